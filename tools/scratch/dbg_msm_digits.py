"""development probe: per-MSM sizes / digit counts / times of one wrap (GPW_DEBUG_WRAP=1 lines from libgpw)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "gnark-plonky2-verifier_b200"))
import gpw  # noqa: E402

T = os.path.join(ROOT, "tests", "golden", "testdata", "step")
ctx = gpw.Context(0)
rd = lambda f: open(os.path.join(T, f), "rb").read()
circ = gpw.Circuit.compile_verifier(ctx, rd("common_circuit_data.json"))
key = gpw.WrapKey(ctx, circ, seed=1)
inp = circ.parse_inputs(rd("proof_with_public_inputs.json"), rd("verifier_only_circuit_data.json"))
key.prove(inp, 1, 2)
os.environ["GPW_DEBUG_WRAP"] = "1"
key.prove(inp, 1, 2)
