import os, sys
import numpy as np, torch
sys.path.insert(0, "gnark-plonky2-verifier_b200")
import gpw
d = "tests/golden/testdata/step"
rd = lambda f: open(os.path.join(d, f), "rb").read()
ctx = gpw.Context(0)
circ = gpw.Circuit.compile_verifier(ctx, rd("common_circuit_data.json"))
key = gpw.WrapKey(ctx, circ, seed=1)
inp = circ.parse_inputs(rd("proof_with_public_inputs.json"), rd("verifier_only_circuit_data.json"))
key.prove(inp, 5, 7)
os.environ["GPW_DEBUG_WRAP"] = "1"
key.prove(inp, 5, 7)
print(key.last_stats())
