#!/usr/bin/env python3
"""ncu target: exactly ONE complete wrap proof of testdata/step between cudaProfilerStart / cudaProfilerStop (run ncu with
`--profile-from-start off`), after a warm-up proof, one proof in flight, synthetic key (the kernels do not depend on the base
values). Used for profiles/r02_launches.csv and the per-kernel `--set full` captures."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gnark-plonky2-verifier_b200"))
import gpw  # noqa: E402

d = os.path.join(ROOT, "tests", "golden", "testdata", "step")
rd = lambda f: open(os.path.join(d, f), "rb").read()
ctx = gpw.Context(0)
circ = gpw.Circuit.compile_verifier(ctx, rd("common_circuit_data.json"))
key = gpw.WrapKey(ctx, circ, seed=1)
inputs = circ.parse_inputs(rd("proof_with_public_inputs.json"), rd("verifier_only_circuit_data.json"))
key.prove(inputs, 3, 4)
torch.cuda.synchronize()
torch.cuda.profiler.start()
p = key.prove(inputs, 3, 4)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
assert p["n_unsatisfied"] == 0
print("one wrap profiled; launches so far:", ctx.launches, key.last_stats())
