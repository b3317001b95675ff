#!/usr/bin/env python3
"""Runs one witness solve (phase 1 + 2) of testdata/step - the target for `ncu -k regex:k_tape_staged`."""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gnark-plonky2-verifier_b200"))
import gpw
d = os.path.join(ROOT, "tests", "golden", "testdata", "step")
rd = lambda f: open(os.path.join(d, f), "rb").read()
ctx = gpw.Context(0)
circ = gpw.Circuit.compile_verifier(ctx, rd("common_circuit_data.json"))
inputs = circ.parse_inputs(rd("proof_with_public_inputs.json"), rd("verifier_only_circuit_data.json"))
inp = torch.from_numpy(inputs.view(np.int64)).cuda()
wires = torch.zeros((circ.n_wires, 4), dtype=torch.int64, device="cuda")
torch.cuda.synchronize()
import time
for it in range(int(sys.argv[1]) if len(sys.argv) > 1 else 1):
    t0 = time.time()
    circ.solve_phase1_dev(inp.data_ptr(), 1, wires.data_ptr(), circ.n_wires)
    ctx.sync()
    print("phase1 %.1f ms" % ((time.time() - t0) * 1e3), flush=True)
circ.solve_phase2_dev([0xabcdef0123456789abcdef], 1, wires.data_ptr(), circ.n_wires)
ctx.sync()
print("unsat", circ.r1cs_eval_dev(wires.data_ptr()))
