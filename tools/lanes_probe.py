#!/usr/bin/env python3
"""Development probe: per-proof phase times of gpw_wrap_prove_many for several lane counts (GPW_DEBUG_LANES=1 prints
one line per proof from libgpw). usage: lanes_probe.py [lanes,lanes,...] [n_proofs]"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gnark-plonky2-verifier_b200"))
import gpw  # noqa: E402

TESTDATA = os.path.join(ROOT, "tests", "golden", "testdata", "step")


def main():
    lanes_list = [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "1,2,4").split(",")]
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 12
    ctx = gpw.Context(0)
    rd = lambda f: open(os.path.join(TESTDATA, f), "rb").read()
    circ = gpw.Circuit.compile_verifier(ctx, rd("common_circuit_data.json"))
    key = gpw.WrapKey(ctx, circ, seed=0x5EED)
    inputs = circ.parse_inputs(rd("proof_with_public_inputs.json"), rd("verifier_only_circuit_data.json"))
    many = torch.from_numpy(np.ascontiguousarray(np.tile(inputs, (n, 1, 1))).view(np.int64)).pin_memory()
    for lanes in lanes_list:
        key.set_lanes(lanes)
        key.prove_many(many.data_ptr(), max(lanes, 2), [1] * max(lanes, 2), [2] * max(lanes, 2))
        torch.cuda.synchronize()
        print("=== lanes %d" % lanes, flush=True)
        os.environ["GPW_DEBUG_LANES"] = "1"
        t0 = time.perf_counter()
        key.prove_many(many.data_ptr(), n, [1] * n, [2] * n)
        dt = time.perf_counter() - t0
        os.environ.pop("GPW_DEBUG_LANES")
        print("=== lanes %d: %d proofs in %.1f ms = %.2f proofs/s" % (lanes, n, dt * 1e3, n / dt), flush=True)


if __name__ == "__main__":
    main()
