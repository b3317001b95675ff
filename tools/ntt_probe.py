#!/usr/bin/env python3
"""Development probe: device time of 2^logn NTTs (CUDA events), the four transform kinds of computeH.
usage: ntt_probe.py [logn] [reps]   (GPW_NTT_TMA=0 switches the TMA tile staging off for comparison)"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gnark-plonky2-verifier_b200"))
import gpw  # noqa: E402


def main():
    logn = int(sys.argv[1]) if len(sys.argv) > 1 else 23
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    ctx = gpw.Context(0)
    side = torch.cuda.Stream()
    torch.cuda.set_stream(side)
    ctx.set_stream(side.cuda_stream)
    n = 1 << logn
    g = torch.Generator(device="cuda").manual_seed(5)
    a = torch.randint(0, 1 << 62, (n, 4), dtype=torch.int64, device="cuda", generator=g)
    a[:, 3] &= (1 << 59) - 1
    kinds = {"inverse nat->bitrev": dict(inverse=True, coset=False, in_bitrev=False, out_bitrev=True),
             "coset fwd bitrev->nat": dict(inverse=False, coset=True, in_bitrev=True, out_bitrev=False),
             "coset inv nat->bitrev": dict(inverse=True, coset=True, in_bitrev=False, out_bitrev=True),
             "forward nat->bitrev": dict(inverse=False, coset=False, in_bitrev=False, out_bitrev=True)}
    for name, kw in kinds.items():
        ctx.ntt_dev(a.data_ptr(), logn, **kw)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(side)
        for _ in range(reps):
            ctx.ntt_dev(a.data_ptr(), logn, **kw)
        e1.record(side)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        print("NTT 2^%d %-24s %.3f ms  (%.0f GB/s algorithmic, 64 N bytes)" % (logn, name, ms, 64 * n / ms / 1e6), flush=True)


if __name__ == "__main__":
    main()
