#!/usr/bin/env python3
"""Quick device-resident timing probe of the hot kernels (development aid; bench.py is the contract)."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gnark-plonky2-verifier_b200"))
import gpw  # noqa: E402


def rand_scalars(n, kind, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    s = torch.randint(0, 1 << 62, (n, 4), dtype=torch.int64, device="cuda", generator=g)
    s[:, 3] &= (1 << 59) - 1      # < 2^251 < r
    if kind == "witness":          # 15% 0/1, 20% < 2^16, 45% < 2^64, 20% full
        u = torch.rand(n, device="cuda", generator=g)
        small = u < 0.80
        s[small, 1:] = 0
        s[u < 0.35, 0] &= 0xffff
        s[u < 0.15, 0] &= 1
    return s


def main():
    logns = [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "20,23").split(",")]
    ctx = gpw.Context(0)
    side = torch.cuda.Stream()
    torch.cuda.set_stream(side)     # a real (non-NULL) stream, shared by torch and libgpw
    ctx.set_stream(side.cuda_stream)
    for logn in logns:
        n = 1 << logn
        for group, words in ((1, 8), (2, 16)):
            if group == 2 and logn > 22:
                continue
            pts = torch.empty((n, words), dtype=torch.int64, device="cuda")
            t0 = time.time()
            ctx.generator_multiples_dev(group, 1, n, pts.data_ptr())
            torch.cuda.synchronize()
            tgen = time.time() - t0
            for kind in ("uniform", "witness"):
                s = rand_scalars(n, kind, 1)
                torch.cuda.synchronize()
                for c in (0, 14):
                    ctx.msm_dev(group, s.data_ptr(), pts.data_ptr(), n, window_bits=c)   # warm (allocs)
                    t0 = time.time()
                    ctx.msm_dev(group, s.data_ptr(), pts.data_ptr(), n, window_bits=c)
                    dt = time.time() - t0
                    st = ctx.msm_last_stats()
                    bytes_alg = n * (32 + words * 8)
                    print("MSM G%d n=2^%d %-8s c=%2d  total %.2f ms (acc %.2f ms, host wall %.2f ms) digits=%d  alg GB/s=%.1f  gen=%.2fs"
                          % (group, logn, kind, c, st["total_ms"], st["accumulate_ms"], dt * 1e3, st["nonzero_digits"],
                             bytes_alg / st["total_ms"] / 1e6, tgen), flush=True)
            del pts
        a = rand_scalars(n, "uniform", 3)
        ctx.ntt_dev(a.data_ptr(), logn, out_bitrev=True)
        torch.cuda.synchronize()
        for (inv, coset, ib, ob, name) in ((0, 0, 0, 1, "fwd DIF"), (1, 1, 1, 0, "inv coset DIT"), (0, 1, 0, 0, "fwd coset nat->nat")):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                ctx.ntt_dev(a.data_ptr(), logn, inverse=inv, coset=coset, in_bitrev=ib, out_bitrev=ob)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 5
            print("NTT 2^%d %-20s %.3f ms  alg GB/s=%.1f" % (logn, name, ms, 64 * n / ms / 1e6), flush=True)


if __name__ == "__main__":
    main()
