#!/usr/bin/env python3
"""torchrun target: window-split MSM over all ranks (NCCL all-gather of one affine point per rank) must equal the
single-GPU MSM bit for bit. Usage: torchrun --nproc-per-node N tools/sharded_msm_check.py [log2 n]"""
import os, sys
import numpy as np
import torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gnark-plonky2-verifier_b200"))
import gpw
from gpw import sharded

def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    logn = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    n = 1 << logn
    ctx = gpw.Context(local)
    side = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(side)
    ctx.set_stream(side.cuda_stream)
    for group, words in ((1, 8), (2, 16)):
        m = n if group == 1 else n // 4
        pts = torch.empty((m, words), dtype=torch.int64, device=dev)
        ctx.generator_multiples_dev(group, 1, m, pts.data_ptr())
        g = torch.Generator(device=dev).manual_seed(5)          # same scalars on every rank
        s = torch.randint(0, 1 << 62, (m, 4), dtype=torch.int64, device=dev, generator=g)
        s[:, 3] &= (1 << 59) - 1
        torch.cuda.synchronize()
        full = ctx.msm_dev(group, s.data_ptr(), pts.data_ptr(), m, window_bits=16)
        full = ctx.msm_dev(group, s.data_ptr(), pts.data_ptr(), m, window_bits=16)      # second call: scratch is allocated
        full_ms = ctx.msm_last_stats()["total_ms"]
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        dist.barrier(); torch.cuda.synchronize()
        e0.record(side)
        res = sharded.sharded_msm(ctx, group, s.data_ptr(), pts.data_ptr(), m, window_bits=16, dist=dist, device=dev)
        e1.record(side); torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ok = bool((res == full).all())
        oks = torch.tensor([int(ok)], device=dev)
        dist.all_reduce(oks, op=dist.ReduceOp.MIN)
        if rank == 0:
            st = ctx.msm_last_stats()
            print("sharded MSM G%d n=%d over %d GPUs: %s, %.2f ms incl. all-gather + combine (max over ranks; this rank's window "
                  "share alone %.2f ms; the whole MSM on one GPU %.2f ms)"
                  % (group, m, world, "bit-identical to single-GPU" if oks.item() else "MISMATCH", t.item(), st["total_ms"], full_ms),
                  flush=True)
        assert oks.item() == 1
    dist.destroy_process_group()

if __name__ == "__main__":
    main()
