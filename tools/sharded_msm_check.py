#!/usr/bin/env python3
"""torchrun target: ONE MSM split over all ranks INSIDE libgpw (gpw_comm_init + gpw_msm_g{1,2}_sharded: NCCL all-gather of one
affine point per rank on the context's stream, the points added on the device) must equal the single-GPU MSM bit for bit, for
both splits (windows / points), G1 and G2. torch.distributed only ships the 128-byte NCCL id and times the barrier.
Usage: torchrun --nproc-per-node N tools/sharded_msm_check.py [log2 n]"""
import os, sys
import numpy as np
import torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gnark-plonky2-verifier_b200"))
import gpw


def init_comm(ctx, rank, world, dev):
    idt = torch.zeros(128, dtype=torch.uint8, device=dev)
    if rank == 0:
        idt = torch.frombuffer(bytearray(gpw.comm_unique_id()), dtype=torch.uint8).to(dev)
    dist.broadcast(idt, 0)
    ctx.comm_init(world, rank, idt.cpu().numpy().tobytes())


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
    init_comm(ctx, rank, world, dev)
    assert ctx.comm_info()["ranks"] == world
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
        for split, name in ((1, "windows"), (2, "points")):
            ctx.msm_sharded(group, s.data_ptr(), pts.data_ptr(), m, window_bits=16, split=split)   # warm-up
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            dist.barrier(); torch.cuda.synchronize()
            e0.record(side)
            res = ctx.msm_sharded(group, s.data_ptr(), pts.data_ptr(), m, window_bits=16, split=split)
            e1.record(side); torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1)], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            oks = torch.tensor([int(bool((res == full).all()))], device=dev)
            dist.all_reduce(oks, op=dist.ReduceOp.MIN)
            if rank == 0:
                print("sharded MSM G%d n=%d over %d GPUs, split by %s: %s, %.2f ms incl. all-gather + device adds (max over ranks; "
                      "the whole MSM on one GPU %.2f ms)"
                      % (group, m, world, name, "bit-identical to single-GPU" if oks.item() else "MISMATCH", t.item(), full_ms), flush=True)
            assert oks.item() == 1
    ctx.comm_destroy()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
