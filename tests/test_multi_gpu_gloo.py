"""N > 1 host logic on CPU: world_size-2 gloo. The window-split MSM's partition / all-gather / combine path is
exercised with a stand-in for the per-rank partial MSM (the oracle's double-and-add restricted to the rank's windows);
the CUDA partials themselves are covered by tests/test_gpu_primitives.py::test_msm_window_split_matches_full."""
import os
import random
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _signed_digits(s, c):
    nwin = (254 + c) // c
    out, carry = [], 0
    for w in range(nwin):
        raw = ((s >> (c * w)) & ((1 << c) - 1)) + carry
        if raw > (1 << (c - 1)):
            out.append(raw - (1 << c))
            carry = 1
        else:
            out.append(raw)
            carry = 0
    assert carry == 0
    return out


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "gnark-plonky2-verifier_b200"))
    import gpw
    from gpw import sharded
    from oracle import bn254 as ob
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    rng = random.Random(7)
    n, c = 24, 16
    ks = [rng.randrange(1, 1 << 20) for _ in range(n)]
    scalars = [rng.randrange(ob.R) for _ in range(n)]
    pts = [ob.ec_mul(1, ob.G1_GEN, k) for k in ks]

    def local_partial(lo, hi):      # stand-in for gpw_msm_g1_dev(win_lo, win_hi): sum_{w in [lo,hi)} 2^(c w) sum_i d_iw P_i
        acc = None
        for s, p in zip(scalars, pts):
            d = _signed_digits(s, c)
            k = sum(d[w] << (c * w) for w in range(lo, hi)) % ob.R
            acc = ob.ec_add(1, acc, ob.ec_mul(1, p, k))
        return gpw.ints_to_points(1, [ob.point_key(1, acc)])[0]

    res = sharded.sharded_msm(None, 1, 0, 0, n, window_bits=c, dist=dist, local_partial=local_partial)
    exp = ob.point_key(1, ob.ec_mul(1, ob.G1_GEN, sum(s * k for s, k in zip(scalars, ks)) % ob.R))
    q.put((rank, gpw.points_to_ints(1, res)[0] == exp))
    dist.barrier()
    dist.destroy_process_group()


def test_window_split_msm_two_ranks_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=240) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(results) == [(0, True), (1, True)]


def test_window_ranges_cover_all_windows():
    sys.path.insert(0, os.path.join(ROOT, "gnark-plonky2-verifier_b200"))
    from gpw import sharded
    for c in (8, 13, 16):
        for world in (1, 2, 3, 4, 8, 32):
            nwin, rs = sharded.window_ranges(c, world)
            cover = [w for r in rs if r for w in range(*r)]
            assert cover == list(range(nwin)) and len(rs) == world
