"""Multi-GPU plumbing. Two ways the path shards:

  1. independent proofs: one proof stream per GPU, no collective (bench.py --gpus N, weak scaling);
  2. ONE large MSM split across GPUs by Pippenger WINDOW ranges: every rank holds all scalars and bases, computes
     the windows [lo, hi) it owns with gpw_msm_g{1,2}_dev (the C ABI folds them to sum_w 2^(c w) W_w), then ONE
     all-gather of a single affine point per rank (64 B for G1, 128 B for G2) over NCCL / NVLink; every rank adds
     the partial points. EC addition is not an NCCL reduce op, hence all-gather + local adds (exact: the result is
     bit-identical to the single-GPU MSM).

torch.distributed is plumbing only; the partial MSM is the CUDA kernel path of libgpw.
"""
import numpy as np

from . import host_ec_add, _pt_words


def window_ranges(window_bits, world):
    """Contiguous split of the ceil(255 / c) signed-digit windows over `world` ranks; empty ranges are (0, 0)."""
    nwin = (254 + window_bits) // window_bits
    base, extra = divmod(nwin, world)
    out, lo = [], 0
    for r in range(world):
        n = base + (1 if r < extra else 0)
        out.append((lo, lo + n) if n else None)
        lo += n
    return nwin, out


def combine_partials(group, partials):
    """Sum of the gathered affine partial points (host, a handful of group additions)."""
    acc = np.zeros(_pt_words(group), dtype=np.uint64)
    for p in partials:
        acc = host_ec_add(group, acc, np.asarray(p, dtype=np.uint64))
    return acc


def sharded_msm(ctx, group, scalars_ptr, points_ptr, n, window_bits=16, scalars_mont=False, dist=None, device=None,
                local_partial=None):
    """MSM of n points split by windows over dist.get_world_size() ranks. Returns the full result on every rank.
    local_partial(lo, hi) -> affine point limbs: defaults to the CUDA path (ctx.msm_dev); tests on CPU inject a stand-in."""
    import torch
    world = dist.get_world_size() if dist is not None else 1
    rank = dist.get_rank() if dist is not None else 0
    _, ranges = window_ranges(window_bits, world)
    words = _pt_words(group)
    rng = ranges[rank]
    if rng is None:
        mine = np.zeros(words, dtype=np.uint64)
    elif local_partial is not None:
        mine = np.asarray(local_partial(*rng), dtype=np.uint64)
    else:
        mine = ctx.msm_dev(group, scalars_ptr, points_ptr, n, scalars_mont=scalars_mont, window_bits=window_bits,
                           win_lo=rng[0], win_hi=rng[1])
    if world == 1:
        return mine
    t = torch.from_numpy(mine.view(np.int64).copy())
    if device is not None:
        t = t.to(device)
    gathered = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(gathered, t)
    return combine_partials(group, [g.cpu().numpy().view(np.uint64) for g in gathered])
