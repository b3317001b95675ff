"""Groth16 prover core (computeH + 5 MSMs + assembly) against the oracle, using the synthetic key's known
discrete logs: every proof element is recomputed 'in the exponent' with Python integers."""
import random

import numpy as np
import pytest

import gpw
from oracle import bn254 as ob

pytestmark = pytest.mark.gpu
R = ob.R


@pytest.fixture(scope="module")
def ctx():
    c = gpw.Context(0)
    yield c
    c.close()


def _to_dev(torch, arr):
    return torch.from_numpy(np.ascontiguousarray(arr).view(np.int64)).cuda()


def _h_coeffs(a_evals, b_evals, c_evals):
    """(A*B - C) / (x^N - 1) by schoolbook polynomial arithmetic on the interpolants."""
    n = len(a_evals)
    A = ob.ntt_fast(a_evals, inverse=True)
    B = ob.ntt_fast(b_evals, inverse=True)
    Cc = ob.ntt_fast(c_evals, inverse=True)
    prod = [0] * (2 * n - 1)
    for i, x in enumerate(A):
        if x:
            for j, y in enumerate(B):
                prod[i + j] = (prod[i + j] + x * y) % R
    for i, x in enumerate(Cc):
        prod[i] = (prod[i] - x) % R
    h = [prod[j + n] for j in range(n - 1)]
    # exact division check: P = h x^N - h
    for k in range(n):
        low = (-(h[k] if k < n - 1 else 0)) % R
        assert prod[k] == low
    return h


@pytest.mark.parametrize("m,n_pub,logn", [(50, 3, 6), (300, 37, 8)])
def test_groth16_prove_matches_exponent_arithmetic(ctx, m, n_pub, logn):
    import torch
    rng = random.Random(m)
    n = 1 << logn
    w = [1] + [rng.randrange(R) if i % 2 else rng.randrange(1 << 16) for i in range(m - 1)]
    a = [rng.randrange(R) for _ in range(n)]
    b = [rng.randrange(R) for _ in range(n)]
    c = [x * y % R for x, y in zip(a, b)]
    h = _h_coeffs(a, b, c)
    seed = 77
    pk = ctx.groth16_pk_synthetic(m, n_pub, logn, seed=seed)
    mont = lambda v: gpw.host_ff_to_mont(0, gpw.ints_to_limbs(v))
    dw, da, db, dc = (_to_dev(torch, mont(v)) for v in (w, a, b, c))
    torch.cuda.synchronize()
    # computeH alone
    da2, db2, dc2 = da.clone(), db.clone(), dc.clone()
    torch.cuda.synchronize()
    ctx.compute_h_dev(da2.data_ptr(), db2.data_ptr(), dc2.data_ptr(), logn)
    ctx.sync()
    got_h = gpw.limbs_to_ints(gpw.host_ff_from_mont(0, da2.cpu().numpy().view(np.uint64)))
    assert got_h[:n - 1] == h and got_h[n - 1] == 0
    r_, s_ = rng.randrange(R), rng.randrange(R)
    Ar, Bs, Krs = pk.prove_dev(dw.data_ptr(), da.data_ptr(), db.data_ptr(), dc.data_ptr(), r_, s_)
    alpha, beta, delta = seed + 1, seed + 2, seed + 3
    sA = (alpha + sum(x * (1 + i) for i, x in enumerate(w)) + r_ * delta) % R
    sB1 = (beta + sum(x * (1 + m + i) for i, x in enumerate(w)) + s_ * delta) % R
    sB2 = (beta + sum(x * (1 + i) for i, x in enumerate(w)) + s_ * delta) % R
    sK = (sum(w[i] * (1 + 2 * m + i) for i in range(n_pub, m)) + sum(hj * (1 + 3 * m + j) for j, hj in enumerate(h))
          + s_ * sA + r_ * sB1 - r_ * s_ * delta) % R
    assert gpw.points_to_ints(1, Ar)[0] == ob.point_key(1, ob.ec_mul(1, ob.G1_GEN, sA))
    assert gpw.points_to_ints(2, Bs)[0] == ob.point_key(2, ob.ec_mul(2, ob.G2_GEN, sB2))
    assert gpw.points_to_ints(1, Krs)[0] == ob.point_key(1, ob.ec_mul(1, ob.G1_GEN, sK))
    pk.close()


def test_generator_multiples_dev_matches_host(ctx):
    import torch
    for group, words in ((1, 8), (2, 16)):
        n = 1000
        out = torch.empty((n, words), dtype=torch.int64, device="cuda")
        ctx.generator_multiples_dev(group, 12345, n, out.data_ptr())
        ctx.sync()
        got = out.cpu().numpy().view(np.uint64)
        exp = gpw.host_ec_generator_multiples(group, 12345, n)
        assert (got == exp).all()
