"""CPU tests of the host paths of the shared host/device arithmetic (ff.cuh / ec.cuh / gl.cuh) through the
C ABI, against Python big-int arithmetic (oracle/bn254.py). No GPU calls."""
import random

import numpy as np
import pytest

import gpw
from oracle import bn254 as ob

MODS = {0: gpw.R_MOD, 1: gpw.P_MOD}


def _rand_elems(rng, mod, n):
    vals = [rng.randrange(mod) for _ in range(n)]
    vals[:4] = [0, 1, mod - 1, mod - 2]
    return vals


@pytest.mark.parametrize("field", [0, 1])
@pytest.mark.parametrize("impl", [0, 1])
def test_mont_mul_matches_bigint(field, impl):
    # impl 0 = the even/odd IMAD.WIDE schedule the GPU runs (host emulation), impl 1 = plain CIOS
    rng = random.Random(100 + field)
    mod = MODS[field]
    a = _rand_elems(rng, mod, 2000)
    b = list(reversed(_rand_elems(rng, mod, 2000)))
    am = gpw.host_ff_to_mont(field, gpw.ints_to_limbs(a))
    bm = gpw.host_ff_to_mont(field, gpw.ints_to_limbs(b))
    assert gpw.limbs_to_ints(gpw.host_ff_from_mont(field, am)) == a
    prod = gpw.host_ff_from_mont(field, gpw.host_ff_mul(field, impl, am, bm))
    assert gpw.limbs_to_ints(prod) == [x * y % mod for x, y in zip(a, b)]


@pytest.mark.parametrize("field", [0, 1])
def test_dedicated_squaring_matches_bigint(field):
    # mont_sqr_wide (36 + 64 IMAD.WIDE: cross products once, doubled, plain REDC) - host emulation of the GPU schedule.
    # Limb patterns of all-ones / single bits exercise every carry of the cross-sum rows, the merge and the doubling.
    rng = random.Random(300 + field)
    mod = MODS[field]
    a = _rand_elems(rng, mod, 3000)
    special = [mod // 2, mod // 2 + 1, (1 << 253) - 1, (1 << 224) - 1, (1 << 192) - 1, (1 << 32) - 1, 1 << 31, 1 << 63, 1 << 252,
               0xffffffff00000000ffffffff00000000ffffffff00000000ffffffff % mod,
               0x00000000ffffffff00000000ffffffff00000000ffffffff00000000ffffffff % mod]
    for k in range(8):
        special.append(0xffffffff << (32 * k) if k < 7 else 0x0fffffff << 224)
        special.append(((1 << 254) - 1) ^ (0xffffffff << (32 * k)))
    special = [x % mod for x in special]
    a[4:4 + len(special)] = special
    # canonical inputs near the modulus matter, and so do Montgomery forms: feed the raw integers as limbs as well
    raw = gpw.ints_to_limbs(a)
    am = gpw.host_ff_to_mont(field, raw)
    for limbs in (am, raw):
        sq = gpw.limbs_to_ints(gpw.host_ff_mul(field, 2, limbs, limbs))
        ref = gpw.limbs_to_ints(gpw.host_ff_mul(field, 1, limbs, limbs))
        assert sq == ref
    got = gpw.limbs_to_ints(gpw.host_ff_from_mont(field, gpw.host_ff_mul(field, 2, am, am)))
    assert got == [x * x % mod for x in a]


@pytest.mark.parametrize("field", [0, 1])
def test_dual_product_mul_matches_bigint(field):
    # a b - c d under ONE Montgomery reduction (mont_mul2): extreme operands exercise the accumulator's head-room
    rng = random.Random(200 + field)
    mod = MODS[field]
    n = 3000
    ext = [0, 1, mod - 1, mod - 2, mod // 2]
    vals = [[rng.choice(ext) if i < 700 else rng.randrange(mod) for i in range(n)] for _ in range(4)]
    for i in range(len(ext) ** 4):          # every combination of the extreme values
        for k in range(4):
            vals[k][700 + i] = ext[(i // len(ext) ** k) % len(ext)]
    mont = [gpw.host_ff_to_mont(field, gpw.ints_to_limbs(v)) for v in vals]
    out = gpw.limbs_to_ints(gpw.host_ff_from_mont(field, gpw.host_ff_mul_sub2(field, *mont)))
    a, b, c, d = vals
    assert out == [(a[i] * b[i] - c[i] * d[i]) % mod for i in range(n)]


@pytest.mark.parametrize("field", [0, 1])
def test_montgomery_constants(field):
    # to_mont(1) must equal R mod p from SURVEY A.1
    mod = MODS[field]
    one = gpw.limbs_to_ints(gpw.host_ff_to_mont(field, gpw.ints_to_limbs([1])))[0]
    assert one == (1 << 256) % mod


@pytest.mark.parametrize("field", [0, 1])
def test_inverse(field):
    rng = random.Random(7)
    mod = MODS[field]
    a = _rand_elems(rng, mod, 50)
    am = gpw.host_ff_to_mont(field, gpw.ints_to_limbs(a))
    inv = gpw.limbs_to_ints(gpw.host_ff_from_mont(field, gpw.host_ff_inv(field, am)))
    assert inv == [pow(x, mod - 2, mod) for x in a]


def test_oracle_curve_constants():
    assert ob.ec_on_curve(1, ob.G1_GEN) and ob.ec_on_curve(2, ob.G2_GEN)
    assert ob.ec_mul(1, ob.G1_GEN, ob.R) is None or ob.ec_mul(1, ob.G1_GEN, ob.R - 1) == ob.ec_neg(1, ob.G1_GEN)
    assert ob.ec_add(2, ob.ec_mul(2, ob.G2_GEN, ob.R - 1), ob.G2_GEN) is None
    assert pow(ob.root_of_unity(5), 32, ob.R) == 1 and pow(ob.root_of_unity(5), 16, ob.R) != 1


@pytest.mark.parametrize("group", [1, 2])
def test_generator_multiples_and_scalar_mul(group):
    pts = gpw.host_ec_generator_multiples(group, 5, 6)
    gen = ob.G1_GEN if group == 1 else ob.G2_GEN
    got = gpw.points_to_ints(group, pts)
    for i, g in enumerate(got):
        assert g == ob.point_key(group, ob.ec_mul(group, gen, 5 + i))
        assert gpw.host_ec_is_on_curve(group, pts[i])
    k = 0x1234567890abcdef1234567890abcdef1234567890abcdef
    q = gpw.host_ec_scalar_mul(group, pts[0], k)
    assert gpw.points_to_ints(group, q)[0] == ob.point_key(group, ob.ec_mul(group, gen, 5 * k))
    # addition incl. doubling and cancellation
    s = gpw.host_ec_add(group, pts[0], pts[1])
    assert gpw.points_to_ints(group, s)[0] == ob.point_key(group, ob.ec_mul(group, gen, 11))
    d = gpw.host_ec_add(group, pts[2], pts[2])
    assert gpw.points_to_ints(group, d)[0] == ob.point_key(group, ob.ec_mul(group, gen, 14))
    neg = gpw.ints_to_points(group, [ob.point_key(group, ob.ec_neg(group, ob.ec_mul(group, gen, 7)))])[0]
    z = gpw.host_ec_add(group, pts[2], neg)
    assert not z.any()
    bad = pts[0].copy()
    bad[0] ^= 1
    assert not gpw.host_ec_is_on_curve(group, bad)
