"""The C oracle (oracle/c/bn254_ref.c - the CPU baseline's MSM / FFT) against the Python big-int oracle (oracle/bn254.py:
schoolbook double-and-add, O(n^2) DFT definition), the curve constants re-derived numerically, and libgpw's host-side
hash-to-field against hashlib."""
import ctypes as C
import hashlib
import os
import random
import subprocess

import numpy as np
import pytest

import gpw
from oracle import bn254 as ob
from oracle import pairing as opair

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
R, P = ob.R, ob.P


@pytest.fixture(scope="module")
def lib():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle", "c"), "libbn254_ref.so"], stdout=subprocess.DEVNULL)
    lib = C.CDLL(os.path.join(ROOT, "oracle", "c", "libbn254_ref.so"))
    vp = C.c_void_p
    lib.ref_msm_g1.argtypes = [vp, vp, C.c_size_t, C.c_int, C.c_int, C.c_int, vp]
    lib.ref_msm_g2.argtypes = [vp, vp, C.c_size_t, C.c_int, C.c_int, C.c_int, vp]
    lib.ref_ntt_fr.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int]
    lib.ref_g1_multiples.argtypes = [vp, C.c_uint64, C.c_size_t, vp]
    lib.ref_g2_multiples.argtypes = [vp, C.c_uint64, C.c_size_t, vp]
    lib.ref_g1_scalar_mul.argtypes = [vp, vp, vp]
    lib.ref_g2_scalar_mul.argtypes = [vp, vp, vp]
    for f in ("ref_fr_to_mont", "ref_fr_from_mont", "ref_fp_to_mont", "ref_fp_from_mont"):
        getattr(lib, f).argtypes = [vp, vp, C.c_size_t]
    return lib


def _limbs(vals):
    return gpw.ints_to_limbs([v for v in vals])


def _fp_mont(lib, vals):
    a = _limbs(vals)
    o = np.zeros_like(a)
    lib.ref_fp_to_mont(a.ctypes.data, o.ctypes.data, len(vals))
    return o


def _fp_ints(lib, mont):
    o = np.zeros_like(mont)
    lib.ref_fp_from_mont(mont.ctypes.data, o.ctypes.data, len(mont))
    return gpw.limbs_to_ints(o)


def _g1_points(lib, buf):
    v = _fp_ints(lib, np.ascontiguousarray(buf.reshape(-1, 4)))
    return [(v[2 * i], v[2 * i + 1]) for i in range(len(v) // 2)]


def _g2_points(lib, buf):
    v = _fp_ints(lib, np.ascontiguousarray(buf.reshape(-1, 4)))
    return [((v[4 * i], v[4 * i + 1]), (v[4 * i + 2], v[4 * i + 3])) for i in range(len(v) // 4)]


def test_constants_rederived():
    # field moduli from the BN parameter, 2-adicity and the root of unity gnark-crypto's fr/fft uses, curve membership of
    # the generators, group order
    assert pow(ob.ROOT_2_28, 1 << 28, R) == 1 and pow(ob.ROOT_2_28, 1 << 27, R) == R - 1
    assert (R - 1) % (1 << 28) == 0 and (R - 1) % (1 << 29) != 0
    assert ob.ec_on_curve(1, ob.G1_GEN) and ob.ec_on_curve(2, ob.G2_GEN)
    assert ob.ec_mul(1, ob.G1_GEN, R - 1) == ob.ec_neg(1, ob.G1_GEN)
    assert opair.ec_mul_full(2, ob.G2_GEN, R) is None


def test_c_generator_multiples_and_scalar_mul_match_python(lib):
    out = np.zeros((5, 8), dtype=np.uint64)
    lib.ref_g1_multiples(None, 3, 5, out.ctypes.data)
    assert _g1_points(lib, out) == [ob.point_key(1, ob.ec_mul(1, ob.G1_GEN, 3 + i)) for i in range(5)]
    g2 = _fp_mont(lib, [ob.G2_GEN[0].a, ob.G2_GEN[0].b, ob.G2_GEN[1].a, ob.G2_GEN[1].b]).reshape(-1)
    out2 = np.zeros((3, 16), dtype=np.uint64)
    lib.ref_g2_multiples(g2.ctypes.data, 7, 3, out2.ctypes.data)
    assert _g2_points(lib, out2) == [ob.point_key(2, ob.ec_mul(2, ob.G2_GEN, 7 + i)) for i in range(3)]
    k = random.Random(2).randrange(R)
    o = np.zeros(8, dtype=np.uint64)
    kl = _limbs([k])
    lib.ref_g1_scalar_mul(out[0].ctypes.data, kl.ctypes.data, o.ctypes.data)
    assert _g1_points(lib, o)[0] == ob.point_key(1, ob.ec_mul(1, ob.G1_GEN, 3 * k))
    o2 = np.zeros(16, dtype=np.uint64)
    lib.ref_g2_scalar_mul(g2.ctypes.data, kl.ctypes.data, o2.ctypes.data)
    assert _g2_points(lib, o2)[0] == ob.point_key(2, ob.ec_mul(2, ob.G2_GEN, k))


@pytest.mark.parametrize("n,c,threads", [(1, 0, 1), (37, 4, 2), (300, 0, 3), (3000, 11, 0), (3000, 16, 0)])
def test_c_msm_g1_matches_schoolbook(lib, n, c, threads):
    rng = random.Random(n)
    pts = np.zeros((n, 8), dtype=np.uint64)
    lib.ref_g1_multiples(None, 1, n, pts.ctypes.data)
    # full-width, small, zero, one, r-1 and repeated scalars (the witness mix)
    sc = [rng.choice((rng.randrange(R), rng.randrange(1 << 16), 0, 1, R - 1)) for _ in range(n)]
    out = np.zeros(8, dtype=np.uint64)
    scl = _limbs(sc)     # (keep a reference: .ctypes.data of a temporary dangles)
    assert lib.ref_msm_g1(scl.ctypes.data, pts.ctypes.data, n, 0, c, threads, out.ctypes.data) == 0
    exp = ob.ec_mul(1, ob.G1_GEN, sum(s * (i + 1) for i, s in enumerate(sc)) % R)
    assert _g1_points(lib, out)[0] == ob.point_key(1, exp)
    if n <= 37:   # the defining computation itself
        pyp = [ob.ec_mul(1, ob.G1_GEN, i + 1) for i in range(n)]
        assert ob.point_key(1, ob.msm_naive(1, sc, pyp)) == ob.point_key(1, exp)
    # Montgomery-form scalars give the same point
    scm = np.zeros((n, 4), dtype=np.uint64)
    lib.ref_fr_to_mont(scl.ctypes.data, scm.ctypes.data, n)
    out2 = np.zeros(8, dtype=np.uint64)
    assert lib.ref_msm_g1(scm.ctypes.data, pts.ctypes.data, n, 1, c, threads, out2.ctypes.data) == 0
    assert (out == out2).all()


@pytest.mark.parametrize("n,c", [(1, 0), (50, 5), (500, 0)])
def test_c_msm_g2_matches_schoolbook(lib, n, c):
    rng = random.Random(100 + n)
    g2 = _fp_mont(lib, [ob.G2_GEN[0].a, ob.G2_GEN[0].b, ob.G2_GEN[1].a, ob.G2_GEN[1].b]).reshape(-1)
    pts = np.zeros((n, 16), dtype=np.uint64)
    lib.ref_g2_multiples(g2.ctypes.data, 1, n, pts.ctypes.data)
    sc = [rng.choice((rng.randrange(R), rng.randrange(1 << 32), 0, R - 1)) for _ in range(n)]
    out = np.zeros(16, dtype=np.uint64)
    scl = _limbs(sc)
    assert lib.ref_msm_g2(scl.ctypes.data, pts.ctypes.data, n, 0, c, 0, out.ctypes.data) == 0
    exp = ob.ec_mul(2, ob.G2_GEN, sum(s * (i + 1) for i, s in enumerate(sc)) % R)
    assert _g2_points(lib, out)[0] == ob.point_key(2, exp)


def test_c_ntt_matches_dft_definition(lib):
    rng = random.Random(9)
    for logn in (0, 1, 2, 5, 8, 15, 16):     # 15 / 16 cross the cache-blocked stage split
        n = 1 << logn
        v = [rng.randrange(R) for _ in range(n)]
        vl = _limbs(v)
        for inverse in (0, 1):
            for coset in (0, 1):
                if logn > 8 and (inverse != coset):
                    continue
                m = np.zeros((n, 4), dtype=np.uint64)
                lib.ref_fr_to_mont(vl.ctypes.data, m.ctypes.data, n)
                assert lib.ref_ntt_fr(m.ctypes.data, logn, inverse, coset, 0) == 0
                o = np.zeros_like(m)
                lib.ref_fr_from_mont(m.ctypes.data, o.ctypes.data, n)
                exp = (ob.dft_naive if logn <= 5 else ob.ntt_fast)(v, inverse=bool(inverse), coset=bool(coset))
                assert gpw.limbs_to_ints(o) == exp, (logn, inverse, coset)
    # the fast Python NTT used above is itself the O(n^2) definition at small sizes
    v = [rng.randrange(R) for _ in range(64)]
    for inverse in (False, True):
        for coset in (False, True):
            assert ob.ntt_fast(v, inverse=inverse, coset=coset) == ob.dft_naive(v, inverse=inverse, coset=coset)


def test_libgpw_hash_to_fr_is_rfc9380_xmd_sha256():
    # gpw_hash_to_fr (csrc/wrap.cu: own SHA-256 + expand_message_xmd) against hashlib, the way gnark-crypto's fr.Hash(msg, dst, 1)
    # derives the BSB22 commitment challenge: 48 bytes of XMD output, big-endian, mod r
    from gpw.wrap import hash_to_fr
    rng = random.Random(4)
    for ln in (0, 1, 55, 56, 63, 64, 65, 119, 120, 200, 1000):
        msg = bytes(rng.randrange(256) for _ in range(ln))
        for dst in (b"bsb22-commitment", b"QUUX-V01-CS02-with-expander-SHA256-128", b"x"):
            assert hash_to_fr(msg, dst) == opair.hash_to_fr(msg, dst), (ln, dst)
    # and the oracle's XMD is the RFC's (appendix K.1 vector)
    assert opair.expand_message_xmd(b"abc", b"QUUX-V01-CS02-with-expander-SHA256-128", 32).hex() == \
        "d8ccab23b5985ccea865c6c97b6e5b8350e794e603b4b97902f53a8a0d605615"
    assert hashlib.sha256(b"").hexdigest().startswith("e3b0c442")
