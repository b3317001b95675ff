"""ctypes binding of libgpw.so - the C-ABI drop-in boundary (include/gpw.h).

This module is plumbing only: it loads the in-tree shared library, checks status codes and moves
numpy buffers across the boundary. There is NO fallback: if the library or a CUDA device is
missing, calls raise.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GPW_LIB") or os.path.join(os.path.dirname(_HERE), "libgpw.so")  # (GPW_LIB: development builds)

R_MOD = 21888242871839275222246405745257275088548364400416034343698204186575808495617
P_MOD = 21888242871839275222246405745257275088696311157297823662689037894645226208583
GL_P = (1 << 64) - (1 << 32) + 1


class GpwError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("libgpw error %d: %s" % (code, msg))
        self.code = code


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError("libgpw.so not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "or `make -C gnark-plonky2-verifier_b200/csrc` (no CPU fallback exists)")
    return C.CDLL(LIB_PATH)


_lib = _load()
_u64p = C.POINTER(C.c_uint64)
_vp = C.c_void_p

# every symbol declared in include/gpw.h: name -> (restype, argtypes)
SYMBOLS = {
    "gpw_version": (C.c_int, []),
    "gpw_last_error": (C.c_char_p, []),
    "gpw_device_count": (C.c_int, []),
    "gpw_ctx_create": (C.c_int, [C.c_int, C.POINTER(_vp)]),
    "gpw_ctx_destroy": (None, [_vp]),
    "gpw_ctx_set_stream": (C.c_int, [_vp, _vp]),
    "gpw_ctx_sync": (C.c_int, [_vp]),
    "gpw_ctx_set_option": (C.c_int, [_vp, C.c_char_p, C.c_int64]),
    "gpw_comm_unique_id": (C.c_int, [_vp]),
    "gpw_comm_init": (C.c_int, [_vp, C.c_int, C.c_int, _vp]),
    "gpw_comm_destroy": (C.c_int, [_vp]),
    "gpw_comm_info": (C.c_int, [_vp, _vp]),
    "gpw_msm_g1_sharded": (C.c_int, [_vp, C.c_uint64, C.c_uint64, C.c_size_t, C.c_int, C.c_int, C.c_int, _vp]),
    "gpw_msm_g2_sharded": (C.c_int, [_vp, C.c_uint64, C.c_uint64, C.c_size_t, C.c_int, C.c_int, C.c_int, _vp]),
    "gpw_msm_sharded_partial": (C.c_int, [_vp, C.c_int, C.c_uint64, C.c_uint64, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_int,
                                          C.c_int, _vp]),
    "gpw_ctx_launch_count": (C.c_uint64, [_vp]),
    "gpw_host_ff_mul": (C.c_int, [C.c_int, C.c_int, _vp, _vp, _vp, C.c_size_t]),
    "gpw_host_ff_mul_sub2": (C.c_int, [C.c_int, _vp, _vp, _vp, _vp, _vp, C.c_size_t]),
    "gpw_host_ff_to_mont": (C.c_int, [C.c_int, _vp, _vp, C.c_size_t]),
    "gpw_host_ff_from_mont": (C.c_int, [C.c_int, _vp, _vp, C.c_size_t]),
    "gpw_host_ff_inv": (C.c_int, [C.c_int, _vp, _vp, C.c_size_t]),
    "gpw_host_ff_inv_euclid": (C.c_int, [C.c_int, _vp, _vp, C.c_size_t]),
    "gpw_host_ec_scalar_mul": (C.c_int, [C.c_int, _vp, _vp, _vp]),
    "gpw_host_ec_add": (C.c_int, [C.c_int, _vp, _vp, _vp]),
    "gpw_host_ec_is_on_curve": (C.c_int, [C.c_int, _vp]),
    "gpw_host_ec_generator_multiples": (C.c_int, [C.c_int, C.c_uint64, C.c_size_t, _vp]),
    "gpw_selftest_ff": (C.c_int, [_vp, C.c_size_t, C.c_uint64]),
    "gpw_msm_g1": (C.c_int, [_vp, _vp, _vp, C.c_size_t, C.c_int, C.c_int, _vp]),
    "gpw_msm_g2": (C.c_int, [_vp, _vp, _vp, C.c_size_t, C.c_int, C.c_int, _vp]),
    "gpw_msm_g1_dev": (C.c_int, [_vp, C.c_uint64, C.c_uint64, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_int, _vp]),
    "gpw_msm_g2_dev": (C.c_int, [_vp, C.c_uint64, C.c_uint64, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_int, _vp]),
    "gpw_msm_g1_fixed_table": (C.c_int, [_vp, C.c_uint64, C.c_size_t, C.c_int, C.c_int, C.c_uint64]),
    "gpw_msm_g1_fixed_dev": (C.c_int, [_vp, C.c_uint64, C.c_uint64, C.c_size_t, C.c_int, C.c_int, C.c_int, _vp]),
    "gpw_msm_last_stats": (C.c_int, [_vp, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_uint64)]),
    "gpw_ntt_fr": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    "gpw_ntt_fr_dev": (C.c_int, [_vp, C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    "gpw_fr_h_pointwise_dev": (C.c_int, [_vp, C.c_uint64, C.c_uint64, C.c_uint64, C.c_size_t, _vp]),
    "gpw_fr_convert_dev": (C.c_int, [_vp, C.c_uint64, C.c_size_t, C.c_int]),
    "gpw_ec_generator_multiples_dev": (C.c_int, [_vp, C.c_int, C.c_uint64, C.c_size_t, C.c_uint64]),
    "gpw_groth16_pk_synthetic": (C.c_int, [_vp, C.c_size_t, C.c_size_t, C.c_int, C.c_uint64, C.POINTER(_vp)]),
    "gpw_groth16_pk_free": (None, [_vp]),
    "gpw_groth16_pk_info": (C.c_int, [_vp, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_int)]),
    "gpw_groth16_compute_h_dev": (C.c_int, [_vp, C.c_uint64, C.c_uint64, C.c_uint64, C.c_int]),
    "gpw_groth16_prove_dev": (C.c_int, [_vp, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, _vp, _vp, _vp]),
    "gpw_groth16_last_stats": (C.c_int, [_vp, C.POINTER(C.c_float), C.POINTER(C.c_float)]),
    "gpw_poseidon_bn254": (C.c_int, [_vp, _vp, _vp, C.c_size_t, C.c_int]),
    "gpw_poseidon_bn254_dev": (C.c_int, [_vp, C.c_uint64, C.c_uint64, C.c_size_t, C.c_int]),
    "gpw_merkle_paths_bn254": (C.c_int, [_vp, _vp, _vp, _vp, C.c_size_t, C.c_int, _vp]),
    "gpw_hash_or_noop_bn254": (C.c_int, [_vp, _vp, C.c_size_t, C.c_int, _vp]),
    "gpw_gl_mul_add_hint": (C.c_int, [_vp, _vp, _vp, _vp, C.c_size_t, _vp, _vp]),
    "gpw_gl_reduce_hint": (C.c_int, [_vp, _vp, C.c_size_t, _vp, _vp]),
    "gpw_gl_inverse_hint": (C.c_int, [_vp, _vp, C.c_size_t, _vp]),
    "gpw_gl_split_limbs_hint": (C.c_int, [_vp, _vp, C.c_size_t, _vp, _vp]),
    "gpw_poseidon_gl": (C.c_int, [_vp, _vp, _vp, C.c_size_t]),
}

for _name, (_res, _args) in SYMBOLS.items():
    _fn = getattr(_lib, _name)   # AttributeError here = header/library drift
    _fn.restype = _res
    _fn.argtypes = _args


def comm_unique_id() -> bytes:
    """ncclGetUniqueId (rank 0 calls it and ships the 128 bytes to the other ranks)"""
    buf = np.zeros(128, dtype=np.uint8)
    _check(_lib.gpw_comm_unique_id(_p(buf)))
    return buf.tobytes()


def last_error():
    return _lib.gpw_last_error().decode()


def _check(rc):
    if rc != 0:
        raise GpwError(rc, last_error())


def _p(a):
    return a.ctypes.data_as(_vp) if a is not None else None


def _u64(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.uint64)
    if shape is not None:
        a = a.reshape(shape)
    return a


# ---- integer <-> limb helpers (host-side test/bench plumbing) ----------------------------------------
def ints_to_limbs(vals, nlimbs=4):
    out = np.zeros((len(vals), nlimbs), dtype=np.uint64)
    mask = (1 << 64) - 1
    for i, v in enumerate(vals):
        for k in range(nlimbs):
            out[i, k] = (v >> (64 * k)) & mask
    return out


def limbs_to_ints(arr):
    arr = np.asarray(arr, dtype=np.uint64).reshape(-1, arr.shape[-1])
    return [sum(int(row[k]) << (64 * k) for k in range(arr.shape[1])) for row in arr]


def device_count():
    return _lib.gpw_device_count()


def version():
    return _lib.gpw_version()


# ---- host arithmetic ---------------------------------------------------------------------------------
def host_ff_mul(field, impl, a, b):
    a, b = _u64(a, (-1, 4)), _u64(b, (-1, 4))
    out = np.empty_like(a)
    _check(_lib.gpw_host_ff_mul(field, impl, _p(a), _p(b), _p(out), a.shape[0]))
    return out


def host_ff_mul_sub2(field, a, b, c, d):
    a, b, c, d = (_u64(x, (-1, 4)) for x in (a, b, c, d))
    out = np.empty_like(a)
    _check(_lib.gpw_host_ff_mul_sub2(field, _p(a), _p(b), _p(c), _p(d), _p(out), a.shape[0]))
    return out


def host_ff_to_mont(field, a):
    a = _u64(a, (-1, 4))
    out = np.empty_like(a)
    _check(_lib.gpw_host_ff_to_mont(field, _p(a), _p(out), a.shape[0]))
    return out


def host_ff_from_mont(field, a):
    a = _u64(a, (-1, 4))
    out = np.empty_like(a)
    _check(_lib.gpw_host_ff_from_mont(field, _p(a), _p(out), a.shape[0]))
    return out


def host_ff_inv(field, a):
    a = _u64(a, (-1, 4))
    out = np.empty_like(a)
    _check(_lib.gpw_host_ff_inv(field, _p(a), _p(out), a.shape[0]))
    return out


def host_ff_inv_euclid(field, a):
    a = _u64(a, (-1, 4))
    out = np.empty_like(a)
    _check(_lib.gpw_host_ff_inv_euclid(field, _p(a), _p(out), a.shape[0]))
    return out


def _pt_words(group):
    return 8 if group == 1 else 16


def host_ec_scalar_mul(group, point, scalar_int):
    point = _u64(point, (_pt_words(group),))
    k = ints_to_limbs([scalar_int])[0]
    out = np.empty_like(point)
    _check(_lib.gpw_host_ec_scalar_mul(group, _p(point), _p(k), _p(out)))
    return out


def host_ec_add(group, p, q):
    p, q = _u64(p, (_pt_words(group),)), _u64(q, (_pt_words(group),))
    out = np.empty_like(p)
    _check(_lib.gpw_host_ec_add(group, _p(p), _p(q), _p(out)))
    return out


def host_ec_is_on_curve(group, p):
    p = _u64(p, (_pt_words(group),))
    return _lib.gpw_host_ec_is_on_curve(group, _p(p)) == 1


def host_ec_generator_multiples(group, k0, n):
    out = np.empty((n, _pt_words(group)), dtype=np.uint64)
    _check(_lib.gpw_host_ec_generator_multiples(group, k0, n, _p(out)))
    return out


def points_to_ints(group, pts):
    """affine Montgomery limbs -> list of python coordinates (canonical ints); G2 as ((x0,x1),(y0,y1))"""
    pts = _u64(pts, (-1, _pt_words(group)))
    flat = host_ff_from_mont(1, pts.reshape(-1, 4))
    vals = limbs_to_ints(flat)
    out = []
    step = 2 if group == 1 else 4
    for i in range(0, len(vals), step):
        v = vals[i:i + step]
        out.append((v[0], v[1]) if group == 1 else ((v[0], v[1]), (v[2], v[3])))
    return out


def ints_to_points(group, coords):
    flat = []
    for c in coords:
        flat += [c[0], c[1]] if group == 1 else [c[0][0], c[0][1], c[1][0], c[1][1]]
    return host_ff_to_mont(1, ints_to_limbs(flat)).reshape(-1, _pt_words(group))


# ---- GPU context ---------------------------------------------------------------------------------------
class Context:
    """One per GPU (gpw_ctx)."""

    def __init__(self, device=0):
        h = _vp()
        _check(_lib.gpw_ctx_create(device, C.byref(h)))
        self._h = h
        self.device = device

    def close(self):
        if self._h:
            _lib.gpw_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, cuda_stream_ptr):
        _check(_lib.gpw_ctx_set_stream(self._h, _vp(cuda_stream_ptr)))

    def sync(self):
        _check(_lib.gpw_ctx_sync(self._h))

    def set_option(self, key, value):
        _check(_lib.gpw_ctx_set_option(self._h, key.encode(), int(value)))

    # ---- one MSM over several GPUs (csrc/comm.cu) ----
    def comm_init(self, nranks, rank, unique_id: bytes):
        assert len(unique_id) == 128
        buf = np.frombuffer(unique_id, dtype=np.uint8).copy()
        _check(_lib.gpw_comm_init(self._h, nranks, rank, _p(buf)))

    def comm_destroy(self):
        _check(_lib.gpw_comm_destroy(self._h))

    def comm_info(self):
        a = np.zeros(3, dtype=np.int32)
        _check(_lib.gpw_comm_info(self._h, _p(a)))
        return {"ranks": int(a[0]), "rank": int(a[1]), "nccl_version": int(a[2])}

    def msm_sharded(self, group, scalars_ptr, points_ptr, n, scalars_mont=False, window_bits=0, split=0):
        """collective: the whole MSM on every rank (split 1 = windows, 2 = points, 0 = choose)"""
        out = np.zeros(_pt_words(group), dtype=np.uint64)
        fn = _lib.gpw_msm_g1_sharded if group == 1 else _lib.gpw_msm_g2_sharded
        _check(fn(self._h, scalars_ptr, points_ptr, n, int(scalars_mont), window_bits, split, _p(out)))
        return out

    def msm_sharded_partial(self, group, scalars_ptr, points_ptr, n, split, rank, nranks, scalars_mont=False, window_bits=0):
        out = np.zeros(_pt_words(group), dtype=np.uint64)
        _check(_lib.gpw_msm_sharded_partial(self._h, group, scalars_ptr, points_ptr, n, int(scalars_mont), window_bits, split,
                                            rank, nranks, _p(out)))
        return out

    @property
    def launches(self):
        return int(_lib.gpw_ctx_launch_count(self._h))

    def selftest_ff(self, n=4096, seed=1):
        _check(_lib.gpw_selftest_ff(self._h, n, seed))

    # -- MSM ---------------------------------------------------------------------------------------
    def msm(self, group, scalars, points, scalars_mont=False, window_bits=0):
        scalars = _u64(scalars, (-1, 4))
        points = _u64(points, (-1, _pt_words(group)))
        assert scalars.shape[0] == points.shape[0]
        out = np.zeros(_pt_words(group), dtype=np.uint64)
        fn = _lib.gpw_msm_g1 if group == 1 else _lib.gpw_msm_g2
        _check(fn(self._h, _p(scalars), _p(points), scalars.shape[0], int(scalars_mont), window_bits, _p(out)))
        return out

    def msm_dev(self, group, scalars_ptr, points_ptr, n, scalars_mont=False, window_bits=0, win_lo=0, win_hi=0):
        out = np.zeros(_pt_words(group), dtype=np.uint64)
        fn = _lib.gpw_msm_g1_dev if group == 1 else _lib.gpw_msm_g2_dev
        _check(fn(self._h, scalars_ptr, points_ptr, n, int(scalars_mont), window_bits, win_lo, win_hi, _p(out)))
        return out

    @staticmethod
    def msm_fixed_windows(window_bits):
        return (254 + window_bits) // window_bits

    def msm_g1_fixed_table(self, points_ptr, n, window_bits, table_ptr):
        """table (device, msm_fixed_windows(window_bits) * n G1 affine points) <- 2^(window_bits w) P_i"""
        _check(_lib.gpw_msm_g1_fixed_table(self._h, points_ptr, n, window_bits, self.msm_fixed_windows(window_bits), table_ptr))

    def msm_g1_fixed_dev(self, scalars_ptr, table_ptr, n, window_bits, scalars_mont=False):
        out = np.zeros(_pt_words(1), dtype=np.uint64)
        _check(_lib.gpw_msm_g1_fixed_dev(self._h, scalars_ptr, table_ptr, n, int(scalars_mont), window_bits,
                                         self.msm_fixed_windows(window_bits), _p(out)))
        return out

    def msm_last_stats(self):
        a, t, d = C.c_float(), C.c_float(), C.c_uint64()
        _check(_lib.gpw_msm_last_stats(self._h, C.byref(a), C.byref(t), C.byref(d)))
        return {"accumulate_ms": a.value, "total_ms": t.value, "nonzero_digits": d.value}

    # -- NTT ---------------------------------------------------------------------------------------
    def ntt(self, data, inverse=False, coset=False, in_bitrev=False, out_bitrev=False):
        data = _u64(data, (-1, 4)).copy()
        n = data.shape[0]
        logn = n.bit_length() - 1
        assert 1 << logn == n
        _check(_lib.gpw_ntt_fr(self._h, _p(data), logn, int(inverse), int(coset), int(in_bitrev), int(out_bitrev)))
        return data

    def ntt_dev(self, data_ptr, logn, inverse=False, coset=False, in_bitrev=False, out_bitrev=False):
        _check(_lib.gpw_ntt_fr_dev(self._h, data_ptr, logn, int(inverse), int(coset), int(in_bitrev), int(out_bitrev)))

    def h_pointwise_dev(self, a_ptr, b_ptr, c_ptr, n, k_mont):
        k = _u64(k_mont, (4,))
        _check(_lib.gpw_fr_h_pointwise_dev(self._h, a_ptr, b_ptr, c_ptr, n, _p(k)))

    def fr_convert_dev(self, ptr, n, to_mont=True):
        _check(_lib.gpw_fr_convert_dev(self._h, ptr, n, int(to_mont)))

    def generator_multiples_dev(self, group, k0, n, out_ptr):
        _check(_lib.gpw_ec_generator_multiples_dev(self._h, group, k0, n, out_ptr))

    def compute_h_dev(self, a_ptr, b_ptr, c_ptr, logn):
        _check(_lib.gpw_groth16_compute_h_dev(self._h, a_ptr, b_ptr, c_ptr, logn))

    def groth16_pk_synthetic(self, m, n_pub, logn, seed=0):
        return ProvingKey(self, m, n_pub, logn, seed)

    # -- Poseidon / Merkle ---------------------------------------------------------------------------
    def poseidon_bn254(self, states, mont=False):
        states = _u64(states, (-1, 16))
        out = np.empty_like(states)
        _check(_lib.gpw_poseidon_bn254(self._h, _p(states), _p(out), states.shape[0], int(mont)))
        return out

    def poseidon_bn254_dev(self, in_ptr, out_ptr, n, mont=True):
        _check(_lib.gpw_poseidon_bn254_dev(self._h, in_ptr, out_ptr, n, int(mont)))

    def merkle_paths_bn254(self, leaf_digests, siblings, index_bits, depth):
        leaf_digests = _u64(leaf_digests, (-1, 4))
        n = leaf_digests.shape[0]
        siblings = _u64(siblings, (n, depth, 4)) if depth else np.zeros((n, 0, 4), dtype=np.uint64)
        index_bits = _u64(index_bits, (n,))
        roots = np.empty((n, 4), dtype=np.uint64)
        _check(_lib.gpw_merkle_paths_bn254(self._h, _p(leaf_digests), _p(siblings), _p(index_bits), n, depth, _p(roots)))
        return roots

    def hash_or_noop_bn254(self, leaves):
        leaves = _u64(leaves)
        assert leaves.ndim == 2
        n, leaf_len = leaves.shape
        out = np.empty((n, 4), dtype=np.uint64)
        _check(_lib.gpw_hash_or_noop_bn254(self._h, _p(leaves), n, leaf_len, _p(out)))
        return out

    def poseidon_gl(self, states):
        states = _u64(states, (-1, 12))
        out = np.empty_like(states)
        _check(_lib.gpw_poseidon_gl(self._h, _p(states), _p(out), states.shape[0]))
        return out

    # -- Goldilocks hints ----------------------------------------------------------------------------
    def gl_mul_add_hint(self, a, b, c):
        a, b, c = _u64(a), _u64(b), _u64(c)
        q, r = np.empty_like(a), np.empty_like(a)
        _check(_lib.gpw_gl_mul_add_hint(self._h, _p(a), _p(b), _p(c), a.size, _p(q), _p(r)))
        return q, r

    def gl_reduce_hint(self, x4):
        x4 = _u64(x4, (-1, 4))
        q = np.empty_like(x4)
        r = np.empty(x4.shape[0], dtype=np.uint64)
        _check(_lib.gpw_gl_reduce_hint(self._h, _p(x4), x4.shape[0], _p(q), _p(r)))
        return q, r

    def gl_inverse_hint(self, x):
        x = _u64(x)
        out = np.empty_like(x)
        _check(_lib.gpw_gl_inverse_hint(self._h, _p(x), x.size, _p(out)))
        return out

    def gl_split_limbs_hint(self, x):
        x = _u64(x)
        hi, lo = np.empty_like(x), np.empty_like(x)
        _check(_lib.gpw_gl_split_limbs_hint(self._h, _p(x), x.size, _p(hi), _p(lo)))
        return hi, lo


class ProvingKey:
    """gpw_pk: device-resident Groth16 proving key (synthetic, known discrete logs)."""

    def __init__(self, ctx, m, n_pub, logn, seed=0):
        h = _vp()
        _check(_lib.gpw_groth16_pk_synthetic(ctx._h, m, n_pub, logn, seed, C.byref(h)))
        self._h, self.ctx, self.m, self.n_pub, self.logn, self.seed = h, ctx, m, n_pub, logn, seed

    def close(self):
        if self._h:
            _lib.gpw_groth16_pk_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def prove_dev(self, w_ptr, a_ptr, b_ptr, c_ptr, r_int, s_int):
        """-> (Ar[8], Bs[16], Krs[8]) affine Montgomery limbs"""
        r = ints_to_limbs([r_int])[0]
        s = ints_to_limbs([s_int])[0]
        out = np.zeros(32, dtype=np.uint64)
        _check(_lib.gpw_groth16_prove_dev(self._h, w_ptr, a_ptr, b_ptr, c_ptr, _p(r), _p(s), _p(out)))
        return out[:8].copy(), out[8:24].copy(), out[24:].copy()

    def last_stats(self):
        h = C.c_float()
        ms = (C.c_float * 5)()
        _check(_lib.gpw_groth16_last_stats(self._h, C.byref(h), ms))
        return {"compute_h_ms": h.value, "msm_ms": dict(zip(("A", "B1", "B2", "K", "Z"), [float(x) for x in ms]))}


from .wrap import Circuit, WrapKey, PlonkKey, hash_to_fr  # noqa: E402,F401  (binds the circuit / wrap part of include/gpw.h)
