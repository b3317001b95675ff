"""Host-side mirror of the reference's top-level flow (benchmark.go:27-78, 192-304) over the C ABI:

    circuit = Circuit.compile_verifier(ctx, common_circuit_data_json)      # frontend.Compile          (:55)
    key     = WrapKey(ctx, circuit, seed)                                   # groth16.DummySetup        (:214)
    inputs  = circuit.parse_inputs(proof_json, verifier_only_json)          # variables.Deserialize*    (:30-31)
    proof   = key.prove(inputs, r, s)                                       # NewWitness + groth16.Prove (:240-249)

Plumbing only - all computation happens in libgpw.so on the GPU.
"""
import ctypes as C

import numpy as np

from . import _lib, _check, _p, _vp, ints_to_limbs, SYMBOLS

_NEW = {
    "gpw_circuit_compile_verifier": (C.c_int, [_vp, C.c_char_p, C.POINTER(_vp)]),
    "gpw_circuit_compile_verifier_bound": (C.c_int, [_vp, C.c_char_p, C.c_char_p, C.c_char_p, C.POINTER(_vp)]),
    "gpw_circuit_compile_gadget": (C.c_int, [_vp, C.c_char_p, C.POINTER(_vp)]),
    "gpw_circuit_free": (None, [_vp]),
    "gpw_circuit_save": (C.c_int, [_vp, C.c_char_p]),
    "gpw_circuit_load": (C.c_int, [_vp, C.c_char_p, C.POINTER(_vp)]),
    "gpw_circuit_info": (C.c_int, [_vp, _vp]),
    "gpw_circuit_parse_inputs": (C.c_int, [_vp, C.c_char_p, C.c_char_p, _vp, C.c_size_t]),
    "gpw_witness_solve_phase1_dev": (C.c_int, [_vp, C.c_uint64, C.c_int, C.c_uint64, C.c_size_t]),
    "gpw_witness_solve_phase2_dev": (C.c_int, [_vp, _vp, C.c_int, C.c_uint64, C.c_size_t]),
    "gpw_r1cs_eval_dev": (C.c_int, [_vp, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, C.POINTER(C.c_uint64)]),
    "gpw_circuit_supports": (C.c_int, [_vp, C.c_int, _vp, C.c_size_t, C.POINTER(C.c_size_t)]),
    "gpw_circuit_hint_wires": (C.c_int, [_vp, C.c_int, _vp, C.c_size_t, C.POINTER(C.c_size_t)]),
    "gpw_wrap_key_synthetic": (C.c_int, [_vp, _vp, C.c_uint64, C.POINTER(_vp)]),
    "gpw_wrap_key_free": (None, [_vp]),
    "gpw_wrap_key_setup": (C.c_int, [_vp, _vp, C.c_char_p, C.POINTER(_vp)]),
    "gpw_wrap_key_save": (C.c_int, [_vp, C.c_char_p, C.c_char_p]),
    "gpw_wrap_key_load": (C.c_int, [_vp, _vp, C.c_char_p, C.c_char_p, C.POINTER(_vp)]),
    "gpw_wrap_key_vk_write_raw": (C.c_int, [_vp, _vp, C.c_size_t, C.POINTER(C.c_size_t)]),
    "gpw_wrap_proof_write_raw": (C.c_int, [_vp, _vp, _vp, C.c_size_t, C.POINTER(C.c_size_t)]),
    "gpw_wrap_key_info": (C.c_int, [_vp, _vp]),
    "gpw_wrap_key_wires_dev": (C.c_uint64, [_vp]),
    "gpw_wrap_key_h_dev": (C.c_uint64, [_vp]),
    "gpw_wrap_prove": (C.c_int, [_vp, _vp, _vp, _vp, C.c_int, _vp]),
    "gpw_wrap_prove_dev": (C.c_int, [_vp, C.c_uint64, _vp, _vp, C.c_int, _vp]),
    "gpw_msm_cumulative_stats": (C.c_int, [_vp, C.c_int, C.c_int, _vp]),
    "gpw_wrap_prove_many": (C.c_int, [_vp, _vp, C.c_int, _vp, _vp, C.c_int, _vp]),
    "gpw_witness_solve_phase1_on": (C.c_int, [_vp, _vp, C.c_uint64, C.c_int, C.c_uint64, C.c_size_t]),
    "gpw_witness_solve_phase2_on": (C.c_int, [_vp, _vp, _vp, C.c_int, C.c_uint64, C.c_size_t]),
    "gpw_r1cs_eval_on": (C.c_int, [_vp, _vp, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, _vp]),
    "gpw_wrap_set_lanes": (C.c_int, [_vp, C.c_int]),
    "gpw_ntt_share_tables": (C.c_int, [_vp, _vp, C.c_int]),
    "gpw_wrap_last_stats": (C.c_int, [_vp, _vp]),
    "gpw_hash_to_fr": (C.c_int, [C.c_char_p, C.c_size_t, C.c_char_p, _vp]),
    "gpw_plonk_setup": (C.c_int, [_vp, _vp, C.c_char_p, C.POINTER(_vp)]),
    "gpw_plonk_key_free": (None, [_vp]),
    "gpw_plonk_key_info": (C.c_int, [_vp, _vp]),
    "gpw_plonk_vk_write": (C.c_int, [_vp, _vp, C.c_size_t, C.POINTER(C.c_size_t)]),
    "gpw_plonk_prove": (C.c_int, [_vp, _vp, _vp, C.c_size_t]),
    "gpw_plonk_last_stats": (C.c_int, [_vp, _vp]),
}
for _name, (_res, _args) in _NEW.items():
    _fn = getattr(_lib, _name)
    _fn.restype = _res
    _fn.argtypes = _args
SYMBOLS.update(_NEW)

INFO_NAMES = ("wires public secret constraints instructions levels limb_wires limb_start count_start commit_wire "
              "narrow_segments wide_segments muladd reduce inverse split").split()
OP_MULADD, OP_REDUCE, OP_INVERSE, OP_SPLIT = 1, 2, 3, 4


def hash_to_fr(msg: bytes, dst: bytes = b"bsb22-commitment"):
    out = np.zeros(4, dtype=np.uint64)
    _check(_lib.gpw_hash_to_fr(msg, len(msg), dst, _p(out)))
    return sum(int(out[i]) << (64 * i) for i in range(4))


class Circuit:
    def __init__(self, ctx, handle):
        self.ctx, self._h = ctx, handle
        a = np.zeros(16, dtype=np.uint64)
        _check(_lib.gpw_circuit_info(self._h, _p(a)))
        self.info = dict(zip(INFO_NAMES, map(int, a)))
        self.n_inputs = self.info["public"] + self.info["secret"]
        self.n_wires = self.info["wires"]

    @classmethod
    def compile_verifier(cls, ctx, common_circuit_data_json, verifier_only_json=None, proof_json=None):
        """verifier_only_json: bake VerifierOnlyCircuitData in as constants (the reference's form; binds the statement to
        one inner circuit). proof_json: bake the proof in too (benchmark.go's ExampleVerifierCircuit)."""
        enc = lambda x: x.encode() if isinstance(x, str) else x
        h = _vp()
        _check(_lib.gpw_circuit_compile_verifier_bound(ctx._h, enc(common_circuit_data_json), enc(verifier_only_json),
                                                       enc(proof_json), C.byref(h)))
        return cls(ctx, h)

    def save(self, path):
        """compile cache (the r1cs.WriteTo of benchmark.go:94-99)"""
        _check(_lib.gpw_circuit_save(self._h, path.encode()))

    @classmethod
    def load(cls, ctx, path):
        h = _vp()
        _check(_lib.gpw_circuit_load(ctx._h, path.encode(), C.byref(h)))
        return cls(ctx, h)

    @classmethod
    def compile_gadget(cls, ctx, name):
        h = _vp()
        _check(_lib.gpw_circuit_compile_gadget(ctx._h, name.encode(), C.byref(h)))
        return cls(ctx, h)

    def close(self):
        if self._h:
            _lib.gpw_circuit_free(self._h)
            self._h = None

    def parse_inputs(self, proof_json, verifier_only_json):
        if isinstance(proof_json, str):
            proof_json = proof_json.encode()
        if isinstance(verifier_only_json, str):
            verifier_only_json = verifier_only_json.encode()
        out = np.zeros((self.n_inputs, 4), dtype=np.uint64)
        _check(_lib.gpw_circuit_parse_inputs(self._h, proof_json, verifier_only_json, _p(out), out.size))
        return out

    def inputs_from_ints(self, public, secret):
        assert len(public) == self.info["public"] and len(secret) == self.info["secret"]
        return ints_to_limbs(list(public) + list(secret))

    def solve_phase1_dev(self, inputs_ptr, n_proofs, wires_ptr, wire_stride):
        _check(_lib.gpw_witness_solve_phase1_dev(self._h, inputs_ptr, n_proofs, wires_ptr, wire_stride))

    def solve_phase2_dev(self, challenges, n_proofs, wires_ptr, wire_stride):
        ch = ints_to_limbs(list(challenges)) if challenges is not None else None
        _check(_lib.gpw_witness_solve_phase2_dev(self._h, _p(ch), n_proofs, wires_ptr, wire_stride))

    def r1cs_eval_dev(self, wires_ptr, a_ptr=0, b_ptr=0, c_ptr=0):
        """-> number of unsatisfied constraints (raises GpwError(-6) if any)"""
        n = C.c_uint64()
        _check(_lib.gpw_r1cs_eval_dev(self._h, wires_ptr, a_ptr, b_ptr, c_ptr, C.byref(n)))
        return n.value

    def hint_wires(self, op):
        n = C.c_size_t()
        _check(_lib.gpw_circuit_hint_wires(self._h, op, None, 0, C.byref(n)))
        out = np.zeros(n.value, dtype=np.uint32)
        _check(_lib.gpw_circuit_hint_wires(self._h, op, _p(out), out.size, C.byref(n)))
        return out

    def supports(self, side):
        out = np.zeros(self.n_wires, dtype=np.uint32)
        n = C.c_size_t()
        _check(_lib.gpw_circuit_supports(self._h, side, _p(out), out.size, C.byref(n)))
        return out[:n.value].copy()


class WrapKey:
    KEY_INFO = "m n_pub n_cons logN nA nB n_committed limb_start".split()

    def __init__(self, ctx, circuit, seed=0x5EED, _handle=None):
        """groth16.DummySetup analogue (synthetic bases with known discrete logs). Real keys: WrapKey.setup / WrapKey.load."""
        h = _handle
        if h is None:
            h = _vp()
            _check(_lib.gpw_wrap_key_synthetic(ctx._h, circuit._h, seed, C.byref(h)))
        self._h, self.ctx, self.circuit, self.seed = h, ctx, circuit, seed
        a = np.zeros(8, dtype=np.uint64)
        _check(_lib.gpw_wrap_key_info(self._h, _p(a)))
        self.info = dict(zip(self.KEY_INFO, map(int, a)))

    @classmethod
    def setup(cls, ctx, circuit, seed32=None):
        """groth16.Setup (benchmark.go:217). seed32: 32 bytes (reproducible toxic waste, tests) or None (OS entropy)."""
        assert seed32 is None or len(seed32) == 32
        h = _vp()
        _check(_lib.gpw_wrap_key_setup(ctx._h, circuit._h, seed32, C.byref(h)))
        return cls(ctx, circuit, None, _handle=h)

    @classmethod
    def load(cls, ctx, circuit, pk_path, vk_path=None):
        h = _vp()
        _check(_lib.gpw_wrap_key_load(ctx._h, circuit._h, pk_path.encode(), vk_path.encode() if vk_path else None, C.byref(h)))
        return cls(ctx, circuit, None, _handle=h)

    def save(self, pk_path, vk_path=None):
        _check(_lib.gpw_wrap_key_save(self._h, pk_path.encode(), vk_path.encode() if vk_path else None))

    def vk_raw(self) -> bytes:
        """vk.WriteRawTo bytes"""
        n = C.c_size_t()
        _check(_lib.gpw_wrap_key_vk_write_raw(self._h, None, 0, C.byref(n)))
        buf = np.zeros(n.value, dtype=np.uint8)
        _check(_lib.gpw_wrap_key_vk_write_raw(self._h, _p(buf), buf.size, C.byref(n)))
        return buf.tobytes()

    def proof_raw(self, proof) -> bytes:
        """proof.WriteRawTo bytes of a proof dict returned by prove*()"""
        raw = np.ascontiguousarray(proof["raw"] if isinstance(proof, dict) else proof, dtype=np.uint64)
        n = C.c_size_t()
        _check(_lib.gpw_wrap_proof_write_raw(self._h, _p(raw), None, 0, C.byref(n)))
        buf = np.zeros(n.value, dtype=np.uint8)
        _check(_lib.gpw_wrap_proof_write_raw(self._h, _p(raw), _p(buf), buf.size, C.byref(n)))
        return buf.tobytes()

    def close(self):
        if self._h:
            _lib.gpw_wrap_key_free(self._h)
            self._h = None

    @property
    def wires_ptr(self):
        return int(_lib.gpw_wrap_key_wires_dev(self._h))

    @property
    def h_ptr(self):
        """device address of the quotient coefficients h (N - 1 Fr, Montgomery) of the last prove()"""
        return int(_lib.gpw_wrap_key_h_dev(self._h))

    def prove(self, inputs, r_int=None, s_int=None, check=True):
        """inputs: (n_inputs, 4) u64 canonical, host. -> dict with Ar, Bs, Krs, commitment, pok (affine Montgomery limbs),
        challenge (int), n_unsatisfied"""
        inputs = np.ascontiguousarray(inputs, dtype=np.uint64)
        assert inputs.shape == (self.circuit.n_inputs, 4)
        r = ints_to_limbs([r_int])[0] if r_int is not None else None     # None: sampled inside libgpw (CSPRNG)
        s = ints_to_limbs([s_int])[0] if s_int is not None else None
        out = np.zeros(64, dtype=np.uint64)
        _check(_lib.gpw_wrap_prove(self._h, _p(inputs), _p(r), _p(s), int(check), _p(out)))
        return {"Ar": out[0:8].copy(), "Bs": out[8:24].copy(), "Krs": out[24:32].copy(), "commitment": out[32:40].copy(),
                "pok": out[40:48].copy(), "challenge": sum(int(out[48 + i]) << (64 * i) for i in range(4)),
                "n_unsatisfied": int(out[52]), "raw": out}

    def _unpack(self, out):
        return {"Ar": out[0:8].copy(), "Bs": out[8:24].copy(), "Krs": out[24:32].copy(), "commitment": out[32:40].copy(),
                "pok": out[40:48].copy(), "challenge": sum(int(out[48 + i]) << (64 * i) for i in range(4)),
                "n_unsatisfied": int(out[52]), "raw": out}

    def prove_ptr(self, inputs_ptr, r_int, s_int, check=True, on_device=False):
        """inputs_ptr: address of the (n_inputs, 4) u64 canonical input block - pinned host memory, or device memory
        when on_device=True."""
        r = ints_to_limbs([r_int])[0]
        s = ints_to_limbs([s_int])[0]
        out = np.zeros(64, dtype=np.uint64)
        if on_device:
            _check(_lib.gpw_wrap_prove_dev(self._h, inputs_ptr, _p(r), _p(s), int(check), _p(out)))
        else:
            _check(_lib.gpw_wrap_prove(self._h, _vp(inputs_ptr), _p(r), _p(s), int(check), _p(out)))
        return self._unpack(out)

    def set_lanes(self, n):
        """proofs gpw_wrap_prove_many keeps in flight (one host thread + stream + scratch each)"""
        _check(_lib.gpw_wrap_set_lanes(self._h, int(n)))

    def prove_many(self, inputs_ptr, n, r_ints=None, s_ints=None, check=True):
        """n independent proofs, several in flight (gpw_wrap_prove_many). inputs_ptr: host address of n x n_inputs x 4 u64."""
        r = ints_to_limbs(list(r_ints)) if r_ints is not None else None
        s = ints_to_limbs(list(s_ints)) if s_ints is not None else None
        out = np.zeros((n, 64), dtype=np.uint64)
        _check(_lib.gpw_wrap_prove_many(self._h, _vp(inputs_ptr), n, _p(r), _p(s), int(check), _p(out)))
        return [self._unpack(out[i]) for i in range(n)]

    def msm_cumulative_stats(self, group, reset=False):
        out = np.zeros(5, dtype=np.float64)
        _check(_lib.gpw_msm_cumulative_stats(self.ctx._h, group, int(reset), _p(out)))
        return dict(zip(("accumulate_ms", "total_ms", "points", "digits", "calls"), map(float, out)))

    def last_stats(self):
        ms = (C.c_float * 6)()
        _check(_lib.gpw_wrap_last_stats(self._h, ms))
        names = ("solve_phase1_ms", "commitment_ms", "solve_phase2_ms", "r1cs_eval_ms", "compute_h_ms", "msm_ms")
        return dict(zip(names, [float(x) for x in ms]))


class PlonkKey:
    """plonk.Setup / plonk.Prove (benchmark.go:130, 162) over the library's lowering of a compiled circuit (csrc/plonk.cu)."""
    PROOF_BYTES = 10 * 64 + 18 * 32
    INFO = "logN gates variables public_rows qcp_rows inputs has_commit chain_levels".split()

    def __init__(self, ctx, circuit, seed32=None):
        assert seed32 is None or len(seed32) == 32
        h = _vp()
        _check(_lib.gpw_plonk_setup(ctx._h, circuit._h, seed32, C.byref(h)))
        self._h, self.ctx, self.circuit = h, ctx, circuit
        a = np.zeros(8, dtype=np.uint64)
        _check(_lib.gpw_plonk_key_info(self._h, _p(a)))
        self.info = dict(zip(self.INFO, map(int, a)))

    def close(self):
        if self._h:
            _lib.gpw_plonk_key_free(self._h)
            self._h = None

    def vk(self) -> bytes:
        n = C.c_size_t()
        _check(_lib.gpw_plonk_vk_write(self._h, None, 0, C.byref(n)))
        buf = np.zeros(n.value, dtype=np.uint8)
        _check(_lib.gpw_plonk_vk_write(self._h, _p(buf), buf.size, C.byref(n)))
        return buf.tobytes()

    def prove(self, inputs) -> bytes:
        inputs = np.ascontiguousarray(inputs, dtype=np.uint64)
        assert inputs.shape == (self.circuit.n_inputs, 4)
        out = np.zeros(self.PROOF_BYTES, dtype=np.uint8)
        _check(_lib.gpw_plonk_prove(self._h, _p(inputs), _p(out), out.size))
        return out.tobytes()

    def last_stats(self):
        ms = (C.c_float * 6)()
        _check(_lib.gpw_plonk_last_stats(self._h, ms))
        return dict(zip(("witness_ms", "wires_ms", "grand_product_ms", "quotient_ms", "evaluations_ms", "openings_ms"), map(float, ms)))
