"""The C++ gadget library + frontend (the product's host side) validated on CPU through the TEST-ONLY sequential
tape interpreter (tests/hostlib): every constraint of the compiled verifier circuit is satisfied on the real
testdata/step proof, and the ordered outputs of all 449k reference hints equal the oracle's trace."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import gpw
from oracle.engine import Api
from oracle.poseidon import GoldilocksChip, BN254Chip
from oracle import goldilocks as ogl
from oracle.verifier import verify_testdata

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOSTLIB = os.path.join(ROOT, "tests", "hostlib")
NAMES = ("wires public secret constraints tape levels commit_level limb_wires muladd reduce glinv split invzero bits "
         "div decomp mul coeffs le_terms limb_start count_start commit_wire").split()
OPS = {"muladd": 1, "reduce": 2, "inverse": 3, "split": 4}


@pytest.fixture(scope="module")
def lib():
    subprocess.check_call(["make", "-C", HOSTLIB], stdout=subprocess.DEVNULL)
    lib = C.CDLL(os.path.join(HOSTLIB, "libgpw_circuit_test.so"))
    lib.ct_compile.restype = C.c_void_p
    lib.ct_compile.argtypes = [C.c_char_p]
    lib.ct_compile_small.restype = C.c_void_p
    lib.ct_compile_small.argtypes = [C.c_int]
    lib.ct_free.argtypes = [C.c_void_p]
    lib.ct_last_error.restype = C.c_char_p
    lib.ct_stats.argtypes = [C.c_void_p, C.c_void_p]
    lib.ct_solve_inputs.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p]
    lib.ct_solve_testdata.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_void_p]
    lib.ct_check.restype = C.c_uint64
    lib.ct_check.argtypes = [C.c_void_p, C.c_void_p]
    lib.ct_hint_outputs.restype = C.c_size_t
    lib.ct_hint_outputs.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t]
    return lib


def stats(lib, h):
    a = np.zeros(len(NAMES), dtype=np.uint64)
    lib.ct_stats(h, a.ctypes.data)
    return dict(zip(NAMES, map(int, a)))


def solve(lib, h, pub, sec, x=123456789):
    pl, sl, xl = gpw.ints_to_limbs(pub), gpw.ints_to_limbs(sec), gpw.ints_to_limbs([x])
    rc = lib.ct_solve_inputs(h, pl.ctypes.data, len(pub), sl.ctypes.data, len(sec), xl.ctypes.data)
    fb = C.c_int64()
    bad = lib.ct_check(h, C.byref(fb)) if rc == 0 else None
    return rc, bad


def test_small_circuits_satisfied_and_reject_wrong_outputs(lib, kats):
    # Poseidon-GL KAT (poseidon/goldilocks_test.go:37-59): 130 MulAdd + 630 Reduce + 890 SplitLimbs per permutation
    h = lib.ct_compile_small(0)
    st = stats(lib, h)
    assert (st["muladd"], st["reduce"], st["split"]) == (130, 630, 890)
    out = [int(x) for x in kats["poseidon_gl_perm_zero"]]
    assert solve(lib, h, out, [0] * 12) == (0, 0)
    rc, bad = solve(lib, h, [out[0] ^ 1] + out[1:], [0] * 12)
    assert rc == 0 and bad == 1
    lib.ct_free(h)
    # Poseidon-BN254 KATs (poseidon/bn254_test.go:31-97)
    h = lib.ct_compile_small(1)
    for case in kats["poseidon_bn254"]:
        assert solve(lib, h, [int(x) for x in case["out"]], [int(x) for x in case["in"]]) == (0, 0)
    lib.ct_free(h)
    # QE mul / div (goldilocks/quadratic_extension_test.go)
    h = lib.ct_compile_small(2)
    a = tuple(map(int, kats["qe_mul"]["a"]))
    b = tuple(map(int, kats["qe_mul"]["b"]))
    c = ogl.Chip(Api(trace=False))
    m = c.MulExtension(a, b)
    d, _ = c.DivExtension(a, b)
    assert m == tuple(map(int, kats["qe_mul"]["out"]))
    assert solve(lib, h, list(m) + list(d), list(a) + list(b)) == (0, 0)
    lib.ct_free(h)
    # RangeCheck accepts 0, 1, p-1 and the hint rejects p (goldilocks/base_test.go:26-44)
    h = lib.ct_compile_small(3)
    for x in (0, 1, ogl.P - 1):
        assert solve(lib, h, [], [x]) == (0, 0)
    rc, _ = solve(lib, h, [], [ogl.P])
    assert rc == -5 and b"SplitLimbsHint" in lib.ct_last_error()
    lib.ct_free(h)


def test_poseidon_gl_macro_host(lib, kats):
    # the native evaluation behind the OP_POSEIDON_GL macro instruction (csrc/poseidon_gl_macro.cuh), host build:
    # the specialised ReduceHint equals the general one; a permutation fills each of its 1992 output slots exactly once
    # and ends in the reference's state (poseidon/goldilocks_test.go:47-53 and the oracle on random states)
    import random
    from oracle.engine import Api
    from oracle.poseidon import GoldilocksChip
    lib.ct_glm_reduce192_mismatches.restype = C.c_uint64
    lib.ct_glm_reduce192_mismatches.argtypes = [C.c_uint64, C.c_uint64]
    lib.ct_glm_permute.restype = C.c_uint32
    lib.ct_glm_permute.argtypes = [C.c_void_p, C.c_void_p]
    assert lib.ct_glm_reduce192_mismatches(7, 2_000_000) == 0
    rng = random.Random(3)
    P = (1 << 64) - (1 << 32) + 1
    for state in ([0] * 12, [P - 1] * 12, [rng.randrange(P) for _ in range(12)]):
        a = np.array(state, dtype=np.uint64)
        out = np.zeros(12, dtype=np.uint64)
        assert lib.ct_glm_permute(a.ctypes.data, out.ctypes.data) == 1992
        exp = GoldilocksChip(Api(trace=False)).Poseidon(list(state))
        assert [int(x) for x in out] == [int(x) for x in exp]
    out = np.zeros(12, dtype=np.uint64)
    lib.ct_glm_permute(np.zeros(12, dtype=np.uint64).ctypes.data, out.ctypes.data)
    assert [int(x) for x in out] == [int(x) for x in kats["poseidon_gl_perm_zero"]]


def _schedule_check(lib, h):
    lib.ct_schedule_check.argtypes = [C.c_void_p, C.c_void_p]
    out = np.zeros(3, dtype=np.uint64)
    lib.ct_schedule_check(h, out.ctypes.data)
    return [int(x) for x in out]


def test_tape_schedule_is_consistent(lib, testdata_dir):
    # the tape the GPU executor walks: after the ASAP levels of the builder, after ALAP and after the spine-and-tail
    # schedule every operand comes from a strictly lower level and no wire has two producers; the verifier circuit of
    # `step` carries one Poseidon-Goldilocks macro per permutation (129 challenger + 5 public-input hash) and one
    # Poseidon-BN254 macro per Merkle / leaf hash permutation
    lib.ct_schedule_spine_tail.argtypes = [C.c_void_p]
    h = lib.ct_compile_small(0)
    assert _schedule_check(lib, h)[:2] == [0, 0]
    lib.ct_schedule_spine_tail(h)
    assert _schedule_check(lib, h)[:2] == [0, 0]
    lib.ct_free(h)
    h = lib.ct_compile(open(os.path.join(testdata_dir, "step", "common_circuit_data.json"), "rb").read())
    bad, dup, macros = _schedule_check(lib, h)
    assert (bad, dup) == (0, 0) and macros == 134
    lib.ct_schedule_spine_tail(h)
    bad, dup, macros = _schedule_check(lib, h)
    assert (bad, dup) == (0, 0) and macros == 134
    lib.ct_free(h)


def test_compile_rejects_unsupported_circuits(lib, testdata_dir):
    # types/common_data.go:121-124 panics on hiding = true; an unknown gate id is refused by GateInstanceFromId
    # (plonk/gates/gates.go:37-54); both must fail loudly here too, in the C++ frontend and in the oracle
    import json
    from oracle.types import CommonCircuitData
    src = open(os.path.join(testdata_dir, "step", "common_circuit_data.json")).read()
    d = json.loads(src)
    d["fri_params"]["hiding"] = True
    assert not lib.ct_compile(json.dumps(d).encode())
    assert b"hiding" in lib.ct_last_error()
    with pytest.raises(ValueError):
        CommonCircuitData(d)
    d = json.loads(src)
    d["gates"][0] = "NotAGate { num_things: 3 }"
    assert not lib.ct_compile(json.dumps(d).encode())
    assert lib.ct_last_error()


def test_full_verifier_circuit_on_step(lib, testdata_dir):
    d = os.path.join(testdata_dir, "step")
    rd = lambda f: open(os.path.join(d, f), "rb").read()
    h = lib.ct_compile(rd("common_circuit_data.json"))
    assert h, lib.ct_last_error()
    st = stats(lib, h)
    # same hint counts as the oracle's run of the reference dataflow
    assert (st["muladd"], st["reduce"], st["split"], st["glinv"]) == (44186, 151410, 251232, 1849)
    assert st["limb_wires"] == 2462493 and st["public"] == 36
    assert 5_000_000 < st["constraints"] < 6_500_000
    x = gpw.ints_to_limbs([0x1234567890abcdef1234567890abcdef])
    rc = lib.ct_solve_testdata(h, rd("proof_with_public_inputs.json"), rd("verifier_only_circuit_data.json"), x.ctypes.data)
    assert rc == 0, lib.ct_last_error()
    fb = C.c_int64()
    assert lib.ct_check(h, C.byref(fb)) == 0, "first unsatisfied constraint: %d" % fb.value
    # ordered hint outputs == oracle trace
    api, _ = verify_testdata(d)
    for kind, op in OPS.items():
        exp = [o for k, _, outs in api.hints if k == kind for o in outs]
        buf = np.zeros((len(exp), 4), dtype=np.uint64)
        n = lib.ct_hint_outputs(h, op, buf.ctypes.data, buf.size)
        assert n == buf.size
        assert gpw.limbs_to_ints(buf) == exp, kind
    # a tampered proof must leave constraints unsatisfied
    import json
    p = json.loads(rd("proof_with_public_inputs.json"))
    p["proof"]["openings"]["plonk_zs"][0][0] ^= 1
    rc = lib.ct_solve_testdata(h, json.dumps(p).encode(), rd("verifier_only_circuit_data.json"), x.ctypes.data)
    assert rc == 0 and lib.ct_check(h, C.byref(fb)) > 0
    lib.ct_free(h)


def test_baked_circuit_form_on_decode_block(lib, testdata_dir):
    # verifier/util.go:10-24 as benchmark.go compiles it: proof + verifier-only data are compile-time constants. Same hint
    # counts as the runtime-input form, no secret inputs left, every constraint satisfied. (Form 1 - only the verifier data
    # baked - is exercised on the GPU, tests/test_gpu_setup_verify.py.)
    d = os.path.join(testdata_dir, "decode_block")
    rd = lambda f: open(os.path.join(d, f), "rb").read()
    lib.ct_compile_baked.restype = C.c_void_p
    lib.ct_compile_baked.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_int]
    h = lib.ct_compile_baked(rd("common_circuit_data.json"), rd("proof_with_public_inputs.json"), rd("verifier_only_circuit_data.json"), 2)
    assert h, lib.ct_last_error()
    st = stats(lib, h)
    assert st["secret"] == 0 and st["muladd"] > 40000 and st["reduce"] > 140000
    x = gpw.ints_to_limbs([0x1234567890abcdef1234567890abcdef])
    rc = lib.ct_solve_testdata(h, rd("proof_with_public_inputs.json"), rd("verifier_only_circuit_data.json"), x.ctypes.data)
    assert rc == 0, lib.ct_last_error()
    fb = C.c_int64()
    assert lib.ct_check(h, C.byref(fb)) == 0, "first unsatisfied constraint: %d" % fb.value
    lib.ct_free(h)


def test_compile_cache_round_trip(lib, kats, tmp_path):
    # fe::API::Serialize / Deserialize (the compile cache behind gpw_circuit_save / _load; benchmark.go:94-99 is the
    # r1cs.WriteTo the reference had to comment out): a reloaded circuit has the same shape, solves and is satisfied; a
    # truncated or foreign file is refused. Scheduled and unscheduled tapes both survive.
    lib.ct_save.argtypes = [C.c_void_p, C.c_char_p]
    lib.ct_load.restype = C.c_void_p
    lib.ct_load.argtypes = [C.c_char_p]
    lib.ct_schedule_spine_tail.argtypes = [C.c_void_p]
    out = [int(x) for x in kats["poseidon_gl_perm_zero"]]
    for schedule in (False, True):
        h = lib.ct_compile_small(0)
        if schedule:
            lib.ct_schedule_spine_tail(h)
        path = str(tmp_path / ("circ%d.bin" % schedule)).encode()
        assert lib.ct_save(h, path) == 0
        h2 = lib.ct_load(path)
        assert h2, lib.ct_last_error()
        assert stats(lib, h2) == stats(lib, h)
        assert _schedule_check(lib, h2) == _schedule_check(lib, h)
        assert solve(lib, h2, out, [0] * 12) == (0, 0)
        rc, bad = solve(lib, h2, [out[0] ^ 1] + out[1:], [0] * 12)
        assert rc == 0 and bad == 1
        lib.ct_free(h)
        lib.ct_free(h2)
    blob = open(path, "rb").read()
    open(path, "wb").write(blob[:len(blob) // 2])
    assert not lib.ct_load(path) and b"circuit cache" in lib.ct_last_error()
    open(path, "wb").write(b"\0" * 64 + blob[64:])
    assert not lib.ct_load(path)


def test_plonk_lowering_of_small_circuits(lib, kats):
    # csrc/host/scs.cc (the scs.NewBuilder counterpart, benchmark.go:44-45): the compiled circuit lowered to PLONK gates. On the
    # solved witness every gate equation holds (addition chains with the constant on qC, multiplication rows, public rows, Qcp
    # rows for the committed range-check wires), sigma is a permutation of the 3 N slots that stays inside each variable's slots,
    # and a wrong witness breaks gates. (The full verifier circuits are checked the same way by hand - 33.46 M gates for step,
    # 2^25 rows - and end to end on the GPU, tests/test_gpu_plonk.py.)
    lib.ct_scs_check.argtypes = [C.c_void_p, C.c_void_p]
    out = [int(x) for x in kats["poseidon_gl_perm_zero"]]
    h = lib.ct_compile_small(0)
    assert solve(lib, h, out, [0] * 12) == (0, 0)
    o = np.zeros(6, dtype=np.uint64)
    lib.ct_scs_check(h, o.ctypes.data)
    gates, nvars, logn, bad, first_bad, perm_bad = map(int, o)
    st = stats(lib, h)
    assert bad == 0 and perm_bad == 0
    assert (1 << (logn - 1)) < gates <= (1 << logn) and nvars > st["wires"]
    assert gates >= st["constraints"] + st["limb_wires"] + 65536            # one row per R1CS row and per committed wire at least
    rc, r1cs_bad = solve(lib, h, [out[0] ^ 1] + out[1:], [0] * 12)
    assert rc == 0 and r1cs_bad == 1
    lib.ct_scs_check(h, o.ctypes.data)
    assert int(o[3]) >= 1 and int(o[5]) == 0
    lib.ct_free(h)
    h = lib.ct_compile_small(1)                                             # Poseidon-BN254: no range checks, no Qcp rows
    case = kats["poseidon_bn254"][0]
    assert solve(lib, h, [int(x) for x in case["out"]], [int(x) for x in case["in"]]) == (0, 0)
    lib.ct_scs_check(h, o.ctypes.data)
    assert int(o[3]) == 0 and int(o[5]) == 0
    lib.ct_free(h)
