"""GPU witness synthesis + wrap proving, through the C ABI, against the oracle and the test-only host interpreter."""
import ctypes as C
import os
import random
import subprocess

import numpy as np
import pytest

import gpw
from oracle import bn254 as ob
from oracle import goldilocks as ogl
from oracle.engine import Api
from oracle.poseidon import GoldilocksChip, BN254Chip
from oracle.verifier import verify_testdata

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
R = ob.R


@pytest.fixture(scope="module")
def ctx():
    c = gpw.Context(0)
    yield c
    c.close()


def _solve(ctx, circ, inputs, challenge=0xabcdef0123456789abcdef0123456789):
    import torch
    inp = torch.from_numpy(np.ascontiguousarray(inputs).view(np.int64)).cuda()
    wires = torch.zeros((circ.n_wires, 4), dtype=torch.int64, device="cuda")
    torch.cuda.synchronize()
    circ.solve_phase1_dev(inp.data_ptr(), 1, wires.data_ptr(), circ.n_wires)
    circ.solve_phase2_dev([challenge] if circ.info["limb_wires"] else None, 1, wires.data_ptr(), circ.n_wires)
    ctx.sync()
    return wires


def _wire_ints(wires, idx):
    import torch
    sel = wires[torch.from_numpy(idx.astype(np.int64)).cuda()].cpu().numpy().view(np.uint64)
    return gpw.limbs_to_ints(gpw.host_ff_from_mont(0, sel))


def test_gadget_circuits_on_gpu(ctx, kats):
    # Poseidon-GL KAT circuit: satisfied with the right outputs, unsatisfied (-6) with a wrong one
    circ = gpw.Circuit.compile_gadget(ctx, "poseidon_gl")
    assert (circ.info["muladd"], circ.info["reduce"], circ.info["split"]) == (130, 630, 890)
    out = [int(x) for x in kats["poseidon_gl_perm_zero"]]
    w = _solve(ctx, circ, circ.inputs_from_ints(out, [0] * 12))
    assert circ.r1cs_eval_dev(w.data_ptr()) == 0
    # the hint outputs on the GPU == the oracle's hint trace, in order
    api = Api()
    GoldilocksChip(api).Poseidon([0] * 12)
    for kind, op in (("muladd", 1), ("reduce", 2), ("split", 4)):
        exp = [o for k, _, outs in api.hints if k == kind for o in outs]
        assert _wire_ints(w, circ.hint_wires(op)) == exp
    w = _solve(ctx, circ, circ.inputs_from_ints([out[0] ^ 1] + out[1:], [0] * 12))
    with pytest.raises(gpw.GpwError) as e:
        circ.r1cs_eval_dev(w.data_ptr())
    assert e.value.code == -6
    circ.close()
    circ = gpw.Circuit.compile_gadget(ctx, "poseidon_bn254")
    for case in kats["poseidon_bn254"]:
        w = _solve(ctx, circ, circ.inputs_from_ints([int(x) for x in case["out"]], [int(x) for x in case["in"]]))
        assert circ.r1cs_eval_dev(w.data_ptr()) == 0
    circ.close()
    circ = gpw.Circuit.compile_gadget(ctx, "qe_mul_div")
    a = tuple(map(int, kats["qe_mul"]["a"]))
    b = tuple(map(int, kats["qe_mul"]["b"]))
    ch = ogl.Chip(Api(trace=False))
    m = ch.MulExtension(a, b)
    d, _ = ch.DivExtension(a, b)
    w = _solve(ctx, circ, circ.inputs_from_ints(list(m) + list(d), list(a) + list(b)))
    assert circ.r1cs_eval_dev(w.data_ptr()) == 0
    circ.close()
    # RangeCheck: accepts 0, 1, p-1; the SplitLimbs hint rejects p (goldilocks/base_test.go:26-44)
    circ = gpw.Circuit.compile_gadget(ctx, "range_check")
    for x in (0, 1, ogl.P - 1):
        w = _solve(ctx, circ, circ.inputs_from_ints([], [x]))
        assert circ.r1cs_eval_dev(w.data_ptr()) == 0
    with pytest.raises(gpw.GpwError) as e:
        _solve(ctx, circ, circ.inputs_from_ints([], [ogl.P]))
    assert e.value.code == -5 and "SplitLimbsHint" in str(e.value)
    circ.close()


@pytest.fixture(scope="module")
def step(ctx, testdata_dir):
    d = os.path.join(testdata_dir, "step")
    rd = lambda f: open(os.path.join(d, f), "rb").read()
    circ = gpw.Circuit.compile_verifier(ctx, rd("common_circuit_data.json"))
    inputs = circ.parse_inputs(rd("proof_with_public_inputs.json"), rd("verifier_only_circuit_data.json"))
    yield circ, inputs, d
    circ.close()


def test_step_witness_on_gpu_matches_oracle_trace_and_satisfies_r1cs(ctx, step):
    circ, inputs, d = step
    assert (circ.info["muladd"], circ.info["reduce"], circ.info["split"], circ.info["inverse"]) == (44186, 151410, 251232, 1849)
    w = _solve(ctx, circ, inputs, challenge=0x1234567890abcdef1234567890abcdef)
    assert circ.r1cs_eval_dev(w.data_ptr()) == 0          # all ~5.6 M constraints hold on the device
    api, _ = verify_testdata(d)
    for kind, op in (("muladd", 1), ("reduce", 2), ("inverse", 3), ("split", 4)):
        exp = [o for k, _, outs in api.hints if k == kind for o in outs]
        assert _wire_ints(w, circ.hint_wires(op)) == exp, kind


def test_step_witness_equals_host_interpreter(ctx, step):
    # wire-for-wire equality of the CUDA executor with the sequential test interpreter
    circ, inputs, d = step
    hostlib = os.path.join(ROOT, "tests", "hostlib")
    subprocess.check_call(["make", "-C", hostlib], stdout=subprocess.DEVNULL)
    lib = C.CDLL(os.path.join(hostlib, "libgpw_circuit_test.so"))
    lib.ct_compile.restype = C.c_void_p
    lib.ct_compile.argtypes = [C.c_char_p]
    lib.ct_solve_inputs.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p]
    lib.ct_wires_mont.argtypes = [C.c_void_p, C.c_void_p]
    h = lib.ct_compile(open(os.path.join(d, "common_circuit_data.json"), "rb").read())
    x = 987654321987654321
    xl = gpw.ints_to_limbs([x])
    npub = circ.info["public"]
    pub, sec = np.ascontiguousarray(inputs[:npub]), np.ascontiguousarray(inputs[npub:])
    assert lib.ct_solve_inputs(h, pub.ctypes.data, npub, sec.ctypes.data, len(sec), xl.ctypes.data) == 0
    ref = np.zeros((circ.n_wires, 4), dtype=np.uint64)
    lib.ct_wires_mont(h, ref.ctypes.data)
    w = _solve(ctx, circ, inputs, challenge=x).cpu().numpy().view(np.uint64)
    assert (w == ref).all()


def test_tampered_proof_is_rejected_on_gpu(ctx, step):
    import json
    circ, inputs, d = step
    rd = lambda f: open(os.path.join(d, f), "rb").read()
    p = json.loads(rd("proof_with_public_inputs.json"))
    p["proof"]["openings"]["plonk_zs"][0][0] ^= 1
    bad = circ.parse_inputs(json.dumps(p), rd("verifier_only_circuit_data.json"))
    with pytest.raises(gpw.GpwError) as e:      # unsatisfiable: caught by a range check in the solve or by the R1CS check
        w = _solve(ctx, circ, bad)
        circ.r1cs_eval_dev(w.data_ptr())
    assert e.value.code == -6
    # a structurally wrong proof is refused by the codec
    del p["proof"]["openings"]["wires"][0]
    with pytest.raises(gpw.GpwError):
        circ.parse_inputs(json.dumps(p), rd("verifier_only_circuit_data.json"))


def test_wrap_prove_step(ctx, step):
    # full wrap: witness + commitment + Groth16; proof elements recomputed in the exponent from the witness
    import torch
    circ, inputs, d = step
    key = gpw.WrapKey(ctx, circ, seed=99)
    r_, s_ = 0x1111222233334444, 0x5555666677778888
    pr = key.prove(inputs, r_, s_, check=True)
    assert pr["n_unsatisfied"] == 0
    pr2 = key.prove(inputs, r_, s_, check=True)
    assert (pr["raw"] == pr2["raw"]).all()            # deterministic given (r, s)
    for g, name in ((1, "Ar"), (2, "Bs"), (1, "Krs"), (1, "commitment"), (1, "pok")):
        assert gpw.host_ec_is_on_curve(g, pr[name]) and pr[name].any()
    # challenge = hash_to_field(serialised commitment)
    cx, cy = gpw.points_to_ints(1, pr["commitment"])[0]
    assert pr["challenge"] == gpw.hash_to_fr(cx.to_bytes(32, "big") + cy.to_bytes(32, "big"))
    # exponent check of Ar and of the commitment from the device-resident witness
    info = key.info
    wires = torch.empty((info["m"], 4), dtype=torch.int64, device="cuda")
    import ctypes
    ctypes.cdll.LoadLibrary("libcudart.so").cudaMemcpy(ctypes.c_void_p(wires.data_ptr()), ctypes.c_void_p(key.wires_ptr),
                                                       ctypes.c_size_t(info["m"] * 32), ctypes.c_int(3))
    wv = gpw.limbs_to_ints(gpw.host_ff_from_mont(0, wires.cpu().numpy().view(np.uint64)))
    suppA = circ.supports(0)
    sA = (99 + 1 + sum(wv[w_] * (1 + j) for j, w_ in enumerate(suppA)) + r_ * (99 + 3)) % R
    assert gpw.points_to_ints(1, pr["Ar"])[0] == ob.point_key(1, ob.ec_mul(1, ob.G1_GEN, sA))
    ls, nc = info["limb_start"], info["n_committed"]
    sD = sum(wv[ls + i] * ((1 << 35) + ls + i) for i in range(nc)) % R
    assert gpw.points_to_ints(1, pr["commitment"])[0] == ob.point_key(1, ob.ec_mul(1, ob.G1_GEN, sD))
    # Bs (G2) and Krs in the exponent too. Discrete logs of the synthetic key (csrc/wrap.cu): B1_j = 2^32 + j,
    # B2_j = 1 + j over B's support, K_i = 2^33 + i, Z_j = 2^34 + j, alpha/beta/delta = seed + 1/2/3.
    suppB = circ.supports(1)
    sB2 = (99 + 2 + sum(wv[w_] * (1 + j) for j, w_ in enumerate(suppB)) + s_ * (99 + 3)) % R
    assert gpw.points_to_ints(2, pr["Bs"])[0] == ob.point_key(2, ob.ec_mul(2, ob.G2_GEN, sB2))
    sB1 = (99 + 2 + sum(wv[w_] * ((1 << 32) + j) for j, w_ in enumerate(suppB)) + s_ * (99 + 3)) % R
    N = 1 << info["logN"]
    hbuf = torch.empty((N - 1, 4), dtype=torch.int64, device="cuda")
    ctypes.cdll.LoadLibrary("libcudart.so").cudaMemcpy(ctypes.c_void_p(hbuf.data_ptr()), ctypes.c_void_p(key.h_ptr),
                                                       ctypes.c_size_t((N - 1) * 32), ctypes.c_int(3))
    hv = gpw.limbs_to_ints(gpw.host_ff_from_mont(0, hbuf.cpu().numpy().view(np.uint64)))
    sZ = sum(h * ((1 << 34) + j) for j, h in enumerate(hv)) % R
    # K: private wires that are neither committed nor the commitment challenge (a public input of the verifier)
    npub, m = info["n_pub"], info["m"]
    k_wires = list(range(1 + npub, ls)) + list(range(ls + nc + 1, m))
    assert circ.info["commit_wire"] == ls + nc
    sK = sum(wv[i] * ((1 << 33) + i) for i in k_wires) % R
    sKrs = (sK + sZ + s_ * sA + r_ * sB1 - r_ * s_ * (99 + 3)) % R
    assert gpw.points_to_ints(1, pr["Krs"])[0] == ob.point_key(1, ob.ec_mul(1, ob.G1_GEN, sKrs))
    print("wrap stats (ms):", key.last_stats(), "key:", info)
    key.close()


def test_prove_many_lanes_match_single_proofs(ctx, step):
    # a stream of proofs with several in flight (one host thread + stream + scratch per lane) must return, proof for
    # proof, exactly what the one-at-a-time entry point returns - with the inputs in host memory and in device memory
    import torch
    circ, inputs, d = step
    key = gpw.WrapKey(ctx, circ, seed=7)
    n = 5
    rs = [(0x1000 + 17 * i, 0x2000 + 31 * i) for i in range(n)]
    single = [key.prove(inputs, r_, s_, check=True)["raw"].copy() for r_, s_ in rs[:3]]
    many_host = np.ascontiguousarray(np.tile(inputs, (n, 1, 1)))
    key.set_lanes(3)
    out = key.prove_many(many_host.ctypes.data, n, [r for r, _ in rs], [s for _, s in rs], check=True)
    assert len(out) == n and all(p["n_unsatisfied"] == 0 for p in out)
    for i in range(3):
        assert (out[i]["raw"] == single[i]).all(), i
    assert len({p["raw"].tobytes() for p in out}) == n          # distinct (r, s) -> distinct proofs
    dev = torch.from_numpy(many_host.view(np.int64)).cuda()
    torch.cuda.synchronize()
    key.set_lanes(2)
    out2 = key.prove_many(dev.data_ptr(), n, [r for r, _ in rs], [s for _, s in rs], check=True)
    assert all((a["raw"] == b["raw"]).all() for a, b in zip(out, out2))
    # deferred MSMs (a lone proof overlaps the tail of each MSM with the next one, csrc/common.cuh) change no byte: forced off,
    # forced on for a stream of two lanes, and the one-lane stream (automatic: on) all reproduce the single proofs
    ctx.set_option("msm_overlap", 0)
    assert (key.prove(inputs, *rs[0], check=True)["raw"] == single[0]).all()
    ctx.set_option("msm_overlap", 1)
    out3 = key.prove_many(dev.data_ptr(), 3, [r for r, _ in rs[:3]], [s for _, s in rs[:3]], check=True)
    ctx.set_option("msm_overlap", -1)
    key.set_lanes(1)
    out4 = key.prove_many(dev.data_ptr(), 2, [r for r, _ in rs[:2]], [s for _, s in rs[:2]], check=True)
    assert all((a["raw"] == b).all() for a, b in zip(out3, single)) and all((a["raw"] == b).all() for a, b in zip(out4, single))
    key.set_lanes(2)
    # a bad proof in the stream fails the call loudly (unsatisfiable -> GPW_EUNSAT), it is not skipped
    bad = many_host.copy()
    bad[2, circ.info["public"] + 40, 0] ^= 1
    with pytest.raises(gpw.GpwError) as e:
        key.prove_many(bad.ctypes.data, n, [r for r, _ in rs], [s for _, s in rs], check=True)
    assert e.value.code in (-5, -6) and "proof 2" in str(e.value)
    key.close()


def test_wrap_prove_decode_block(ctx, testdata_dir):
    # BASELINE.json configs[0]: the reference's other fixture (degree_bits 12, ConstantGate, no ExponentiationGate)
    d = os.path.join(testdata_dir, "decode_block")
    rd = lambda f: open(os.path.join(d, f), "rb").read()
    circ = gpw.Circuit.compile_verifier(ctx, rd("common_circuit_data.json"))
    inputs = circ.parse_inputs(rd("proof_with_public_inputs.json"), rd("verifier_only_circuit_data.json"))
    w = _solve(ctx, circ, inputs)
    assert circ.r1cs_eval_dev(w.data_ptr()) == 0
    api, _ = verify_testdata(d)
    for kind, op in (("muladd", 1), ("reduce", 2), ("inverse", 3), ("split", 4)):
        exp = [o for k, _, outs in api.hints if k == kind for o in outs]
        assert _wire_ints(w, circ.hint_wires(op)) == exp, kind
    key = gpw.WrapKey(ctx, circ, seed=5)
    pr = key.prove(inputs, 3, 4, check=True)
    assert pr["n_unsatisfied"] == 0
    for g, name in ((1, "Ar"), (2, "Bs"), (1, "Krs"), (1, "commitment"), (1, "pok")):
        assert gpw.host_ec_is_on_curve(g, pr[name]) and pr[name].any()
    key.close()
    circ.close()


def test_poseidon_gl_macro_rejects_noncanonical_state(ctx):
    # the permutation starts with gl.Add = MulAddHint, which refuses operands >= p (goldilocks/base.go:228-232)
    circ = gpw.Circuit.compile_gadget(ctx, "poseidon_gl")
    with pytest.raises(gpw.GpwError) as e:
        _solve(ctx, circ, circ.inputs_from_ints([0] * 12, [ogl.P] + [0] * 11))
    assert e.value.code == -5 and "MulAddHint" in str(e.value)
    circ.close()


def test_gate_circuits_on_gpu(ctx):
    # the three gates without a reference vector (Exponentiation, Constant, Noop) as one-gate circuits solved on the GPU, against
    # the independent evaluation of their constraint polynomials (tests/test_gate_vectors.py); a perturbed expectation is rejected
    from test_gate_vectors import gate_cases, gate_circuit_inputs, qadd, ONE
    rng = random.Random(22)
    for spec, expected, consts, wires in gate_cases(rng):
        circ = gpw.Circuit.compile_gadget(ctx, "gate:" + spec)
        pub, sec = gate_circuit_inputs(expected, consts, wires)
        w = _solve(ctx, circ, circ.inputs_from_ints(pub, sec))
        assert circ.r1cs_eval_dev(w.data_ptr()) == 0, spec
        if expected:
            pub2, _ = gate_circuit_inputs([qadd(expected[0], ONE)] + expected[1:], consts, wires)
            w = _solve(ctx, circ, circ.inputs_from_ints(pub2, sec))
            with pytest.raises(gpw.GpwError) as e:
                circ.r1cs_eval_dev(w.data_ptr())
            assert e.value.code == -6
        circ.close()


def test_circuit_cache_round_trip_on_gpu(ctx, step, tmp_path):
    # gpw_circuit_save / gpw_circuit_load: the reloaded step circuit has the same shape, parses the same inputs and produces the
    # same proof bytes with the same key seed, without re-running the gadget code
    import time
    circ, inputs, d = step
    path = str(tmp_path / "step.circuit")
    circ.save(path)
    t0 = time.perf_counter()
    circ2 = gpw.Circuit.load(ctx, path)
    t_load = time.perf_counter() - t0
    assert circ2.info == circ.info
    rd = lambda f: open(os.path.join(d, f), "rb").read()
    assert (circ2.parse_inputs(rd("proof_with_public_inputs.json"), rd("verifier_only_circuit_data.json")) == inputs).all()
    k1, k2 = gpw.WrapKey(ctx, circ, seed=3), gpw.WrapKey(ctx, circ2, seed=3)
    p1, p2 = k1.prove(inputs, 5, 6), k2.prove(inputs, 5, 6)
    assert (p1["raw"] == p2["raw"]).all()
    print("circuit cache: %.1f MB, load %.2f s" % (os.path.getsize(path) / 1e6, t_load))
    k1.close()
    k2.close()
    circ2.close()
    with pytest.raises(gpw.GpwError):
        gpw.Circuit.load(ctx, os.path.join(d, "common_circuit_data.json"))
