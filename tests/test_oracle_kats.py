"""Pins the oracle against every golden vector the reference's own tests hold for the hot path
(SURVEY 8c). CPU-only."""
import os

import pytest

from oracle.engine import Api, AssertionFailed
from oracle import goldilocks as gl
from oracle import gates as og
from oracle.poseidon import GoldilocksChip, BN254Chip
from oracle.types import read_common_circuit_data
from oracle.verifier import verify_testdata


def test_poseidon_gl_perm_zero(kats):
    # poseidon/goldilocks_test.go:37-59
    out = GoldilocksChip(Api()).Poseidon([0] * 12)
    assert out == [int(x) for x in kats["poseidon_gl_perm_zero"]]


def test_public_inputs_hash(kats):
    # poseidon/public_inputs_hash_test.go:43-60
    k = kats["public_inputs_hash"]
    assert GoldilocksChip(Api()).HashNoPad([int(x) for x in k["in"]]) == [int(x) for x in k["out"]]


def test_poseidon_bn254(kats):
    # poseidon/bn254_test.go:31-97
    for case in kats["poseidon_bn254"]:
        out = BN254Chip(Api()).Poseidon([int(x) for x in case["in"]])
        assert out == [int(x) for x in case["out"]]


def test_qe_mul_div(kats):
    # goldilocks/quadratic_extension_test.go:25-94
    c = gl.Chip(Api())
    k = kats["qe_mul"]
    assert c.MulExtension(tuple(map(int, k["a"])), tuple(map(int, k["b"]))) == tuple(map(int, k["out"]))
    k = kats["qe_div"]
    res, has = c.DivExtension(tuple(map(int, k["a"])), tuple(map(int, k["b"])))
    assert res == tuple(map(int, k["out"])) and has == 1


def test_muladd(kats):
    # goldilocks/base_test.go:97-116
    k = kats["muladd"]
    assert gl.Chip(Api()).MulAdd(int(k["a"]), int(k["b"]), int(k["c"])) == int(k["out"])


def test_range_check_boundaries():
    # goldilocks/base_test.go:26-44: accepts 0, 1, p-1; rejects p
    for x in (0, 1, gl.P - 1):
        gl.Chip(Api()).RangeCheck(x)
    with pytest.raises(ValueError):
        gl.Chip(Api()).RangeCheck(gl.P)
    # hi == 2^32-1 forces lo == 0: p-1 passes (lo=0); 2^64-1 is >= p so the hint errors
    with pytest.raises(ValueError):
        gl.split_limbs_hint((1 << 64) - 1)


def test_hint_edge_cases():
    # SURVEY 8a parity trap 10
    assert gl.inverse_hint(0) == 0
    assert gl.reduce_hint(gl.P) == (1, 0)
    assert gl.reduce_hint((1 << 200) + 5) == (((1 << 200) + 5) // gl.P, ((1 << 200) + 5) % gl.P)
    with pytest.raises(ValueError):
        gl.mul_add_hint(gl.P, 1, 0)
    # Sub(a, b) multiplies b by p-1 (trap 1): quotient ~ b
    api = Api()
    gl.Chip(api).Sub(5, 3)
    kind, inp, out = api.hints[0]
    assert kind == "muladd" and inp == (3, gl.P - 1, 5) and out == (3, 2)


_GATE_CTORS = {
    "PublicInputGate": lambda a, w: og.PublicInputGate(),
    "BaseSumGate": lambda a, w: og.BaseSumGate(*a),
    "ArithmeticGate": lambda a, w: og.ArithmeticGate(*a),
    "RandomAccessGate": lambda a, w: og.RandomAccessGate(*a),
    "PoseidonGate": lambda a, w: og.PoseidonGate(),
    "ArithmeticExtensionGate": lambda a, w: og.ArithmeticExtensionGate(*a),
    "MultiplicationExtensionGate": lambda a, w: og.MultiplicationExtensionGate(*a),
    "ReducingExtensionGate": lambda a, w: og.ReducingExtensionGate(*a),
    "ReducingGate": lambda a, w: og.ReducingGate(*a),
    "CosetInterpolationGate": lambda a, w: og.CosetInterpolationGate(a[0], a[1], [int(x) for x in w]),
    "PoseidonMdsGate": lambda a, w: og.PoseidonMdsGate(),
}


def test_gates(kats, testdata_dir):
    # plonk/gates/gates_test.go:712-768
    g = kats["gates"]
    common = read_common_circuit_data(os.path.join(testdata_dir, g["common_data"], "common_circuit_data.json"))
    num_selectors = len(common.SelectorGroups)
    qe = lambda v: [(int(a), int(b)) for a, b in v]
    consts = qe(g["vectors"]["localConstants"])[num_selectors:]
    wires = qe(g["vectors"]["localWires"])
    pih = [int(x) for x in g["public_inputs_hash"]]
    assert len(g["tests"]) == 11
    for t in g["tests"]:
        api = Api(trace=False)
        gate = _GATE_CTORS[t["gate"]](t["args"], t.get("weights"))
        got = gate.EvalUnfiltered(api, gl.Chip(api), consts, wires, pih)
        assert got == qe(g["vectors"][t["expected"]]), t["gate"]


def test_gate_ids_parse(testdata_dir):
    # plonk/gates/gates.go:37-54 on both fixtures' gate lists
    for d in ("step", "decode_block"):
        common = read_common_circuit_data(os.path.join(testdata_dir, d, "common_circuit_data.json"))
        insts = [og.GateInstanceFromId(x) for x in common.GateIds]
        assert len(insts) == 13
    with pytest.raises(ValueError):
        og.GateInstanceFromId("FooGate")


def test_verify_decode_block_with_challenger_goldens(kats, testdata_dir):
    # fri/fri_test.go:23-133 (8 challenger goldens + full FRI), plonk/plonk_test.go:39-66
    api, chip = verify_testdata(os.path.join(testdata_dir, "decode_block"))
    k = kats["challenger_decode_block"]
    ch = chip.challenges
    assert ch.PlonkBetas[0] == int(k["plonk_beta0"])
    assert ch.PlonkGammas[0] == int(k["plonk_gamma0"])
    assert ch.PlonkAlphas[0] == int(k["plonk_alpha0"])
    assert ch.PlonkZeta[0] == int(k["plonk_zeta0"])
    f = ch.FriChallenges
    assert f.FriAlpha[0] == int(k["fri_alpha0"])
    assert f.FriBetas[0][0] == int(k["fri_beta00"])
    assert f.FriPowResponse == int(k["fri_pow_response"])
    assert f.FriQueryIndices[0] == int(k["fri_query_index0"])
    assert api.counts == {"muladd": 42424, "reduce": 144072, "split": 240308, "inverse": 1849,
                          "poseidon_gl": 125, "poseidon_bn254": 2604}


def test_verify_step_end_to_end(testdata_dir):
    # verifier/verifier_test.go:13-41: all in-circuit assertions hold on testdata/step
    api, chip = verify_testdata(os.path.join(testdata_dir, "step"))
    assert api.counts == {"muladd": 44186, "reduce": 151410, "split": 251232, "inverse": 1849,
                          "poseidon_gl": 134, "poseidon_bn254": 2772}
    assert chip.n_duplex == 129
    assert len(chip.friChip.merkle_roots) == 28 * 6
    assert all(a == b for a, b in chip.friChip.merkle_roots)
    assert sum((b + 15) // 16 for _, b in api.range_checks) == 2462493


def test_verify_rejects_tampered_proof(testdata_dir, tmp_path):
    import json, shutil
    src = os.path.join(testdata_dir, "decode_block")
    for f in os.listdir(src):
        shutil.copy(os.path.join(src, f), tmp_path / f)
    p = json.load(open(tmp_path / "proof_with_public_inputs.json"))
    p["proof"]["openings"]["wires"][3][0] ^= 1
    json.dump(p, open(tmp_path / "proof_with_public_inputs.json", "w"))
    with pytest.raises(AssertionFailed):
        verify_testdata(str(tmp_path), trace=False)
