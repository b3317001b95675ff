"""Real Groth16 setup -> GPU proof -> gnark raw bytes -> verified by the oracle's pairing check.

Mirrors /root/reference/benchmark.go:214-217 (groth16.Setup), :249 (Prove), :266 (Verify), :272-291 (WriteRawTo and the
a/b/c layout read back from it). The verifier is oracle/pairing.py (pure Python, shares nothing with csrc/)."""
import os

import numpy as np
import pytest

import gpw
from oracle import bn254 as ob
from oracle import goldilocks as ogl
from oracle import pairing as opair
from oracle.engine import Api

pytestmark = pytest.mark.gpu
SEED = bytes(range(32))


@pytest.fixture(scope="module")
def ctx():
    c = gpw.Context(0)
    yield c
    c.close()


def _public_ints(circ, inputs):
    return [int(x) for x in gpw.limbs_to_ints(inputs[:circ.info["public"]])]


def _verify(key, proof, public):
    vk = opair.parse_vk_raw(key.vk_raw())
    pr = opair.parse_proof_raw(key.proof_raw(proof))
    return opair.groth16_verify(vk, pr, public), vk, pr


def test_small_circuit_real_setup_proof_verifies(ctx, kats, tmp_path):
    circ = gpw.Circuit.compile_gadget(ctx, "qe_mul_div")
    a = tuple(map(int, kats["qe_mul"]["a"]))
    b = tuple(map(int, kats["qe_mul"]["b"]))
    ch = ogl.Chip(Api(trace=False))
    m = ch.MulExtension(a, b)
    d, _ = ch.DivExtension(a, b)
    inputs = circ.inputs_from_ints(list(m) + list(d), list(a) + list(b))
    key = gpw.WrapKey.setup(ctx, circ, SEED)
    proof = key.prove(inputs, 0x1234, 0x5678)
    public = _public_ints(circ, inputs)
    (ok, why), vk, pr = _verify(key, proof, public)
    assert ok, why
    # raw layout: benchmark.go:283-290 reads a = bytes[0:64], b = [64:192] (X.A1, X.A0, Y.A1, Y.A0), c = [192:256]
    raw = key.proof_raw(proof)
    assert len(raw) == 388 and int.from_bytes(raw[256:260], "big") == 1
    ax, ay = gpw.points_to_ints(1, proof["Ar"])[0]
    assert raw[:64] == ax.to_bytes(32, "big") + ay.to_bytes(32, "big")
    (bx0, bx1), (by0, by1) = gpw.points_to_ints(2, proof["Bs"])[0]
    assert raw[64:192] == b"".join(v.to_bytes(32, "big") for v in (bx1, bx0, by1, by0))
    # the verifier recomputes the commitment challenge with hashlib: it must be the wire value the prover used
    assert opair.hash_to_fr(opair.g1_raw(pr["commitments"][0])) == proof["challenge"]
    # a wrong public input, a perturbed proof element, a wrong commitment: all rejected
    assert not opair.groth16_verify(vk, pr, [public[0] ^ 1] + public[1:])[0]
    bad = dict(pr, Krs=ob.ec_add(1, pr["Krs"], ob.G1_GEN))
    assert not opair.groth16_verify(vk, bad, public)[0]
    bad = dict(pr, commitments=[ob.ec_add(1, pr["commitments"][0], ob.G1_GEN)])
    assert not opair.groth16_verify(vk, bad, public)[0]
    # library-sampled blinding (r = s = NULL): two proofs of the same statement differ and both verify
    p1, p2 = key.prove(inputs), key.prove(inputs)
    assert (p1["Ar"] != p2["Ar"]).any() and (p1["commitment"] == p2["commitment"]).all()
    assert _verify(key, p1, public)[0][0] and _verify(key, p2, public)[0][0]
    # same seed -> same key; pk / vk files round-trip and the loaded key proves the same bytes
    pk_path, vk_path = str(tmp_path / "proving.key"), str(tmp_path / "verifying.key")
    key.save(pk_path, vk_path)
    assert open(vk_path, "rb").read() == key.vk_raw()
    key2 = gpw.WrapKey.load(ctx, circ, pk_path, vk_path)
    assert key2.vk_raw() == key.vk_raw()
    assert key2.proof_raw(key2.prove(inputs, 0x1234, 0x5678)) == raw
    key2.close()
    key3 = gpw.WrapKey.setup(ctx, circ, SEED)
    assert key3.vk_raw() == key.vk_raw()
    key3.close()
    key4 = gpw.WrapKey.setup(ctx, circ, bytes(32))
    assert key4.vk_raw() != key.vk_raw()
    key4.close()
    # a key file of another circuit is refused
    other = gpw.Circuit.compile_gadget(ctx, "poseidon_bn254")
    with pytest.raises(gpw.GpwError):
        gpw.WrapKey.load(ctx, other, pk_path, vk_path)
    other.close()
    # the DummySetup analogue has no verifying key
    dummy = gpw.WrapKey(ctx, circ, seed=1)
    with pytest.raises(gpw.GpwError):
        dummy.vk_raw()
    dummy.close()
    # an unsatisfied witness never yields a proof, whatever `check` says
    wrong = circ.inputs_from_ints([m[0] ^ 1] + list(m[1:]) + list(d), list(a) + list(b))
    for check in (True, False):
        with pytest.raises(gpw.GpwError) as e:
            key.prove(wrong, 1, 2, check=check)
        assert e.value.code == -6
    key.close()
    circ.close()


@pytest.mark.parametrize("name", ["step", "decode_block"])
def test_wrap_proof_of_reference_fixture_verifies(ctx, testdata_dir, name):
    # benchmark.go:192-266 end to end on the reference's own fixtures: Compile, Setup, NewWitness + Prove, Verify
    d = os.path.join(testdata_dir, name)
    rd = lambda f: open(os.path.join(d, f), "rb").read()
    circ = gpw.Circuit.compile_verifier(ctx, rd("common_circuit_data.json"))
    inputs = circ.parse_inputs(rd("proof_with_public_inputs.json"), rd("verifier_only_circuit_data.json"))
    key = gpw.WrapKey.setup(ctx, circ, SEED)
    proof = key.prove(inputs)  # r, s from the library's CSPRNG
    public = _public_ints(circ, inputs)
    (ok, why), vk, pr = _verify(key, proof, public)
    assert ok, why
    assert len(vk["K"]) == 1 + circ.info["public"] + 1
    if public:   # (decode_block has no public inputs)
        assert not opair.groth16_verify(vk, pr, public[:-1] + [public[-1] ^ 1])[0]
    assert not opair.groth16_verify(vk, dict(pr, Ar=ob.ec_add(1, pr["Ar"], ob.G1_GEN)), public)[0]
    # several proofs in flight with a real key: every one verifies
    many = np.ascontiguousarray(np.tile(inputs, (3, 1, 1)))
    key.set_lanes(3)
    for p in key.prove_many(many.ctypes.data, 3):
        assert _verify(key, p, public)[0][0]
    key.close()
    circ.close()


def test_bound_and_baked_circuit_forms(ctx, testdata_dir):
    # verifier/util.go:10-24: Proof and VerifierOnlyCircuitData are `gnark:"-"` (compile-time constants) in the reference.
    # Form 1 bakes the verifier-only data (statement bound to one inner circuit), form 2 the proof too (benchmark.go:33-55).
    # Both produce the same 449 k reference hints and proofs that verify; documents of another circuit / proof are refused.
    import json
    d = os.path.join(testdata_dir, "decode_block")
    rd = lambda f: open(os.path.join(d, f), "rb").read()
    common, proof_json, vod = rd("common_circuit_data.json"), rd("proof_with_public_inputs.json"), rd("verifier_only_circuit_data.json")
    free = gpw.Circuit.compile_verifier(ctx, common)
    counts = {k: free.info[k] for k in ("muladd", "reduce", "inverse", "split")}
    cap = 1 << json.loads(common)["config"]["fri_config"]["cap_height"]
    n_free_secret = free.info["secret"]
    free.close()
    for baked_proof in (False, True):
        circ = gpw.Circuit.compile_verifier(ctx, common, vod, proof_json if baked_proof else None)
        assert {k: circ.info[k] for k in counts} == counts
        assert circ.info["secret"] == (0 if baked_proof else n_free_secret - cap - 1)
        inputs = circ.parse_inputs(proof_json, vod)
        assert inputs.shape[0] == circ.info["public"] + circ.info["secret"]
        key = gpw.WrapKey.setup(ctx, circ, SEED)
        proof = key.prove(inputs)
        (ok, why), vk, pr = _verify(key, proof, _public_ints(circ, inputs))
        assert ok, why
        v2 = json.loads(vod)
        v2["circuit_digest"] = str(int(v2["circuit_digest"]) ^ 1)
        with pytest.raises(gpw.GpwError):
            circ.parse_inputs(proof_json, json.dumps(v2))
        if baked_proof:
            p2 = json.loads(proof_json)
            p2["proof"]["opening_proof"]["pow_witness"] ^= 1
            with pytest.raises(gpw.GpwError):
                circ.parse_inputs(json.dumps(p2), vod)
        key.close()
        circ.close()
