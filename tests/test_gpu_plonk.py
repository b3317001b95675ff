"""PLONK / KZG backend (BASELINE.json configs[3]; benchmark.go:80-190: NewKZGSRS, plonk.Setup, plonk.Prove, plonk.Verify):
proofs made on the GPU are checked by the independent verifier oracle/plonk_verify.py (field arithmetic on Python integers, the
pairing of oracle/pairing.py)."""
import os

import numpy as np
import pytest

import gpw
from oracle import goldilocks as ogl
from oracle import plonk_verify as oplonk
from oracle.engine import Api
from oracle.poseidon import BN254Chip

pytestmark = pytest.mark.gpu
SEED = bytes(range(32, 64))


@pytest.fixture(scope="module")
def ctx():
    c = gpw.Context(0)
    yield c
    c.close()


def _check(key, proof, public):
    return oplonk.verify(oplonk.parse_vk(key.vk()), oplonk.parse_proof(proof), public)


def test_plonk_circuit_without_commitment(ctx, kats):
    # Poseidon-BN254 gadget: multiplication gates + addition chains + public rows, no range checks -> no P2 column
    circ = gpw.Circuit.compile_gadget(ctx, "poseidon_bn254")
    key = gpw.PlonkKey(ctx, circ, SEED)
    assert key.info["has_commit"] == 0 and key.info["qcp_rows"] == 0 and key.info["public_rows"] == 5
    case = kats["poseidon_bn254"][0]
    out, inp = [int(x) for x in case["out"]], [int(x) for x in case["in"]]
    proof = key.prove(circ.inputs_from_ints(out, inp))
    ok, why = _check(key, proof, out)
    assert ok, why
    assert not _check(key, proof, [out[0] ^ 1] + out[1:])[0]                    # another statement
    bad = bytearray(proof)
    bad[10 * 64 + 31] ^= 1                                                      # a(zeta) perturbed
    assert not _check(key, bytes(bad), out)[0]
    bad = bytearray(proof)
    bad[64 * 8: 64 * 9] = proof[64 * 9: 64 * 10]                                # wrong opening proof
    assert not _check(key, bytes(bad), out)[0]
    # a wrong witness is refused by the prover (the quotient is not a polynomial)
    with pytest.raises(gpw.GpwError) as e:
        key.prove(circ.inputs_from_ints([out[0] ^ 1] + out[1:], inp))
    assert e.value.code == -6
    # a second statement with the same key
    case = kats["poseidon_bn254"][1]
    out2, inp2 = [int(x) for x in case["out"]], [int(x) for x in case["in"]]
    ok, why = _check(key, key.prove(circ.inputs_from_ints(out2, inp2)), out2)
    assert ok, why
    key.close()
    circ.close()


def test_plonk_circuit_with_range_check_commitment(ctx, kats):
    # QE mul / div gadget: Goldilocks hints + range checks -> committed limb wires on Qcp rows, challenge = hash([P2]) is a
    # public input the verifier derives itself
    circ = gpw.Circuit.compile_gadget(ctx, "qe_mul_div")
    a = tuple(map(int, kats["qe_mul"]["a"]))
    b = tuple(map(int, kats["qe_mul"]["b"]))
    ch = ogl.Chip(Api(trace=False))
    m = ch.MulExtension(a, b)
    d, _ = ch.DivExtension(a, b)
    key = gpw.PlonkKey(ctx, circ, SEED)
    assert key.info["has_commit"] == 1 and key.info["qcp_rows"] == circ.info["limb_wires"] + 65536
    pub = list(m) + list(d)
    proof = key.prove(circ.inputs_from_ints(pub, list(a) + list(b)))
    ok, why = _check(key, proof, pub)
    assert ok, why
    bad = bytearray(proof)
    bad[64 * 3: 64 * 4] = proof[0:64]                                           # another P2 commitment -> another challenge
    assert not _check(key, bytes(bad), pub)[0]
    assert not _check(key, proof, pub[:-1] + [pub[-1] ^ 1])[0]
    print("plonk qe_mul_div:", key.info, key.last_stats())
    key.close()
    circ.close()


@pytest.mark.parametrize("name", ["decode_block", "step"])
def test_plonk_wrap_of_reference_fixture(ctx, testdata_dir, name):
    # BASELINE configs[3] on the reference's fixtures: the whole verifier circuit under PLONK (2^25 rows each; step fits with
    # 92 k rows to spare thanks to the constant folding into qC)
    d = os.path.join(testdata_dir, name)
    rd = lambda f: open(os.path.join(d, f), "rb").read()
    circ = gpw.Circuit.compile_verifier(ctx, rd("common_circuit_data.json"), rd("verifier_only_circuit_data.json"))
    inputs = circ.parse_inputs(rd("proof_with_public_inputs.json"), rd("verifier_only_circuit_data.json"))
    key = gpw.PlonkKey(ctx, circ, SEED)
    proof = key.prove(inputs)
    public = [int(x) for x in gpw.limbs_to_ints(inputs[:circ.info["public"]])]
    ok, why = _check(key, proof, public)
    assert ok, why
    assert key.info["logN"] == 25
    if public:
        assert not _check(key, proof, public[:-1] + [public[-1] ^ 1])[0]
    print("plonk %s:" % name, key.info, key.last_stats())
    key.close()
    circ.close()
