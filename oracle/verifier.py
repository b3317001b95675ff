"""ORACLE (test infrastructure only). Top-level dataflow, restating verifier/verifier.go."""
import os

from .engine import Api
from . import goldilocks as gl
from . import types
from .poseidon import GoldilocksChip, BN254Chip
from .challenger import Chip as ChallengerChip
from .fri import Chip as FriChip
from .plonk import PlonkChip


class ProofChallenges:
    pass


class VerifierChip:
    def __init__(self, api, common_data):
        """verifier.go:24-39"""
        self.api = api
        self.glChip = gl.Chip(api)
        self.friChip = FriChip(api, common_data, common_data.FriParams)
        self.plonkChip = PlonkChip(api, common_data)
        self.poseidonGlChip = GoldilocksChip(api)
        self.poseidonBN254Chip = BN254Chip(api)
        self.commonData = common_data
        self.phase_counts = {}

    def GetPublicInputsHash(self, public_inputs):
        return self.poseidonGlChip.HashNoPad(public_inputs)

    def GetChallenges(self, proof, public_inputs_hash, verifier_data):
        """verifier.go:45-82"""
        n = self.commonData.NumChallenges
        ch = ChallengerChip(self.api)
        ch.ObserveBN254Hash(verifier_data.CircuitDigest)
        ch.ObserveHash(public_inputs_hash)
        ch.ObserveCap(proof.WiresCap)
        pc = ProofChallenges()
        pc.PlonkBetas = ch.GetNChallenges(n)
        pc.PlonkGammas = ch.GetNChallenges(n)
        ch.ObserveCap(proof.PlonkZsPartialProductsCap)
        pc.PlonkAlphas = ch.GetNChallenges(n)
        ch.ObserveCap(proof.QuotientPolysCap)
        pc.PlonkZeta = ch.GetExtensionChallenge()
        ch.ObserveOpenings(self.friChip.ToOpenings(proof.Openings))
        pc.FriChallenges = ch.GetFriChallenges(proof.OpeningProof.CommitPhaseMerkleCaps, proof.OpeningProof.FinalPoly,
                                               proof.OpeningProof.PowWitness, self.commonData.FriConfig)
        self.n_duplex = ch.n_duplex
        return pc

    def rangeCheckProof(self, proof):
        """verifier.go:84-141"""
        o = proof.Openings
        for group in (o.Constants, o.PlonkSigmas, o.Wires, o.PlonkZs, o.PlonkZsNext, o.PartialProducts,
                      o.QuotientPolys):
            for qe in group:
                self.glChip.RangeCheckQE(qe)
        for qr in proof.OpeningProof.QueryRoundProofs:
            for leaf, _ in qr.EvalsProofs:
                for e in leaf:
                    self.glChip.RangeCheck(e)
            for evals, _ in qr.Steps:
                for e in evals:
                    self.glChip.RangeCheckQE(e)
        for c in proof.OpeningProof.FinalPoly:
            self.glChip.RangeCheckQE(c)
        self.glChip.RangeCheck(proof.OpeningProof.PowWitness)

    def _snap(self, name):
        self.phase_counts[name] = dict(self.api.counts)

    def Verify(self, proof, public_inputs, verifier_data):
        """verifier.go:143-170"""
        self.rangeCheckProof(proof)
        self._snap("rangeCheckProof")
        pih = self.GetPublicInputsHash(public_inputs)
        self._snap("publicInputsHash")
        ch = self.GetChallenges(proof, pih, verifier_data)
        self._snap("challenger")
        self.challenges = ch
        self.plonkChip.Verify(ch, proof.Openings, pih)
        self._snap("plonk")
        caps = [verifier_data.ConstantSigmasCap, proof.WiresCap, proof.PlonkZsPartialProductsCap,
                proof.QuotientPolysCap]
        self.friChip.VerifyFriProof(self.friChip.GetInstance(ch.PlonkZeta), self.friChip.ToOpenings(proof.Openings),
                                    ch.FriChallenges, caps, proof.OpeningProof)
        self._snap("fri")


def verify_testdata(dirpath, trace=True):
    """ExampleVerifierCircuit.Define (verifier/util.go:19-24) on one testdata directory."""
    common = types.read_common_circuit_data(os.path.join(dirpath, "common_circuit_data.json"))
    proof, pis = types.read_proof_with_public_inputs(os.path.join(dirpath, "proof_with_public_inputs.json"))
    vd = types.read_verifier_only_circuit_data(os.path.join(dirpath, "verifier_only_circuit_data.json"))
    api = Api(trace=trace)
    chip = VerifierChip(api, common)
    chip.Verify(proof, pis, vd)
    return api, chip
