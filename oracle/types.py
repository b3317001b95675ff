"""ORACLE (test infrastructure only). Input wire format: restates types/common_data.go,
types/types.go, types/deserialize.go and variables/deserialize.go (JSON -> plain Python values).
"""
import json
import re


class FriConfig:
    def __init__(self, d):
        self.RateBits = d["rate_bits"]
        self.CapHeight = d["cap_height"]
        self.ProofOfWorkBits = d["proof_of_work_bits"]
        self.NumQueryRounds = d["num_query_rounds"]

    def Rate(self):
        return 1.0 / float(1 << self.RateBits)


class FriParams:
    """types/types.go:21-60"""

    def __init__(self, d):
        self.Config = FriConfig(d["config"])
        self.Hiding = d["hiding"]
        self.DegreeBits = d["degree_bits"]
        self.ReductionArityBits = list(d["reduction_arity_bits"])

    def TotalArities(self):
        return sum(self.ReductionArityBits)

    def LdeBits(self):
        return self.DegreeBits + self.Config.RateBits

    def FinalPolyLen(self):
        return 1 << (self.DegreeBits - self.TotalArities())


class CommonCircuitData:
    """types/common_data.go:61-127 / types/types.go:74-86"""

    def __init__(self, raw):
        c = raw["config"]
        self.NumWires = c["num_wires"]
        self.NumRoutedWires = c["num_routed_wires"]
        self.NumChallenges = c["num_challenges"]
        self.FriConfig = FriConfig(c["fri_config"])
        self.FriParams = FriParams(raw["fri_params"])
        self.DegreeBits = raw["fri_params"]["degree_bits"]
        self.GateIds = list(raw["gates"])
        self.SelectorIndices = list(raw["selectors_info"]["selector_indices"])
        self.SelectorGroups = [(g["start"], g["end"]) for g in raw["selectors_info"]["groups"]]
        self.QuotientDegreeFactor = raw["quotient_degree_factor"]
        self.NumGateConstraints = raw["num_gate_constraints"]
        self.NumConstants = raw["num_constants"]
        self.NumPublicInputs = raw["num_public_inputs"]
        self.KIs = list(raw["k_is"])
        self.NumPartialProducts = raw["num_partial_products"]
        if raw["fri_params"]["hiding"]:
            raise ValueError("Circuit has hiding enabled, which is not supported")  # common_data.go:121-124


def _qe_list(xs):
    return [(int(a), int(b)) for a, b in xs]


class Proof:
    pass


def read_common_circuit_data(path):
    return CommonCircuitData(json.load(open(path)))


def read_proof_with_public_inputs(path):
    """types/deserialize.go:92-108 + variables/deserialize.go:114-147"""
    raw = json.load(open(path))
    p = raw["proof"]
    proof = Proof()
    proof.WiresCap = [int(x) for x in p["wires_cap"]]
    proof.PlonkZsPartialProductsCap = [int(x) for x in p["plonk_zs_partial_products_cap"]]
    proof.QuotientPolysCap = [int(x) for x in p["quotient_polys_cap"]]
    o = p["openings"]
    op = Proof()
    op.Constants = _qe_list(o["constants"])
    op.PlonkSigmas = _qe_list(o["plonk_sigmas"])
    op.Wires = _qe_list(o["wires"])
    op.PlonkZs = _qe_list(o["plonk_zs"])
    op.PlonkZsNext = _qe_list(o["plonk_zs_next"])
    op.PartialProducts = _qe_list(o["partial_products"])
    op.QuotientPolys = _qe_list(o["quotient_polys"])
    proof.Openings = op
    f = p["opening_proof"]
    fp = Proof()
    fp.CommitPhaseMerkleCaps = [[int(x) for x in cap] for cap in f["commit_phase_merkle_caps"]]
    fp.QueryRoundProofs = []
    for q in f["query_round_proofs"]:
        qr = Proof()
        qr.EvalsProofs = []
        for leaf, mp in q["initial_trees_proof"]["evals_proofs"]:   # 2-tuple, deserialize.go:45-72
            qr.EvalsProofs.append(([int(x) for x in leaf], [int(x) for x in mp["siblings"]]))
        qr.Steps = [(_qe_list(s["evals"]), [int(x) for x in s["merkle_proof"]["siblings"]]) for s in q["steps"]]
        fp.QueryRoundProofs.append(qr)
    fp.FinalPoly = _qe_list(f["final_poly"]["coeffs"])
    fp.PowWitness = int(f["pow_witness"])
    proof.OpeningProof = fp
    return proof, [int(x) for x in raw["public_inputs"]]


class VerifierOnlyCircuitData:
    pass


def read_verifier_only_circuit_data(path):
    raw = json.load(open(path))
    v = VerifierOnlyCircuitData()
    v.ConstantSigmasCap = [int(x) for x in raw["constants_sigmas_cap"]]
    v.CircuitDigest = int(raw["circuit_digest"])
    return v
