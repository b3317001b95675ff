"""ORACLE (test infrastructure only). FRI verifier, restating fri/fri.go and fri/fri_utils.go.

Openings are a list of two batches (lists of QE); an InstanceInfo is (oracles, batches) with
batches = [(point, [(oracle_idx, poly_idx), ...]), ...].
"""
from . import goldilocks as gl
from .poseidon import BN254Chip


# ---- fri_utils.go ---------------------------------------------------------------------------
def _range(oracle, start, end):
    return [(oracle, i) for i in range(start, end)]


def num_preprocessed_polys(c):
    # fri_utils.go:60-74: last element of sigmasRange = NumConstants + NumRoutedWires
    return c.NumConstants + c.NumRoutedWires


def num_zs_partial_products_polys(c):
    return c.NumChallenges * (1 + c.NumPartialProducts)


def num_quotient_polys(c):
    return c.NumChallenges * c.QuotientDegreeFactor


def fri_all_polys(c):
    """fri_utils.go:141-149"""
    return (_range(0, 0, num_preprocessed_polys(c)) + _range(1, 0, c.NumWires)
            + _range(2, 0, num_zs_partial_products_polys(c)) + _range(3, 0, num_quotient_polys(c)))


def fri_zs_polys(c):
    return _range(2, 0, c.NumChallenges)


def fri_oracles(c):
    """fri_utils.go:120-139 -> [(num_polys, blinding)]"""
    return [(num_preprocessed_polys(c), False), (c.NumWires, True),
            (num_zs_partial_products_polys(c), True), (num_quotient_polys(c), True)]


def assert_noncanonical_indices_ok(fri_params):
    """fri_utils.go:153-163 (the only floating point on the path; a config sanity check)"""
    num_ambiguous = float((1 << 64) - gl.P)
    if num_ambiguous / float(gl.P) >= fri_params.Config.Rate() * 1e-5:
        raise ValueError("non-negligible portion of field elements permits non-canonical encodings")


def validate_fri_proof_shape(proof, instance, params):
    """fri_utils.go:167-228"""
    cap_height = params.Config.CapHeight
    for cap in proof.CommitPhaseMerkleCaps:
        if (1 << cap_height) != len(cap):
            raise ValueError("config cap_height does not match commit_phase_merkle_caps")
    oracles = instance[0]
    for qr in proof.QueryRoundProofs:
        if len(qr.EvalsProofs) != len(oracles):
            raise ValueError("eval proofs length is not equal to instance oracles length")
        for (leaf, siblings), (num_polys, blinding) in zip(qr.EvalsProofs, oracles):
            salt = 4 if (blinding and params.Hiding) else 0
            if len(leaf) != num_polys + salt:
                raise ValueError("eval proof leaf length doesn't match oracle info")
            if len(siblings) + cap_height != params.LdeBits():
                raise ValueError("length of merkle proof + capHeight doesn't match lde_bits from params")
        if len(qr.Steps) != len(params.ReductionArityBits):
            raise ValueError("length of steps != params.reduction_arity_bits")
        codeword_len_bits = params.LdeBits()
        for (evals, siblings), arity_bits in zip(qr.Steps, params.ReductionArityBits):
            codeword_len_bits -= arity_bits
            if len(evals) != (1 << arity_bits):
                raise ValueError("len evals doesn't match arity")
            if len(siblings) + cap_height != codeword_len_bits:
                raise ValueError("len merkleProof doesn't match codewordLenBits")
    if len(proof.FinalPoly) != params.FinalPolyLen():
        raise ValueError("len finalPoly doesn't match params FinalPolyLen")


def _bitrev8(i):
    return int("{:08b}".format(i)[::-1], 2)


class Chip:
    """fri/fri.go:17-38"""

    def __init__(self, api, common_data, fri_params):
        self.api = api
        self.gl = gl.Chip(api)
        self.poseidonBN254Chip = BN254Chip(api)
        self.commonData = common_data
        self.friParams = fri_params
        self.merkle_roots = []   # (computed root, selected cap entry) for tests

    def GetInstance(self, zeta):
        """fri.go:40-61"""
        g = gl.primitive_root_of_unity(self.commonData.DegreeBits)
        zeta_next = self.gl.MulExtension((g, 0), zeta)
        return (fri_oracles(self.commonData),
                [(zeta, fri_all_polys(self.commonData)), (zeta_next, fri_zs_polys(self.commonData))])

    def ToOpenings(self, c):
        """fri.go:63-73"""
        values = list(c.Constants) + list(c.PlonkSigmas) + list(c.Wires) + list(c.PlonkZs) \
            + list(c.PartialProducts) + list(c.QuotientPolys)
        return [values, list(c.PlonkZsNext)]

    def assertLeadingZeros(self, pow_witness, fri_config):
        self.gl.RangeCheckWithMaxBits(pow_witness, 64 - fri_config.ProofOfWorkBits)

    def fromOpeningsAndAlpha(self, openings, alpha):
        return [self.gl.ReduceWithPowers(batch, alpha) for batch in openings]

    def verifyMerkleProofToCapWithCapIndex(self, leaf_data, leaf_index_bits, cap_index_bits, merkle_cap, siblings):
        """fri.go:97-144"""
        current = self.poseidonBN254Chip.HashOrNoop(leaf_data)
        for i, sibling in enumerate(siblings):
            bit = leaf_index_bits[i]
            inputs = [0, 0, self.api.Select(bit, sibling, current), self.api.Select(bit, current, sibling)]
            current = self.poseidonBN254Chip.Poseidon(inputs)[0]
        if len(cap_index_bits) != 4 or len(merkle_cap) != 16:
            raise ValueError("capIndexBits length should be 4 and the merkleCap length should be 16")
        leaf_lookups = [self.api.Lookup2(cap_index_bits[0], cap_index_bits[1], *merkle_cap[4 * i:4 * i + 4])
                        for i in range(4)]
        entry = self.api.Lookup2(cap_index_bits[2], cap_index_bits[3], *leaf_lookups)
        self.merkle_roots.append((current, entry))
        self.api.AssertIsEqual(current, entry)

    def verifyInitialProof(self, x_index_bits, evals_proofs, initial_merkle_caps, cap_index_bits):
        if len(evals_proofs) != len(initial_merkle_caps):
            raise ValueError("length of eval proofs in fri proof should equal length of initial merkle caps")
        for (evals, siblings), cap in zip(evals_proofs, initial_merkle_caps):
            self.verifyMerkleProofToCapWithCapIndex(evals, x_index_bits, cap_index_bits, cap, siblings)

    def expFromBitsConstBase(self, base, exponent_bits):
        """fri.go:159-185"""
        product = 1
        for i, bit in enumerate(exponent_bits):
            base_pow = pow(base, 1 << i, gl.P)
            base_pow_var = (base_pow - 1) % (1 << 64)     # Go uint64 arithmetic: basePow.Uint64() - 1
            product = self.gl.Add(self.gl.Mul(self.gl.Mul(base_pow_var, product), bit), product)
        return product

    def calculateSubgroupX(self, x_index_bits, n_log):
        """fri.go:187-206"""
        base = gl.primitive_root_of_unity(n_log)
        product = self.expFromBitsConstBase(base, list(reversed(x_index_bits)))
        return self.gl.Mul(gl.MULTIPLICATIVE_GROUP_GENERATOR, product)

    def friCombineInitial(self, instance, evals_proofs, fri_alpha, subgroup_x_qe, precomputed_reduced_eval):
        """fri.go:208-251"""
        s = gl.ZERO_QE
        batches = instance[1]
        assert len(batches) == len(precomputed_reduced_eval)
        for (point, polys), reduced_openings in zip(batches, precomputed_reduced_eval):
            evals = [(evals_proofs[o][0][p], 0) for (o, p) in polys]
            reduced_evals = self.gl.ReduceWithPowers(evals, fri_alpha)
            numerator = self.gl.SubExtensionNoReduce(reduced_evals, reduced_openings)
            denominator = self.gl.SubExtension(subgroup_x_qe, point)
            s = self.gl.MulExtension(self.gl.ExpExtension(fri_alpha, len(evals)), s)
            inv, has_inv = self.gl.InverseExtension(denominator)
            self.api.AssertIsEqual(has_inv, 1)
            s = self.gl.MulAddExtension(numerator, inv, s)
        return s

    def finalPolyEval(self, final_poly, point):
        ret = gl.ZERO_QE
        for c in reversed(final_poly):
            ret = self.gl.MulAddExtension(ret, point, c)
        return ret

    def interpolate(self, x, x_points, y_points, barycentric_weights):
        """fri.go:261-312"""
        assert len(x_points) == len(y_points) == len(barycentric_weights)
        l_x = gl.ONE_QE
        for xp in x_points:
            l_x = self.gl.SubMulExtension(x, xp, l_x)
        s = gl.ZERO_QE
        lookup_from_points = 1
        for i in range(len(x_points)):
            quotient, has_quotient = self.gl.DivExtension(barycentric_weights[i], self.gl.SubExtension(x, x_points[i]))
            lookup_from_points = self.api.Mul(has_quotient, lookup_from_points)
            s = self.gl.AddExtension(self.gl.MulExtension(y_points[i], quotient), s)
        interpolation = self.gl.MulExtension(l_x, s)
        lookup_val = gl.ZERO_QE
        for i in range(len(x_points)):
            lookup_val = self.gl.Lookup(self.gl.IsZero(self.gl.SubExtension(x, x_points[i])), lookup_val, y_points[i])
        return self.gl.Lookup(lookup_from_points, lookup_val, interpolation)

    def computeEvaluation(self, x, x_index_within_coset_bits, arity_bits, evals, beta):
        """fri.go:314-384"""
        arity = 1 << arity_bits
        assert len(evals) == arity and arity_bits <= 8
        g = gl.primitive_root_of_unity(arity_bits)
        g_inv = pow(g, arity - 1, gl.P)
        permuted = [None] * arity
        for i in range(arity):
            permuted[_bitrev8(i) >> (8 - arity_bits)] = evals[i]
        rev_bits = list(reversed(x_index_within_coset_bits))
        start = self.expFromBitsConstBase(g_inv, rev_bits)
        coset_start = self.gl.Mul(start, x)
        x_points = [None] * arity
        x_points[0] = (coset_start, 0)
        for i in range(1, arity):
            x_points[i] = self.gl.MulExtension(x_points[i - 1], (g, 0))
        weights = [None] * arity
        for i in range(arity):
            w = gl.ONE_QE
            for j in range(arity):
                if i != j:
                    w = self.gl.SubMulExtension(x_points[i], x_points[j], w)
            inv, has_inv = self.gl.InverseExtension(w)
            self.api.AssertIsEqual(has_inv, 1)
            weights[i] = inv
        return self.interpolate(beta, x_points, permuted, weights)

    def verifyQueryRound(self, instance, challenges, precomputed_reduced_eval, initial_merkle_caps, proof,
                         x_index, n, n_log, round_proof):
        """fri.go:386-498"""
        assert_noncanonical_indices_ok(self.friParams)
        x_index = self.gl.Reduce(x_index)
        x_index_bits = self.api.ToBinary(x_index, 64)[0:self.friParams.DegreeBits + self.friParams.Config.RateBits]
        cap_index_bits = x_index_bits[len(x_index_bits) - self.friParams.Config.CapHeight:]
        self.verifyInitialProof(x_index_bits, round_proof.EvalsProofs, initial_merkle_caps, cap_index_bits)
        subgroup_x = self.calculateSubgroupX(x_index_bits, n_log)
        old_eval = self.friCombineInitial(instance, round_proof.EvalsProofs, challenges.FriAlpha,
                                          (subgroup_x, 0), precomputed_reduced_eval)
        for i, arity_bits in enumerate(self.friParams.ReductionArityBits):
            evals = round_proof.Steps[i][0]
            coset_index_bits = x_index_bits[arity_bits:]
            within = x_index_bits[:arity_bits]
            if arity_bits != 4:
                raise ValueError("assuming arity bits is 4")
            leaf = [self.gl.Lookup2(within[0], within[1], *evals[4 * k:4 * k + 4]) for k in range(4)]
            new_eval = self.gl.Lookup2(within[2], within[3], *leaf)
            self.gl.AssertIsEqual(new_eval[0], old_eval[0])
            self.gl.AssertIsEqual(new_eval[1], old_eval[1])
            old_eval = self.computeEvaluation(subgroup_x, within, arity_bits, evals, challenges.FriBetas[i])
            field_evals = []
            for e in evals:
                field_evals += [e[0], e[1]]
            self.verifyMerkleProofToCapWithCapIndex(field_evals, coset_index_bits, cap_index_bits,
                                                    proof.CommitPhaseMerkleCaps[i], round_proof.Steps[i][1])
            for _ in range(arity_bits):
                subgroup_x = self.gl.Mul(subgroup_x, subgroup_x)
            x_index_bits = coset_index_bits
        final_eval = self.finalPolyEval(proof.FinalPoly, (subgroup_x, 0))
        self.gl.AssertIsEqual(old_eval[0], final_eval[0])
        self.gl.AssertIsEqual(old_eval[1], final_eval[1])

    def VerifyFriProof(self, instance, openings, fri_challenges, initial_merkle_caps, fri_proof):
        """fri.go:500-548"""
        validate_fri_proof_shape(fri_proof, instance, self.friParams)
        self.assertLeadingZeros(fri_challenges.FriPowResponse, self.friParams.Config)
        if self.friParams.Config.NumQueryRounds != len(fri_proof.QueryRoundProofs):
            raise ValueError("Number of query rounds does not match config.")
        precomputed = self.fromOpeningsAndAlpha(openings, fri_challenges.FriAlpha)
        n_log = self.friParams.DegreeBits + self.friParams.Config.RateBits
        n = 1 << n_log
        if len(fri_challenges.FriQueryIndices) != len(fri_proof.QueryRoundProofs):
            raise ValueError("Number of query indices should equal number of query round proofs")
        for x_index, round_proof in zip(fri_challenges.FriQueryIndices, fri_proof.QueryRoundProofs):
            self.verifyQueryRound(instance, fri_challenges, precomputed, initial_merkle_caps, fri_proof,
                                  x_index, n, n_log, round_proof)
