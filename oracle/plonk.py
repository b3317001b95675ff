"""ORACLE (test infrastructure only). Plonk vanishing-polynomial check, restating plonk/plonk.go."""
from . import goldilocks as gl
from . import gates


class PlonkChip:
    def __init__(self, api, common_data):
        """plonk.go:27-53"""
        self.api = api
        self.commonData = common_data
        created = [gates.GateInstanceFromId(gid) for gid in common_data.GateIds]
        self.evaluateGatesChip = gates.EvaluateGatesChip(api, created, common_data.NumGateConstraints,
                                                         common_data.SelectorIndices, common_data.SelectorGroups)
        self.DEGREE = 1 << common_data.DegreeBits
        self.DEGREE_QE = (1 << common_data.DegreeBits, 0)
        self.commonDataKIs = list(common_data.KIs)

    def expPowerOf2Extension(self, x):
        g = gl.Chip(self.api)
        for _ in range(self.commonData.DegreeBits):
            x = g.MulExtension(x, x)
        return x

    def evalL0(self, x, x_pow_n):
        """plonk.go:63-83"""
        g = gl.Chip(self.api)
        eval_zero_poly = g.SubExtension(x_pow_n, gl.ONE_QE)
        denominator = g.SubExtension(g.ScalarMulExtension(x, self.DEGREE), self.DEGREE_QE)
        quotient, has_quotient = g.DivExtension(eval_zero_poly, denominator)
        self.api.AssertIsEqual(has_quotient, 1)
        return quotient

    def checkPartialProducts(self, numerators, denominators, challenge_num, openings):
        """plonk.go:85-119"""
        g = gl.Chip(self.api)
        npp = self.commonData.NumPartialProducts
        qdf = self.commonData.QuotientDegreeFactor
        accs = [openings.PlonkZs[challenge_num]]
        accs += openings.PartialProducts[challenge_num * npp:(challenge_num + 1) * npp]
        accs.append(openings.PlonkZsNext[challenge_num])
        checks = []
        for i in range(npp + 1):
            s = i * qdf
            nume, deno = numerators[s], denominators[s]
            for j in range(1, qdf):
                nume = g.MulExtension(nume, numerators[s + j])
                deno = g.MulExtension(deno, denominators[s + j])
            checks.append(g.SubExtension(g.MulExtension(accs[i], nume), g.MulExtension(accs[i + 1], deno)))
        return checks

    def evalVanishingPoly(self, consts, wires, pih, ch, openings, zeta_pow_n):
        """plonk.go:121-207"""
        g = gl.Chip(self.api)
        cd = self.commonData
        constraint_terms = self.evaluateGatesChip.EvaluateGateConstraints(consts, wires, pih)
        s_ids = [g.ScalarMulExtension(ch.PlonkZeta, self.commonDataKIs[i]) for i in range(cd.NumRoutedWires)]
        l0_zeta = self.evalL0(ch.PlonkZeta, zeta_pow_n)
        z1_terms, pp_terms = [], []
        for i in range(cd.NumChallenges):
            z1_terms.append(g.MulExtension(l0_zeta, g.SubExtension(openings.PlonkZs[i], gl.ONE_QE)))
            nums, dens = [], []
            for j in range(cd.NumRoutedWires):
                wire_plus_gamma = g.AddExtension(openings.Wires[j], (ch.PlonkGammas[i], 0))
                nums.append(g.AddExtension(g.MulExtension((ch.PlonkBetas[i], 0), s_ids[j]), wire_plus_gamma))
                dens.append(g.AddExtension(g.MulExtension((ch.PlonkBetas[i], 0), openings.PlonkSigmas[j]),
                                           wire_plus_gamma))
            pp_terms += self.checkPartialProducts(nums, dens, i, openings)
        terms = z1_terms + pp_terms + constraint_terms
        reduced = [gl.ZERO_QE] * cd.NumChallenges
        for t in reversed(terms):
            for j in range(cd.NumChallenges):
                reduced[j] = g.AddExtension(t, g.ScalarMulExtension(reduced[j], ch.PlonkAlphas[j]))
        return reduced

    def Verify(self, ch, openings, pih):
        """plonk.go:209-250"""
        g = gl.Chip(self.api)
        zeta_pow_n = self.expPowerOf2Extension(ch.PlonkZeta)
        vanishing = self.evalVanishingPoly(openings.Constants, openings.Wires, pih, ch, openings, zeta_pow_n)
        z_h_zeta = g.SubExtension(zeta_pow_n, gl.ONE_QE)
        qdf = self.commonData.QuotientDegreeFactor
        for i, v in enumerate(vanishing):
            chunk = openings.QuotientPolys[i * qdf:(i + 1) * qdf]
            prod = g.MulExtension(z_h_zeta, g.ReduceWithPowers(chunk, zeta_pow_n))
            g.AssertIsEqualExtension(v, prod)
