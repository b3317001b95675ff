"""ORACLE (test infrastructure only). Plonky2 custom-gate constraint evaluation, restating
plonk/gates/*.go of the reference. EvaluationVars = (localConstants, localWires, publicInputsHash).
"""
import re

from . import goldilocks as gl
from . import poseidon

UNUSED_SELECTOR = (1 << 32) - 1   # gates/types.go:3
D = gl.D


def _alg(wires, start):
    """vars.go:30-42 GetLocalExtAlgebra for Range{start, start+D}"""
    return (wires[start], wires[start + 1])


class NoopGate:
    def EvalUnfiltered(self, api, g, consts, wires, pih):
        return []


class ConstantGate:
    def __init__(self, num_consts):
        self.numConsts = num_consts

    def EvalUnfiltered(self, api, g, consts, wires, pih):
        return [g.SubExtension(consts[i], wires[i]) for i in range(self.numConsts)]


class PublicInputGate:
    def EvalUnfiltered(self, api, g, consts, wires, pih):
        """public_input_gate.go:32-51"""
        return [g.SubExtension(wires[i], (pih[i], 0)) for i in range(4)]


class ArithmeticGate:
    def __init__(self, num_ops):
        self.numOps = num_ops

    def EvalUnfiltered(self, api, g, consts, wires, pih):
        """arithmetic_gate.go:60-84"""
        c0, c1 = consts[0], consts[1]
        out = []
        for i in range(self.numOps):
            m0, m1, addend, output = wires[4 * i], wires[4 * i + 1], wires[4 * i + 2], wires[4 * i + 3]
            computed = g.AddExtension(g.MulExtension(g.MulExtension(m0, m1), c0), g.MulExtension(addend, c1))
            out.append(g.SubExtension(output, computed))
        return out


class ArithmeticExtensionGate:
    def __init__(self, num_ops):
        self.numOps = num_ops

    def EvalUnfiltered(self, api, g, consts, wires, pih):
        """arithmetic_extension_gate.go:59-86"""
        c0, c1 = consts[0], consts[1]
        out = []
        for i in range(self.numOps):
            m0 = _alg(wires, 4 * D * i)
            m1 = _alg(wires, 4 * D * i + D)
            addend = _alg(wires, 4 * D * i + 2 * D)
            output = _alg(wires, 4 * D * i + 3 * D)
            mul = g.MulExtensionAlgebra(m0, m1)
            scaled_mul = g.ScalarMulExtensionAlgebra(c0, mul)
            computed = g.ScalarMulExtensionAlgebra(c1, addend)
            computed = g.AddExtensionAlgebra(computed, scaled_mul)
            diff = g.SubExtensionAlgebra(output, computed)
            out += [diff[0], diff[1]]
        return out


class MultiplicationExtensionGate:
    def __init__(self, num_ops):
        self.numOps = num_ops

    def EvalUnfiltered(self, api, g, consts, wires, pih):
        """multiplication_extension_gate.go:55-76"""
        c0 = consts[0]
        out = []
        for i in range(self.numOps):
            m0 = _alg(wires, 3 * D * i)
            m1 = _alg(wires, 3 * D * i + D)
            output = _alg(wires, 3 * D * i + 2 * D)
            mul = g.MulExtensionAlgebra(m0, m1)
            computed = g.ScalarMulExtensionAlgebra(c0, mul)
            diff = g.SubExtensionAlgebra(output, computed)
            out += [diff[0], diff[1]]
        return out


class BaseSumGate:
    def __init__(self, num_limbs, base):
        self.numLimbs, self.base = num_limbs, base

    def EvalUnfiltered(self, api, g, consts, wires, pih):
        """base_sum_gate.go:66-96"""
        s = wires[0]
        limbs = [wires[1 + i] for i in range(self.numLimbs)]
        computed = g.ReduceWithPowers(limbs, (self.base, 0))
        out = [g.SubExtension(computed, s)]
        for limb in limbs:
            acc = gl.ONE_QE
            for i in range(self.base):
                acc = g.MulExtension(acc, g.SubExtension(limb, (i, 0)))
            out.append(acc)
        return out


class CosetInterpolationGate:
    def __init__(self, subgroup_bits, degree, weights):
        self.subgroupBits, self.degree, self.barycentricWeights = subgroup_bits, degree, weights

    def numPoints(self):
        return 1 << self.subgroupBits

    def numIntermediates(self):
        return (self.numPoints() - 2) // (self.degree - 1)

    def EvalUnfiltered(self, api, g, consts, wires, pih):
        """coset_interpolation_gate.go:151-226"""
        n = self.numPoints()
        start_values = 1
        start_eval_point = start_values + n * D
        start_eval_value = start_eval_point + D
        start_inter = start_eval_value + D
        ni = self.numIntermediates()
        out = []
        shift = wires[0]
        evaluation_point = _alg(wires, start_eval_point)
        shifted_point = _alg(wires, start_inter + D * 2 * ni)
        neg_shift = g.ScalarMulExtension(shift, gl.NEG_ONE)
        tmp = g.ScalarMulExtensionAlgebra(neg_shift, shifted_point)
        tmp = g.AddExtensionAlgebra(tmp, evaluation_point)
        out += [tmp[0], tmp[1]]
        domain = gl.two_adic_subgroup(self.subgroupBits)
        values = [_alg(wires, start_values + i * D) for i in range(n)]
        weights = self.barycentricWeights
        deg = self.degree
        c_eval, c_prod = g.PartialInterpolateExtAlgebra(domain[:deg], values[:deg], weights[:deg], shifted_point,
                                                       gl.ZERO_ALG, gl.ONE_ALG)
        for i in range(ni):
            inter_eval = _alg(wires, start_inter + D * i)
            inter_prod = _alg(wires, start_inter + D * (ni + i))
            d = g.SubExtensionAlgebra(inter_eval, c_eval)
            out += [d[0], d[1]]
            d = g.SubExtensionAlgebra(inter_prod, c_prod)
            out += [d[0], d[1]]
            s = 1 + (deg - 1) * (i + 1)
            e = min(s + deg - 1, n)
            c_eval, c_prod = g.PartialInterpolateExtAlgebra(domain[s:e], values[s:e], weights[s:e], shifted_point,
                                                           inter_eval, inter_prod)
        evaluation_value = _alg(wires, start_eval_value)
        d = g.SubExtensionAlgebra(evaluation_value, c_eval)
        out += [d[0], d[1]]
        return out


class ExponentiationGate:
    def __init__(self, num_power_bits):
        self.numPowerBits = num_power_bits

    def EvalUnfiltered(self, api, g, consts, wires, pih):
        """exponentiation_gate.go:80-128"""
        n = self.numPowerBits
        base = wires[0]
        power_bits = [wires[1 + i] for i in range(n)]
        inter = [wires[2 + n + i] for i in range(n)]
        output = wires[1 + n]
        out = []
        for i in range(n):
            prev = gl.ONE_QE if i == 0 else g.MulExtension(inter[i - 1], inter[i - 1])
            cur_bit = power_bits[n - i - 1]
            tmp = g.MulExtension(cur_bit, gl.ONE_QE)
            tmp = g.SubExtension(tmp, gl.ONE_QE)
            mul_by = g.MulExtension(cur_bit, base)
            mul_by = g.SubExtension(mul_by, tmp)
            diff = g.MulExtension(prev, mul_by)
            diff = g.SubExtension(diff, inter[i])
            out.append(diff)
        out.append(g.SubExtension(output, inter[n - 1]))
        return out


class PoseidonGate:
    W = poseidon.SPONGE_WIDTH
    START_DELTA = 2 * W + 1
    START_FULL_0 = START_DELTA + 4
    START_PARTIAL = START_FULL_0 + (poseidon.HALF_N_FULL_ROUNDS - 1) * W
    START_FULL_1 = START_PARTIAL + poseidon.N_PARTIAL_ROUNDS

    def EvalUnfiltered(self, api, g, consts, wires, pih):
        """poseidon_gate.go:92-181"""
        W = self.W
        pc = poseidon.GoldilocksChip(api)
        out = []
        swap = wires[2 * W]
        out.append(g.MulExtension(swap, g.SubExtension(swap, gl.ONE_QE)))
        for i in range(4):
            lhs, rhs, delta = wires[i], wires[i + 4], wires[self.START_DELTA + i]
            diff = g.SubExtension(rhs, lhs)
            out.append(g.SubExtension(g.MulExtension(swap, diff), delta))
        state = [None] * W
        for i in range(4):
            delta, lhs, rhs = wires[self.START_DELTA + i], wires[i], wires[i + 4]
            state[i] = g.AddExtension(lhs, delta)
            state[i + 4] = g.SubExtension(rhs, delta)
        for i in range(8, W):
            state[i] = wires[i]
        rc = [0]
        for r in range(poseidon.HALF_N_FULL_ROUNDS):
            state = pc.ConstantLayerExtension(state, rc)
            if r != 0:
                for i in range(W):
                    sbox_in = wires[self.START_FULL_0 + (r - 1) * W + i]
                    out.append(g.SubExtension(state[i], sbox_in))
                    state[i] = sbox_in
            state = pc.SBoxLayerExtension(state)
            state = pc.MdsLayerExtension(state)
            rc[0] += 1
        state = pc.PartialFirstConstantLayerExtension(state)
        state = pc.MdsPartialLayerInitExtension(state)
        for r in range(poseidon.N_PARTIAL_ROUNDS - 1):
            sbox_in = wires[self.START_PARTIAL + r]
            out.append(g.SubExtension(state[0], sbox_in))
            state[0] = pc.SBoxMonomialExtension(sbox_in)
            state[0] = g.AddExtension(state[0], (poseidon.FAST_PARTIAL_ROUND_CONSTANTS[r], 0))
            state = pc.MdsPartialLayerFastExtension(state, r)
        sbox_in = wires[self.START_PARTIAL + poseidon.N_PARTIAL_ROUNDS - 1]
        out.append(g.SubExtension(state[0], sbox_in))
        state[0] = pc.SBoxMonomialExtension(sbox_in)
        state = pc.MdsPartialLayerFastExtension(state, poseidon.N_PARTIAL_ROUNDS - 1)
        rc[0] += poseidon.N_PARTIAL_ROUNDS
        for r in range(poseidon.HALF_N_FULL_ROUNDS):
            state = pc.ConstantLayerExtension(state, rc)
            for i in range(W):
                sbox_in = wires[self.START_FULL_1 + r * W + i]
                out.append(g.SubExtension(state[i], sbox_in))
                state[i] = sbox_in
            state = pc.SBoxLayerExtension(state)
            state = pc.MdsLayerExtension(state)
            rc[0] += 1
        for i in range(W):
            out.append(g.SubExtension(state[i], wires[W + i]))
        return out


class PoseidonMdsGate:
    def EvalUnfiltered(self, api, g, consts, wires, pih):
        """poseidon_mds_gate.go:36-99"""
        W = poseidon.SPONGE_WIDTH
        inputs = [_alg(wires, i * D) for i in range(W)]
        out = []
        computed = []
        for r in range(W):
            res = gl.ZERO_ALG
            for i in range(W):
                coeff = (poseidon.MDS_MATRIX_CIRC[i], 0)
                res = g.AddExtensionAlgebra(res, g.ScalarMulExtensionAlgebra(coeff, inputs[(i + r) % W]))
            coeff = (poseidon.MDS_MATRIX_DIAG[r], 0)
            res = g.AddExtensionAlgebra(res, g.ScalarMulExtensionAlgebra(coeff, inputs[r]))
            computed.append(res)
        for i in range(W):
            output = _alg(wires, (W + i) * D)
            diff = g.SubExtensionAlgebra(output, computed[i])
            out += [diff[0], diff[1]]
        return out


class RandomAccessGate:
    def __init__(self, bits, num_copies, num_extra_constants):
        self.bits, self.numCopies, self.numExtraConstants = bits, num_copies, num_extra_constants

    def EvalUnfiltered(self, api, g, consts, wires, pih):
        """random_access_gate.go:131-190"""
        vec = 1 << self.bits
        two = (2, 0)
        out = []
        start_extra = (2 + vec) * self.numCopies
        num_routed = start_extra + self.numExtraConstants
        for copy in range(self.numCopies):
            access_index = wires[(2 + vec) * copy]
            list_items = [wires[(2 + vec) * copy + 2 + i] for i in range(vec)]
            claimed = wires[(2 + vec) * copy + 1]
            bits = [wires[num_routed + copy * self.bits + i] for i in range(self.bits)]
            for b in bits:
                out.append(g.SubExtension(g.MulExtension(b, b), b))
            reconstructed = g.ReduceWithPowers(bits, two)
            out.append(g.SubExtension(reconstructed, access_index))
            for b in bits:
                tmp = []
                for i in range(0, len(list_items), 2):
                    x, y = list_items[i], list_items[i + 1]
                    diff = g.SubExtension(y, x)
                    mul = g.MulExtension(b, diff)
                    tmp.append(g.AddExtension(x, mul))
                list_items = tmp
            assert len(list_items) == 1
            out.append(g.SubExtension(list_items[0], claimed))
        for i in range(self.numExtraConstants):
            out.append(g.SubExtension(consts[i], wires[start_extra + i]))
        return out


class ReducingExtensionGate:
    def __init__(self, num_coeffs):
        self.numCoeffs = num_coeffs

    def _accs(self, i):
        if i == self.numCoeffs - 1:
            return 0
        return 3 * D + self.numCoeffs * D + D * i

    def EvalUnfiltered(self, api, g, consts, wires, pih):
        """reducing_extension_gate.go:77-109"""
        alpha = _alg(wires, D)
        acc = _alg(wires, 2 * D)
        out = []
        for i in range(self.numCoeffs):
            coeff = _alg(wires, 3 * D + D * i)
            acc_i = _alg(wires, self._accs(i))
            tmp = g.MulExtensionAlgebra(acc, alpha)
            tmp = g.AddExtensionAlgebra(tmp, coeff)
            tmp = g.SubExtensionAlgebra(tmp, acc_i)
            out += [tmp[0], tmp[1]]
            acc = acc_i
        return out


class ReducingGate:
    def __init__(self, num_coeffs):
        self.numCoeffs = num_coeffs

    def _accs(self, i):
        if i == self.numCoeffs - 1:
            return 0
        return 3 * D + self.numCoeffs + D * i

    def EvalUnfiltered(self, api, g, consts, wires, pih):
        """reducing_gate.go:77-110"""
        alpha = _alg(wires, D)
        acc = _alg(wires, 2 * D)
        out = []
        for i in range(self.numCoeffs):
            coeff = (wires[3 * D + i], gl.ZERO_QE)
            acc_i = _alg(wires, self._accs(i))
            tmp = g.MulExtensionAlgebra(acc, alpha)
            tmp = g.AddExtensionAlgebra(tmp, coeff)
            tmp = g.SubExtensionAlgebra(tmp, acc_i)
            out += [tmp[0], tmp[1]]
            acc = acc_i
        return out


# gates.go:20-54 -- the regex handlers (mutually exclusive patterns; order irrelevant)
_HANDLERS = [
    (re.compile(r"ArithmeticGate { num_ops: (?P<numOps>[0-9]+) }"), lambda p: ArithmeticGate(int(p["numOps"]))),
    (re.compile(r"ArithmeticExtensionGate { num_ops: (?P<numOps>[0-9]+) }"),
     lambda p: ArithmeticExtensionGate(int(p["numOps"]))),
    (re.compile(r"BaseSumGate { num_limbs: (?P<numLimbs>[0-9]+) } \+ Base: (?P<base>[0-9]+)"),
     lambda p: BaseSumGate(int(p["numLimbs"]), int(p["base"]))),
    (re.compile(r"ConstantGate { num_consts: (?P<numConsts>[0-9]+) }"), lambda p: ConstantGate(int(p["numConsts"]))),
    (re.compile(r"CosetInterpolationGate { subgroup_bits: (?P<subgroupBits>[0-9]+), degree: (?P<degree>[0-9]+), "
                r"barycentric_weights: \[(?P<barycentricWeights>[0-9, ]+)\], _phantom: "
                r"PhantomData<plonky2_field::goldilocks_field::GoldilocksField> }<D=2>"),
     lambda p: CosetInterpolationGate(int(p["subgroupBits"]), int(p["degree"]),
                                      [int(x.strip()) for x in p["barycentricWeights"].split(",")])),
    (re.compile(r"ExponentiationGate { num_power_bits: (?P<numPowerBits>[0-9]+), _phantom: "
                r"PhantomData<plonky2_field::goldilocks_field::GoldilocksField> }<D=(?P<base>[0-9]+)>"),
     lambda p: ExponentiationGate(int(p["numPowerBits"]))),
    (re.compile(r"MulExtensionGate { num_ops: (?P<numOps>[0-9]+) }"),
     lambda p: MultiplicationExtensionGate(int(p["numOps"]))),
    (re.compile(r"NoopGate"), lambda p: NoopGate()),
    (re.compile(r"PoseidonGate.*"), lambda p: PoseidonGate()),
    (re.compile(r"PoseidonMdsGate.*"), lambda p: PoseidonMdsGate()),
    (re.compile(r"PublicInputGate"), lambda p: PublicInputGate()),
    (re.compile(r"RandomAccessGate { bits: (?P<bits>[0-9]+), num_copies: (?P<numCopies>[0-9]+), "
                r"num_extra_constants: (?P<numExtraConstants>[0-9]+), _phantom: "
                r"PhantomData<plonky2_field::goldilocks_field::GoldilocksField> }<D=(?P<base>[0-9]+)>"),
     lambda p: RandomAccessGate(int(p["bits"]), int(p["numCopies"]), int(p["numExtraConstants"]))),
    (re.compile(r"ReducingExtensionGate { num_coeffs: (?P<numCoeffs>[0-9]+) }"),
     lambda p: ReducingExtensionGate(int(p["numCoeffs"]))),
    (re.compile(r"ReducingGate { num_coeffs: (?P<numCoeffs>[0-9]+) }"), lambda p: ReducingGate(int(p["numCoeffs"]))),
]


def GateInstanceFromId(gate_id):
    for rx, handler in _HANDLERS:
        m = rx.search(gate_id)
        if m:
            return handler(m.groupdict())
    raise ValueError("Unknown gate ID %s" % gate_id)


class EvaluateGatesChip:
    """evaluate_gates.go:8-105"""

    def __init__(self, api, gates, num_gate_constraints, selector_indices, groups):
        self.api = api
        self.gates = gates
        self.numGateConstraints = num_gate_constraints
        self.selectorIndices = selector_indices
        self.groups = groups

    def computeFilter(self, row, group, s, many_selector):
        g = gl.Chip(self.api)
        product = gl.ONE_QE
        for i in range(group[0], group[1]):
            if i == row:
                continue
            product = g.MulExtension(product, g.SubExtension((i, 0), s))
        if many_selector:
            product = g.MulExtension(product, g.SubExtension((UNUSED_SELECTOR, 0), s))
        return product

    def evalFiltered(self, gate, consts, wires, pih, row, selector_index, group, num_selectors):
        g = gl.Chip(self.api)
        filt = self.computeFilter(row, group, consts[selector_index], num_selectors > 1)
        unfiltered = gate.EvalUnfiltered(self.api, g, consts[num_selectors:], wires, pih)
        return [g.MulExtension(u, filt) for u in unfiltered]

    def EvaluateGateConstraints(self, consts, wires, pih):
        g = gl.Chip(self.api)
        constraints = [gl.ZERO_QE] * self.numGateConstraints
        num_selectors = len(self.groups)
        for i, gate in enumerate(self.gates):
            sel = self.selectorIndices[i]
            gate_constraints = self.evalFiltered(gate, consts, wires, pih, i, sel, self.groups[sel], num_selectors)
            for k, c in enumerate(gate_constraints):
                if k >= self.numGateConstraints:
                    raise ValueError("num_constraints() gave too low of a number")
                constraints[k] = g.AddExtension(constraints[k], c)
        return constraints
