"""ORACLE (test infrastructure only). Goldilocks base field, quadratic extension and extension
algebra, restating goldilocks/base.go, goldilocks/quadratic_extension.go and
goldilocks/quadratic_extension_algebra.go of the reference on plain integers.

A gl.Variable is a Python int (the integer held by the Fr wire; "NoReduce" values may exceed P).
A QuadraticExtensionVariable is a 2-tuple, an algebra variable a 2-tuple of 2-tuples.
"""
from .engine import R, AssertionFailed

P = (1 << 64) - (1 << 32) + 1                    # goldilocks/base.go:42 MODULUS
MULTIPLICATIVE_GROUP_GENERATOR = 7               # base.go:33
TWO_ADICITY = 32                                 # base.go:36
POWER_OF_TWO_GENERATOR = 1753635133440165772     # base.go:39
RANGE_CHECK_NB_BITS = 144                        # base.go:48
W = 7                                            # quadratic_extension.go:9
DTH_ROOT = 18446744069414584320                  # quadratic_extension.go:10
NEG_ONE = P - 1                                  # base.go:83
D = 2                                            # quadratic_extension_algebra.go:5

ZERO_QE = (0, 0)
ONE_QE = (1, 0)
ZERO_ALG = (ZERO_QE, ZERO_QE)
ONE_ALG = (ONE_QE, ZERO_QE)


# ---- the four hints, as pure functions (base.go:223-359) ----------------------------------
def mul_add_hint(a, b, c):
    """base.go:223-243. Panics (ValueError) on non-canonical inputs like the reference."""
    for x in (a, b, c):
        if x >= P:
            raise ValueError("%d is not in the field" % x)
    s = a * b + c
    return s // P, s % P


def reduce_hint(x):
    """base.go:284-294 (accepts anything)."""
    return x // P, x % P


def inverse_hint(x):
    """base.go:316-336; gnark-crypto goldilocks Element.Inverse maps 0 -> 0."""
    if x >= P:
        raise ValueError("Input is not in the field")
    return pow(x, P - 2, P) if x else 0


def split_limbs_hint(x):
    """base.go:339-359 (returns an error for x >= P)."""
    if x >= P:
        raise ValueError("input is not in the field")
    return x >> 32, x & 0xFFFFFFFF


def primitive_root_of_unity(n_log):
    """base.go:445-454"""
    assert n_log <= TWO_ADICITY
    res = POWER_OF_TWO_GENERATOR
    for _ in range(TWO_ADICITY - n_log):
        res = res * res % P
    return res


def two_adic_subgroup(n_log):
    """base.go:456-471"""
    g = primitive_root_of_unity(n_log)
    res = [1]
    for _ in range((1 << n_log) - 1):
        res.append(res[-1] * g % P)
    return res


class Chip:
    """goldilocks.Chip (base.go:96-110). Range checks follow the COMMIT_RANGE_CHECKER branch of
    rangeCheckerCheck (base.go:411-421): requests are collected in order."""

    def __init__(self, api):
        self.api = api

    # ---- base field (base.go:160-213) ----------------------------------------------------
    def Add(self, a, b):
        return self.MulAdd(a, 1, b)

    def AddNoReduce(self, a, b):
        return self.api.Add(a, b)

    def Sub(self, a, b):
        return self.MulAdd(b, NEG_ONE, a)

    def SubNoReduce(self, a, b):
        return self.api.Add(a, self.api.Mul(b, NEG_ONE))

    def Mul(self, a, b):
        return self.MulAdd(a, b, 0)

    def MulNoReduce(self, a, b):
        return self.api.Mul(a, b)

    def MulAdd(self, a, b, c):
        """base.go:196-213"""
        q, r = mul_add_hint(a, b, c)
        self.api.record_hint("muladd", (a, b, c), (q, r))
        lhs = self.api.MulAcc(self.api.Mul(c, 1), a, b)
        rhs = self.api.MulAcc(r, P, q)
        self.api.AssertIsEqual(lhs, rhs)
        self.RangeCheck(q)
        self.RangeCheck(r)
        return r

    def MulAddNoReduce(self, a, b, c):
        return self.api.MulAcc(self.api.Mul(c, 1), a, b)

    def Reduce(self, x):
        return self.ReduceWithMaxBits(x, RANGE_CHECK_NB_BITS)

    def ReduceWithMaxBits(self, x, max_nb_bits):
        """base.go:259-281"""
        q, r = reduce_hint(x)
        self.api.record_hint("reduce", (x,), (q, r))
        self.rangeCheckerCheck(q, max_nb_bits)
        self.RangeCheck(r)
        self.api.AssertIsEqual(x, self.api.Add(self.api.Mul(q, P), r))
        return r

    def Inverse(self, x):
        """base.go:297-313 -> (inverse, hasInv)"""
        inv = inverse_hint(x)
        self.api.record_hint("inverse", (x,), (inv,))
        is_zero = self.api.IsZero(x)
        has_inv = self.api.Sub(1, is_zero)
        self.RangeCheck(inv)
        product = self.Mul(inv, x)
        to_check = self.api.Select(has_inv, product, 1)
        self.api.AssertIsEqual(to_check, 1)
        return inv, has_inv

    def RangeCheck(self, x):
        """base.go:362-400"""
        hi, lo = split_limbs_hint(x)
        self.api.record_hint("split", (x,), (hi, lo))
        self.api.AssertIsEqual(self.api.Add(self.api.Mul(hi, 1 << 32), lo), x)
        self.rangeCheckerCheck(hi, 32)
        self.rangeCheckerCheck(lo, 32)
        should_check = self.api.IsZero(self.api.Sub(hi, (1 << 32) - 1))
        self.api.AssertIsEqual(self.api.Select(should_check, lo, 0), 0)

    def RangeCheckWithMaxBits(self, x, max_nb_bits):
        self.rangeCheckerCheck(x, max_nb_bits)

    def AssertIsEqual(self, x, y):
        self.api.AssertIsEqual(x, y)

    def rangeCheckerCheck(self, x, nb_bits):
        self.api.record_range_check(x, nb_bits)

    # ---- quadratic extension (quadratic_extension.go) ---------------------------------------
    def AddExtension(self, a, b):
        return (self.Add(a[0], b[0]), self.Add(a[1], b[1]))

    def AddExtensionNoReduce(self, a, b):
        return (self.AddNoReduce(a[0], b[0]), self.AddNoReduce(a[1], b[1]))

    def SubExtension(self, a, b):
        return (self.Sub(a[0], b[0]), self.Sub(a[1], b[1]))

    def SubExtensionNoReduce(self, a, b):
        return (self.SubNoReduce(a[0], b[0]), self.SubNoReduce(a[1], b[1]))

    def MulExtension(self, a, b):
        return self.ReduceExtension(self.MulExtensionNoReduce(a, b))

    def MulExtensionNoReduce(self, a, b):
        """quadratic_extension.go:65-71"""
        c0o0 = self.MulNoReduce(a[0], b[0])
        c0o1 = self.MulNoReduce(self.MulNoReduce(W, a[1]), b[1])
        c0 = self.AddNoReduce(c0o0, c0o1)
        c1 = self.AddNoReduce(self.MulNoReduce(a[0], b[1]), self.MulNoReduce(a[1], b[0]))
        return (c0, c1)

    def MulAddExtension(self, a, b, c):
        return self.ReduceExtension(self.AddExtensionNoReduce(self.MulExtensionNoReduce(a, b), c))

    def MulAddExtensionNoReduce(self, a, b, c):
        return self.AddExtensionNoReduce(self.MulExtensionNoReduce(a, b), c)

    def SubMulExtension(self, a, b, c):
        return self.ReduceExtension(self.MulExtensionNoReduce(self.SubExtensionNoReduce(a, b), c))

    def ScalarMulExtension(self, a, b):
        return (self.Mul(a[0], b), self.Mul(a[1], b))

    def InnerProductExtension(self, constant, starting_acc, pairs):
        """quadratic_extension.go:107-120"""
        acc = starting_acc
        for a, b in pairs:
            mul = self.ScalarMulExtension(a, constant)
            acc = self.MulAddExtensionNoReduce(mul, b, acc)
        return self.ReduceExtension(acc)

    def InverseExtension(self, a):
        """quadratic_extension.go:123-134"""
        a_is_zero = self.IsZero(a)
        self.api.AssertIsEqual(a_is_zero, 0)
        a_pow_r_minus_1 = (a[0], self.Mul(a[1], DTH_ROOT))
        a_pow_r = self.MulExtension(a_pow_r_minus_1, a)
        inv, has_inv = self.Inverse(a_pow_r[0])
        return self.ScalarMulExtension(a_pow_r_minus_1, inv), has_inv

    def DivExtension(self, a, b):
        b_inv, has_inv = self.InverseExtension(b)
        return self.MulExtension(a, b_inv), has_inv

    def ExpExtension(self, a, exponent):
        """quadratic_extension.go:143-170"""
        if exponent == 0:
            return ONE_QE
        if exponent == 1:
            return a
        if exponent == 2:
            return self.MulExtension(a, a)
        current = a
        product = ONE_QE
        for i in range(exponent.bit_length()):
            if i != 0:
                current = self.MulExtension(current, current)
            if (exponent >> i) & 1:
                product = self.MulExtension(product, current)
        return product

    def ReduceExtension(self, x):
        return (self.Reduce(x[0]), self.Reduce(x[1]))

    def ReduceWithPowers(self, terms, scalar):
        """quadratic_extension.go:177-193"""
        s = ZERO_QE
        for t in reversed(terms):
            s = self.AddExtensionNoReduce(self.MulExtensionNoReduce(s, scalar), t)
            s = self.ReduceExtension(s)
        return s

    def IsZero(self, x):
        return self.api.Mul(self.api.IsZero(x[0]), self.api.IsZero(x[1]))

    def Lookup(self, b, x, y):
        return (self.api.Select(b, y[0], x[0]), self.api.Select(b, y[1], x[1]))

    def Lookup2(self, b0, b1, q0, q1, q2, q3):
        c0 = self.Lookup(b0, q0, q1)
        c1 = self.Lookup(b0, q2, q3)
        return self.Lookup(b1, c0, c1)

    def AssertIsEqualExtension(self, a, b):
        self.AssertIsEqual(a[0], b[0])
        self.AssertIsEqual(a[1], b[1])

    def RangeCheckQE(self, a):
        self.RangeCheck(a[0])
        self.RangeCheck(a[1])

    # ---- extension algebra (quadratic_extension_algebra.go) -----------------------------------
    def AddExtensionAlgebra(self, a, b):
        return tuple(self.AddExtension(a[i], b[i]) for i in range(D))

    def SubExtensionAlgebra(self, a, b):
        return tuple(self.SubExtension(a[i], b[i]) for i in range(D))

    def MulExtensionAlgebra(self, a, b):
        """quadratic_extension_algebra.go:50-75"""
        inner = [[] for _ in range(D)]
        inner_w = [[] for _ in range(D)]
        for i in range(D):
            for j in range(D - i):
                inner[(i + j) % D].append((a[i], b[j]))
            for j in range(D - i, D):
                inner_w[(i + j) % D].append((a[i], b[j]))
        product = []
        for i in range(D):
            acc = self.InnerProductExtension(W, ZERO_QE, inner_w[i])
            product.append(self.InnerProductExtension(1, acc, inner[i]))
        return tuple(product)

    def ScalarMulExtensionAlgebra(self, a, b):
        return tuple(self.MulExtension(a, b[i]) for i in range(D))

    def PartialInterpolateExtAlgebra(self, domain, values, weights, point, initial_eval, initial_prod):
        """quadratic_extension_algebra.go:88-125"""
        n = len(values)
        assert n and n == len(domain) == len(weights)
        new_eval, new_prod = initial_eval, initial_prod
        for i in range(n):
            x_alg = ((domain[i], 0), ZERO_QE)
            weight = (weights[i], 0)
            term = self.SubExtensionAlgebra(point, x_alg)
            weighted_val = self.ScalarMulExtensionAlgebra(weight, values[i])
            new_eval = self.MulExtensionAlgebra(new_eval, term)
            tmp = self.MulExtensionAlgebra(weighted_val, new_prod)
            new_eval = self.AddExtensionAlgebra(new_eval, tmp)
            new_prod = self.MulExtensionAlgebra(new_prod, term)
        return new_eval, new_prod
