"""ORACLE (test infrastructure only - never imported by the product path).

Value-level execution engine for the reference's gadget dataflow: every `frontend.Variable` is a
Python int in BN254 Fr, exactly what gnark's test engine (`test.IsSolved`) computes.
It additionally records the ordered trace of the four reference-owned solver hints
(goldilocks/base.go:223 MulAddHint, :284 ReduceHint, :316 InverseHint, :339 SplitLimbsHint) and
the range-check requests collected by goldilocks/base.go:411-421, which the CUDA witness path is
compared against.
"""

R = 21888242871839275222246405745257275088548364400416034343698204186575808495617  # BN254 Fr


class AssertionFailed(Exception):
    pass


class Api:
    """frontend.API restricted to the calls the reference makes (value semantics, mod R)."""

    def __init__(self, trace=True):
        self.trace_on = trace
        self.hints = []          # (kind, inputs tuple, outputs tuple)
        self.range_checks = []   # (value, bits)   -- rangeCheckerCheck requests, in order
        self.n_asserts = 0
        self.counts = {"muladd": 0, "reduce": 0, "split": 0, "inverse": 0,
                       "poseidon_gl": 0, "poseidon_bn254": 0}

    # -- arithmetic -------------------------------------------------------------------------
    def Add(self, a, b, *rest):
        s = (a + b) % R
        for x in rest:
            s = (s + x) % R
        return s

    def Sub(self, a, b):
        return (a - b) % R

    def Mul(self, a, b):
        return (a * b) % R

    def MulAcc(self, a, b, c):
        return (a + b * c) % R

    def IsZero(self, a):
        return 1 if a % R == 0 else 0

    def Select(self, b, i1, i2):
        if b not in (0, 1):
            raise AssertionFailed("Select: condition not boolean")
        return i1 if b == 1 else i2

    def Lookup2(self, b0, b1, i0, i1, i2, i3):
        if b0 not in (0, 1) or b1 not in (0, 1):
            raise AssertionFailed("Lookup2: bits not boolean")
        return (i0, i1, i2, i3)[b0 + 2 * b1]

    def ToBinary(self, v, n=254):
        v %= R
        if v >> n:
            raise AssertionFailed("ToBinary: value does not fit in %d bits" % n)
        return [(v >> i) & 1 for i in range(n)]

    def FromBinary(self, bits):
        return sum(b << i for i, b in enumerate(bits)) % R

    def AssertIsEqual(self, a, b):
        self.n_asserts += 1
        if (a - b) % R != 0:
            raise AssertionFailed("AssertIsEqual failed: %d != %d" % (a % R, b % R))

    # -- trace ------------------------------------------------------------------------------
    def record_hint(self, kind, inputs, outputs):
        self.counts[kind] += 1
        if self.trace_on:
            self.hints.append((kind, tuple(inputs), tuple(outputs)))

    def record_range_check(self, v, bits):
        if v >> bits:
            raise AssertionFailed("range check failed: %d does not fit in %d bits" % (v, bits))
        if self.trace_on:
            self.range_checks.append((v, bits))
