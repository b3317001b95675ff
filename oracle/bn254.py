"""ORACLE (test infrastructure only). BN254 fields, G1/G2 group law, textbook MSM and NTT on Python
integers.

PARITY NOTE: MultiExp / FFT live in the un-vendored dependency gnark-crypto
v0.12.2-0.20231013160410-1f65e75b6dfb (go.mod:7) and the reference pins no MSM / NTT / proof-byte
vectors, so this part of the oracle is "parity unpinned" by reference goldens. It is
pinned instead by mathematics: an MSM result is a unique group element (schoolbook double-and-add
here), an NTT result is unique given the domain conventions of gnark-crypto's fr/fft (2-adicity 28,
generator 5, w_N = 5^((r-1)/N)), checked by the O(n^2) DFT definition, and the curve / field
constants are re-derived numerically in tests/test_oracle_bn254.py.
"""
R = 21888242871839275222246405745257275088548364400416034343698204186575808495617
P = 21888242871839275222246405745257275088696311157297823662689037894645226208583
T_BN = 4965661367192848881
assert P == 36 * T_BN**4 + 36 * T_BN**3 + 24 * T_BN**2 + 6 * T_BN + 1
assert R == 36 * T_BN**4 + 36 * T_BN**3 + 18 * T_BN**2 + 6 * T_BN + 1

FR_GENERATOR = 5
FR_TWO_ADICITY = 28
ROOT_2_28 = pow(FR_GENERATOR, (R - 1) >> FR_TWO_ADICITY, R)
assert ROOT_2_28 == 19103219067921713944291392827692070036145651957329286315305642004821462161904


def root_of_unity(logn):
    return pow(ROOT_2_28, 1 << (FR_TWO_ADICITY - logn), R)


# ---- Fp2 = Fp[u]/(u^2+1) ----------------------------------------------------------------------------
class Fp2:
    __slots__ = ("a", "b")

    def __init__(self, a, b=0):
        self.a, self.b = a % P, b % P

    def __add__(self, o):
        return Fp2(self.a + o.a, self.b + o.b)

    def __sub__(self, o):
        return Fp2(self.a - o.a, self.b - o.b)

    def __neg__(self):
        return Fp2(-self.a, -self.b)

    def __mul__(self, o):
        if isinstance(o, int):
            return Fp2(self.a * o, self.b * o)
        return Fp2(self.a * o.a - self.b * o.b, self.a * o.b + self.b * o.a)

    def inv(self):
        n = pow(self.a * self.a + self.b * self.b, P - 2, P)
        return Fp2(self.a * n, -self.b * n)

    def __eq__(self, o):
        return self.a == o.a and self.b == o.b

    def is_zero(self):
        return self.a == 0 and self.b == 0

    def tup(self):
        return (self.a, self.b)


class _FpOps:
    zero = 0
    @staticmethod
    def add(a, b): return (a + b) % P
    @staticmethod
    def sub(a, b): return (a - b) % P
    @staticmethod
    def mul(a, b): return a * b % P
    @staticmethod
    def inv(a): return pow(a, P - 2, P)
    @staticmethod
    def neg(a): return (-a) % P
    @staticmethod
    def is_zero(a): return a % P == 0


class _Fp2Ops:
    zero = Fp2(0, 0)
    @staticmethod
    def add(a, b): return a + b
    @staticmethod
    def sub(a, b): return a - b
    @staticmethod
    def mul(a, b): return a * b
    @staticmethod
    def inv(a): return a.inv()
    @staticmethod
    def neg(a): return -a
    @staticmethod
    def is_zero(a): return a.is_zero()


G1_GEN = (1, 2)
G2_GEN = (Fp2(10857046999023057135944570762232829481370756359578518086990519993285655852781,
              11559732032986387107991004021392285783925812861821192530917403151452391805634),
          Fp2(8495653923123431417604973247489272438418190587263600148770280649306958101930,
              4082367875863433681332203403145435568316851327593401208105741076214120093531))
G1_B = 3
G2_B = Fp2(3, 0) * Fp2(9, 1).inv()


def _ops(group):
    return _FpOps if group == 1 else _Fp2Ops


def ec_add(group, p, q):
    """Affine addition; None = infinity."""
    F = _ops(group)
    if p is None:
        return q
    if q is None:
        return p
    x1, y1 = p
    x2, y2 = q
    if F.is_zero(F.sub(x1, x2)):
        if F.is_zero(F.add(y1, y2)):
            return None
        three_x2 = F.mul(F.mul(x1, x1), 3) if group == 1 else F.mul(x1, x1) * 3
        lam = F.mul(three_x2, F.inv(F.add(y1, y1)))
    else:
        lam = F.mul(F.sub(y2, y1), F.inv(F.sub(x2, x1)))
    x3 = F.sub(F.sub(F.mul(lam, lam), x1), x2)
    y3 = F.sub(F.mul(lam, F.sub(x1, x3)), y1)
    return (x3, y3)


def ec_neg(group, p):
    if p is None:
        return None
    return (p[0], _ops(group).neg(p[1]))


def ec_mul(group, p, k):
    k %= R
    acc = None
    add = p
    while k:
        if k & 1:
            acc = ec_add(group, acc, add)
        add = ec_add(group, add, add)
        k >>= 1
    return acc


def ec_on_curve(group, p):
    if p is None:
        return True
    F = _ops(group)
    x, y = p
    b = G1_B if group == 1 else G2_B
    return F.is_zero(F.sub(F.mul(y, y), F.add(F.mul(F.mul(x, x), x), b)))


def msm_naive(group, scalars, points):
    """sum_i [s_i] P_i by schoolbook double-and-add (the defining computation)."""
    acc = None
    for s, p in zip(scalars, points):
        acc = ec_add(group, acc, ec_mul(group, p, s))
    return acc


def point_key(group, p):
    """Canonical comparable form: G1 (x,y); G2 ((x0,x1),(y0,y1)); infinity -> zeros."""
    if p is None:
        return (0, 0) if group == 1 else ((0, 0), (0, 0))
    if group == 1:
        return (p[0] % P, p[1] % P)
    return (p[0].tup(), p[1].tup())


# ---- NTT ---------------------------------------------------------------------------------------------
def dft_naive(a, inverse=False, coset=False):
    """O(n^2) definition, natural order in/out, gnark-crypto conventions."""
    n = len(a)
    logn = n.bit_length() - 1
    w = root_of_unity(logn)
    if not inverse:
        if coset:
            a = [x * pow(FR_GENERATOR, j, R) % R for j, x in enumerate(a)]
        return [sum(a[j] * pow(w, j * k, R) for j in range(n)) % R for k in range(n)]
    winv = pow(w, R - 2, R)
    ninv = pow(n, R - 2, R)
    out = [sum(a[k] * pow(winv, j * k, R) for k in range(n)) * ninv % R for j in range(n)]
    if coset:
        ginv = pow(FR_GENERATOR, R - 2, R)
        out = [x * pow(ginv, j, R) % R for j, x in enumerate(out)]
    return out


def ntt_fast(a, inverse=False, coset=False):
    """Iterative radix-2 (natural in/out) for medium sizes; same conventions as dft_naive."""
    n = len(a)
    logn = n.bit_length() - 1
    a = [x % R for x in a]
    if not inverse and coset:
        g = 1
        for j in range(n):
            a[j] = a[j] * g % R
            g = g * FR_GENERATOR % R
    w = root_of_unity(logn)
    if inverse:
        w = pow(w, R - 2, R)
    # bit-reverse then DIT
    rev = [int("{:0{}b}".format(i, logn)[::-1], 2) if logn else 0 for i in range(n)]
    a = [a[rev[i]] for i in range(n)]
    size = 2
    while size <= n:
        wm = pow(w, n // size, R)
        for start in range(0, n, size):
            t = 1
            for j in range(size // 2):
                u = a[start + j]
                v = a[start + j + size // 2] * t % R
                a[start + j] = (u + v) % R
                a[start + j + size // 2] = (u - v) % R
                t = t * wm % R
        size *= 2
    if inverse:
        ninv = pow(n, R - 2, R)
        a = [x * ninv % R for x in a]
        if coset:
            ginv = pow(FR_GENERATOR, R - 2, R)
            g = 1
            for j in range(n):
                a[j] = a[j] * g % R
                g = g * ginv % R
    return a


def bit_reverse_list(a):
    n = len(a)
    logn = n.bit_length() - 1
    return [a[int("{:0{}b}".format(i, logn)[::-1], 2) if logn else 0] for i in range(n)]
