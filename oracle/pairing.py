"""ORACLE (test infrastructure only). BN254 optimal-ate pairing and the verification equations of gnark's
Groth16 (with the BSB22 range-check commitment) and of KZG openings, on Python integers.

What it restates
  groth16.Verify(proof, vk, publicWitness)                      /root/reference/benchmark.go:266
  plonk.Verify's KZG checks                                     /root/reference/benchmark.go:176
both of which live in the un-vendored dependencies gnark v0.9.1 / gnark-crypto
v0.12.2-0.20231013160410-1f65e75b6dfb (go.mod:6-7); the reference pins no proof bytes and no pairing
vectors, so this file is pinned by mathematics, and the tests say so:
  * the pairing is checked to be bilinear in both arguments, non-degenerate and of order r
    (tests/test_oracle_pairing.py) - any such map decides the verification equations below identically;
  * the equation checked is the published one (Groth16, EUROCRYPT 2016, with gnark's commitment extension):
        e(Ar, Bs) = e(alpha, beta) . e(sum_i x_i K_i + D, gamma) . e(Krs, delta)
        e(D, G) . e(PoK, GRootSigmaNeg) = 1      (Pedersen proof of knowledge: PoK = sigma D, GRootSigmaNeg = -G/sigma)
    where the public vector x is (1, public inputs..., challenge) and challenge = hash_to_field(D).

Construction (deliberately the plainest one): Fp12 = Fp2[w]/(w^6 - xi), xi = 9 + u, as six Fp2 coefficients with
schoolbook multiplication; G2 points are untwisted as (x w^2, y w^3); the Miller loop runs over 6t+2 in affine
coordinates on the twist with the two Frobenius end-steps; the final exponentiation is ONE plain square-and-multiply by
(p^12 - 1)/r. Nothing here shares code with csrc/.
"""
import hashlib

from .bn254 import P, R, T_BN, Fp2, G1_GEN, G2_GEN, ec_add, ec_mul, ec_neg, ec_on_curve

XI = Fp2(9, 1)
ATE_LOOP = 6 * T_BN + 2
FINAL_EXP = (P**12 - 1) // R
assert (P**12 - 1) % R == 0


def _fp2_pow(a, e):
    r = Fp2(1, 0)
    while e:
        if e & 1:
            r = r * a
        a = a * a
        e >>= 1
    return r


def _conj(a):
    return Fp2(a.a, -a.b)


# Frobenius on twist coordinates: pi(x w^2, y w^3) = (conj(x) xi^((p-1)/3) w^2, conj(y) xi^((p-1)/2) w^3)
_G_X1, _G_Y1 = _fp2_pow(XI, (P - 1) // 3), _fp2_pow(XI, (P - 1) // 2)
_G_X2, _G_Y2 = _fp2_pow(XI, (P * P - 1) // 3), _fp2_pow(XI, (P * P - 1) // 2)

_ZERO, _ONE = Fp2(0, 0), Fp2(1, 0)


def f12_one():
    return [_ONE, _ZERO, _ZERO, _ZERO, _ZERO, _ZERO]


def f12_mul(a, b):
    t = [_ZERO] * 11
    for i in range(6):
        ai = a[i]
        if ai.a == 0 and ai.b == 0:
            continue
        for j in range(6):
            bj = b[j]
            if bj.a == 0 and bj.b == 0:
                continue
            t[i + j] = t[i + j] + ai * bj
    return [t[i] + (t[i + 6] * XI if i < 5 else _ZERO) for i in range(6)]


def f12_pow(a, e):
    r = f12_one()
    while e:
        if e & 1:
            r = f12_mul(r, a)
        a = f12_mul(a, a)
        e >>= 1
    return r


def f12_eq(a, b):
    return all(x == y for x, y in zip(a, b))


def _line(t, lam, p):
    """line through twist point t with twist slope lam, evaluated at the G1 point p (see the module docstring):
    y_P - lam x_P w + (lam x_T - y_T) w^3"""
    xp, yp = p
    return [Fp2(yp, 0), -(lam * xp), _ZERO, lam * t[0] - t[1], _ZERO, _ZERO]


def _dbl_step(t, p):
    x, y = t
    lam = (x * x * 3) * (y + y).inv()
    l = _line(t, lam, p)
    x3 = lam * lam - x - x
    return (x3, lam * (x - x3) - y), l


def _add_step(t, q, p):
    lam = (q[1] - t[1]) * (q[0] - t[0]).inv()
    l = _line(t, lam, p)
    x3 = lam * lam - t[0] - q[0]
    return (x3, lam * (t[0] - x3) - t[1]), l


def miller_loop(p, q):
    """f_{6t+2,Q}(P) times the two Frobenius lines. p in G1 (ints), q in G2 (Fp2 pairs); None = infinity -> 1."""
    if p is None or q is None:
        return f12_one()
    f = f12_one()
    t = q
    for bit in bin(ATE_LOOP)[3:]:
        t, l = _dbl_step(t, p)
        f = f12_mul(f12_mul(f, f), l)
        if bit == "1":
            t, l = _add_step(t, q, p)
            f = f12_mul(f, l)
    q1 = (_conj(q[0]) * _G_X1, _conj(q[1]) * _G_Y1)
    q2n = (q[0] * _G_X2, -(q[1] * _G_Y2))
    t, l = _add_step(t, q1, p)
    f = f12_mul(f, l)
    t, l = _add_step(t, q2n, p)
    return f12_mul(f, l)


def final_exponentiation(f):
    return f12_pow(f, FINAL_EXP)


def pairing(p, q):
    return final_exponentiation(miller_loop(p, q))


def pairing_product_is_one(pairs):
    """prod e(P_i, Q_i) == 1 with one final exponentiation."""
    f = f12_one()
    for p, q in pairs:
        f = f12_mul(f, miller_loop(p, q))
    return f12_eq(final_exponentiation(f), f12_one())


# ---- hash to field: RFC 9380 expand_message_xmd(SHA-256), L = 48, big-endian mod r (gnark-crypto fr.Hash(msg, dst, 1)) ----
def expand_message_xmd(msg: bytes, dst: bytes, length: int) -> bytes:
    ell = (length + 31) // 32
    dst_prime = dst + bytes([len(dst)])
    b0 = hashlib.sha256(bytes(64) + msg + length.to_bytes(2, "big") + b"\0" + dst_prime).digest()
    b = [hashlib.sha256(b0 + b"\x01" + dst_prime).digest()]
    for i in range(2, ell + 1):
        b.append(hashlib.sha256(bytes(x ^ y for x, y in zip(b0, b[-1])) + bytes([i]) + dst_prime).digest())
    return b"".join(b)[:length]


def hash_to_fr(msg: bytes, dst: bytes = b"bsb22-commitment") -> int:
    return int.from_bytes(expand_message_xmd(msg, dst, 48), "big") % R


def g1_raw(p) -> bytes:
    """gnark-crypto G1Affine.RawBytes / Marshal: X | Y big-endian, 64 bytes (infinity: 0x40 then zeros)"""
    if p is None:
        return bytes([0x40]) + bytes(63)
    return p[0].to_bytes(32, "big") + p[1].to_bytes(32, "big")


def g2_raw(q) -> bytes:
    """gnark-crypto G2Affine.RawBytes: X.A1 | X.A0 | Y.A1 | Y.A0 big-endian, 128 bytes"""
    if q is None:
        return bytes([0x40]) + bytes(127)
    return b"".join(v.to_bytes(32, "big") for v in (q[0].b, q[0].a, q[1].b, q[1].a))


def g1_from_raw(b: bytes):
    if b[0] & 0xC0 == 0x40:
        return None
    return (int.from_bytes(b[:32], "big"), int.from_bytes(b[32:64], "big"))


def g2_from_raw(b: bytes):
    if b[0] & 0xC0 == 0x40:
        return None
    v = [int.from_bytes(b[32 * i:32 * i + 32], "big") for i in range(4)]
    return (Fp2(v[1], v[0]), Fp2(v[3], v[2]))


# ---- Groth16 (gnark v0.9.1 backend/groth16/bn254/verify.go, recalled; equations in the module docstring) -----------------
def groth16_verify(vk, proof, public_inputs):
    """vk: dict alpha1, beta2, gamma2, delta2 (points), K (list of G1: ONE, publics..., then one per commitment),
    optional pedersen {"g": G2, "g_root_sigma_neg": G2}. proof: dict Ar, Bs, Krs, commitments [G1], pok G1.
    public_inputs: ints (without the leading 1). Returns (ok, reason)."""
    for name, g in (("Ar", 1), ("Krs", 1), ("Bs", 2)):
        if proof[name] is None or not ec_on_curve(g, proof[name]):
            return False, name + " not on curve"
    # (subgroup check of Bs: [r]Bs == infinity)
    if ec_mul_full(2, proof["Bs"], R) is not None:
        return False, "Bs not in the r-torsion subgroup"
    x = [1] + [v % R for v in public_inputs]
    commitments = proof.get("commitments", [])
    for d in commitments:
        if d is not None and not ec_on_curve(1, d):
            return False, "commitment not on curve"
        x.append(hash_to_fr(g1_raw(d)))
    if len(x) != len(vk["K"]):
        return False, "public witness length %d != len(vk.K) %d" % (len(x), len(vk["K"]))
    ksum = None
    for xi, ki in zip(x, vk["K"]):
        ksum = ec_add(1, ksum, ec_mul(1, ki, xi))
    for d in commitments:
        ksum = ec_add(1, ksum, d)
    if commitments:
        ped = vk["pedersen"]
        if len(commitments) != 1:
            return False, "batched proofs of knowledge are not restated here"
        if not pairing_product_is_one([(commitments[0], ped["g"]), (proof["pok"], ped["g_root_sigma_neg"])]):
            return False, "commitment proof of knowledge fails"
    ok = pairing_product_is_one([(proof["Ar"], proof["Bs"]), (ec_neg(1, vk["alpha1"]), vk["beta2"]),
                                 (ec_neg(1, ksum), vk["gamma2"]), (ec_neg(1, proof["Krs"]), vk["delta2"])])
    return ok, "" if ok else "pairing equation fails"


def ec_mul_full(group, p, k):
    """[k]P without reducing k mod r (subgroup checks)"""
    acc, add = None, p
    while k:
        if k & 1:
            acc = ec_add(group, acc, add)
        add = ec_add(group, add, add)
        k >>= 1
    return acc


def parse_proof_raw(buf: bytes):
    """gnark groth16 Proof.WriteRawTo layout (benchmark.go:272-291 reads its first 256 bytes):
    Ar (64) | Bs (128) | Krs (64) | uint32 BE n | n commitments (64 each) | CommitmentPok (64)"""
    ar, bs, krs = g1_from_raw(buf[0:64]), g2_from_raw(buf[64:192]), g1_from_raw(buf[192:256])
    n = int.from_bytes(buf[256:260], "big")
    cs = [g1_from_raw(buf[260 + 64 * i:324 + 64 * i]) for i in range(n)]
    pok = g1_from_raw(buf[260 + 64 * n:324 + 64 * n])
    assert len(buf) == 324 + 64 * n
    return {"Ar": ar, "Bs": bs, "Krs": krs, "commitments": cs, "pok": pok}


def parse_vk_raw(buf: bytes):
    """gnark groth16 VerifyingKey.WriteRawTo layout (recalled): alpha1 | beta1 | beta2 | gamma2 | delta1 | delta2 |
    uint32 len(K) | K... | uint32 n_commitments | per commitment: uint32 len, uint64 BE indices | pedersen vk (G | GRootSigmaNeg)"""
    o = 0

    def g1():
        nonlocal o
        o += 64
        return g1_from_raw(buf[o - 64:o])

    def g2():
        nonlocal o
        o += 128
        return g2_from_raw(buf[o - 128:o])

    def u32():
        nonlocal o
        o += 4
        return int.from_bytes(buf[o - 4:o], "big")

    vk = {"alpha1": g1(), "beta1": g1(), "beta2": g2(), "gamma2": g2(), "delta1": g1(), "delta2": g2()}
    vk["K"] = [g1() for _ in range(u32())]
    pac = []
    for _ in range(u32()):
        n = u32()
        pac.append([int.from_bytes(buf[o + 8 * i:o + 8 * i + 8], "big") for i in range(n)])
        o += 8 * n
    vk["public_and_commitment_committed"] = pac
    if pac:
        vk["pedersen"] = {"g": g2(), "g_root_sigma_neg": g2()}
    assert o == len(buf), (o, len(buf))
    return vk


# ---- KZG (gnark-crypto ecc/bn254/kzg Verify, recalled): e(C - [y]G1 + [z] H, G2) . e(-H, [tau]G2) = 1 ------------------------
def kzg_verify(commitment, z, y, h, g2_tau):
    lhs = ec_add(1, ec_add(1, commitment, ec_neg(1, ec_mul(1, G1_GEN, y))), ec_mul(1, h, z))
    return pairing_product_is_one([(lhs, G2_GEN), (ec_neg(1, h), g2_tau)])
