"""ORACLE (test infrastructure only). Verifier for the PLONK / KZG proofs of csrc/plonk.cu, on Python integers.

Restates plonk.Verify(proof, vk, publicWitness) (/root/reference/benchmark.go:176; gnark v0.9.1 backend/plonk/bn254, un-vendored)
at the level of the published protocol (Gabizon, Williamson, Ciobotaru 2019, section 8) with gnark's BSB22 commitment column:
the proofs checked here are this library's (every polynomial opened at zeta, no linearisation polynomial, own transcript
labels - see the header of csrc/plonk.cu), so "parity" means: the proof convinces an independent implementation of the
verifier's equations. Shares nothing with csrc/: big-int field arithmetic, oracle/pairing.py for the two pairings,
hashlib for the transcript.

  vk bytes:    u32 logN | u32 n_public_rows | u32 has_commit | k1 | k2 | omega (32 B BE) | [qL] [qR] [qM] [qO] [qC] [Qcp] [S1] [S2] [S3]
               (64 B raw each) | [tau]2 (128 B raw)
  proof bytes: [a] [b] [c] [P2] [Z] [t0] [t1] [t2] [W_zeta] [W_zeta_w] (64 B raw each) | a b c z p2 qL qR qM qO qC Qcp S1 S2 S3 t0 t1 t2
               at zeta, z at zeta w (32 B BE each)
"""
import hashlib

from .bn254 import R, G1_GEN, G2_GEN, ec_add, ec_mul, ec_neg, ec_on_curve, root_of_unity
from .pairing import g1_from_raw, g2_from_raw, hash_to_fr, pairing_product_is_one

N_COM, N_EVAL = 10, 18


def parse_vk(buf):
    logn, npub, hc = (int.from_bytes(buf[4 * i:4 * i + 4], "big") for i in range(3))
    o = 12
    k1, k2, omega = (int.from_bytes(buf[o + 32 * i:o + 32 * i + 32], "big") for i in range(3))
    o += 96
    com = [g1_from_raw(buf[o + 64 * i:o + 64 * i + 64]) for i in range(9)]
    o += 9 * 64
    tau2 = g2_from_raw(buf[o:o + 128])
    assert o + 128 == len(buf)
    digest = hashlib.sha256(b"".join(buf[12 + 96 + 64 * i:12 + 96 + 64 * i + 64] for i in range(9))
                            + (1 << logn).to_bytes(4, "little") + npub.to_bytes(4, "little")).digest()
    return {"logN": logn, "n_public_rows": npub, "has_commit": bool(hc), "k1": k1, "k2": k2, "omega": omega, "com": com,
            "tau2": tau2, "digest": digest, "raw_com": [buf[12 + 96 + 64 * i:12 + 96 + 64 * i + 64] for i in range(9)]}


def parse_proof(buf):
    assert len(buf) == N_COM * 64 + N_EVAL * 32
    raw = [buf[64 * i:64 * i + 64] for i in range(N_COM)]
    com = [g1_from_raw(r) for r in raw]
    ev = [int.from_bytes(buf[N_COM * 64 + 32 * i:N_COM * 64 + 32 * i + 32], "big") for i in range(N_EVAL)]
    return {"com": com, "raw_com": raw, "evals": ev}


def _fr(x):
    return (x % R).to_bytes(32, "big")


def verify(vk, proof, public_inputs):
    """public_inputs: the circuit's public inputs (without the leading ONE and without the commitment challenge).
    Returns (ok, reason)."""
    inv = lambda v: pow(v, R - 2, R)
    n = 1 << vk["logN"]
    com, ev = proof["com"], proof["evals"]
    for i, p in enumerate(com):
        if p is not None and not ec_on_curve(1, p):
            return False, "commitment %d not on the curve" % i
    if any(e >= R for e in ev):
        return False, "evaluation not reduced"
    if vk["omega"] != root_of_unity(vk["logN"]):
        return False, "vk.omega is not the domain generator"
    # public vector: ONE, the public inputs, then (as in the Groth16 path) the challenge derived from the P2 commitment
    xs = [1] + [v % R for v in public_inputs]
    if vk["has_commit"]:
        xs.append(hash_to_fr(proof["raw_com"][3]))
    if len(xs) != vk["n_public_rows"]:
        return False, "public input count %d != %d" % (len(xs), vk["n_public_rows"])
    # transcript
    beta = hash_to_fr(vk["digest"] + b"".join(_fr(x) for x in xs) + b"".join(proof["raw_com"][0:4]), b"gpw-plonk-beta")
    gamma = hash_to_fr(_fr(beta), b"gpw-plonk-gamma")
    alpha = hash_to_fr(_fr(gamma) + proof["raw_com"][4], b"gpw-plonk-alpha")
    zeta = hash_to_fr(_fr(alpha) + b"".join(proof["raw_com"][5:8]), b"gpw-plonk-zeta")
    nu = hash_to_fr(_fr(zeta) + b"".join(_fr(e) for e in ev), b"gpw-plonk-nu")
    u = hash_to_fr(_fr(nu) + proof["raw_com"][8] + proof["raw_com"][9], b"gpw-plonk-u")
    a, b, c, z, p2, ql, qr, qm, qo, qc, qcp, s1, s2, s3, t0, t1, t2, zw = ev
    w = vk["omega"]
    zn = pow(zeta, n, R)
    zh = (zn - 1) % R
    if zh == 0:
        return False, "zeta in the domain"
    # PI(zeta) = -sum x_i L_i(zeta),  L_i(zeta) = w^i (zeta^n - 1) / (n (zeta - w^i))
    pi = 0
    wi = 1
    for x in xs:
        pi = (pi - x * wi % R * zh % R * inv(n * (zeta - wi) % R)) % R
        wi = wi * w % R
    l0 = zh * inv(n * (zeta - 1) % R) % R
    gate = (ql * a + qr * b + qm * a % R * b + qo * c + qc + pi + qcp * p2) % R
    k1, k2 = vk["k1"], vk["k2"]
    perm = (z * (a + beta * zeta + gamma) % R * (b + beta * k1 % R * zeta + gamma) % R * (c + beta * k2 % R * zeta + gamma)
            - zw * (a + beta * s1 + gamma) % R * (b + beta * s2 + gamma) % R * (c + beta * s3 + gamma)) % R
    lhs = (gate + alpha * perm + alpha * alpha % R * (z - 1) % R * l0) % R
    rhs = zh * (t0 + zn * t1 + zn * zn % R * t2) % R
    if lhs != rhs:
        return False, "quotient identity fails at zeta"
    # batched KZG openings: e(F - [y] + zeta Wz + u (Z - [zw] + zeta w Wzw), G2) = e(Wz + u Wzw, [tau]2)
    polys = [com[0], com[1], com[2], com[4], com[3]] + vk["com"] + [com[5], com[6], com[7]]   # order of the 17 evaluations
    F, y, nk = None, 0, 1
    for pcom, e in zip(polys, ev[:17]):
        F = ec_add(1, F, ec_mul(1, pcom, nk))
        y = (y + nk * e) % R
        nk = nk * nu % R
    wz, wzw = com[8], com[9]
    left = ec_add(1, F, ec_neg(1, ec_mul(1, G1_GEN, y)))
    left = ec_add(1, left, ec_mul(1, wz, zeta))
    zpart = ec_add(1, com[4], ec_neg(1, ec_mul(1, G1_GEN, zw)))
    zpart = ec_add(1, zpart, ec_mul(1, wzw, zeta * w % R))
    left = ec_add(1, left, ec_mul(1, zpart, u))
    right = ec_add(1, wz, ec_mul(1, wzw, u))
    ok = pairing_product_is_one([(left, G2_GEN), (ec_neg(1, right), vk["tau2"])])
    return ok, "" if ok else "KZG opening check fails"
