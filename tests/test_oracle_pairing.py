"""The pairing oracle (oracle/pairing.py) is pinned by mathematics: bilinear, non-degenerate, order r; RFC 9380 vector for
expand_message_xmd; and a complete toy Groth16 (setup, prove, verify) over a 3-constraint R1CS done on Python integers."""
import random

from oracle import bn254 as ob
from oracle import pairing as op

R = ob.R


def test_pairing_is_bilinear_nondegenerate_of_order_r():
    e = op.pairing(ob.G1_GEN, ob.G2_GEN)
    assert not op.f12_eq(e, op.f12_one())
    assert op.f12_eq(op.f12_pow(e, R), op.f12_one())
    rng = random.Random(5)
    a, b = rng.randrange(R), rng.randrange(R)
    pa, qb = ob.ec_mul(1, ob.G1_GEN, a), ob.ec_mul(2, ob.G2_GEN, b)
    assert op.f12_eq(op.pairing(pa, qb), op.f12_pow(e, a * b % R))
    # additivity in each argument, and the product form with one final exponentiation
    assert op.f12_eq(op.pairing(ob.ec_add(1, pa, ob.G1_GEN), ob.G2_GEN), op.f12_pow(e, (a + 1) % R))
    assert op.pairing_product_is_one([(pa, ob.G2_GEN), (ob.ec_neg(1, ob.G1_GEN), ob.ec_mul(2, ob.G2_GEN, a))])
    assert not op.pairing_product_is_one([(pa, ob.G2_GEN), (ob.G1_GEN, ob.ec_mul(2, ob.G2_GEN, a))])
    assert op.f12_eq(op.pairing(None, ob.G2_GEN), op.f12_one())


def test_expand_message_xmd_rfc9380_vector():
    # RFC 9380 appendix K.1 (SHA-256), DST = "QUUX-V01-CS02-with-expander-SHA256-128", msg = "", len_in_bytes = 0x20
    dst = b"QUUX-V01-CS02-with-expander-SHA256-128"
    assert op.expand_message_xmd(b"", dst, 32).hex() == "68a985b87eb6b46952128911f2a4412bbc302a9d759667f87f7a21d803f07235"
    assert op.expand_message_xmd(b"abc", dst, 32).hex() == "d8ccab23b5985ccea865c6c97b6e5b8350e794e603b4b97902f53a8a0d605615"


def test_raw_point_codecs_round_trip():
    p = ob.ec_mul(1, ob.G1_GEN, 77)
    q = ob.ec_mul(2, ob.G2_GEN, 99)
    assert op.g1_from_raw(op.g1_raw(p)) == p and op.g1_from_raw(op.g1_raw(None)) is None
    q2 = op.g2_from_raw(op.g2_raw(q))
    assert ob.point_key(2, q2) == ob.point_key(2, q) and op.g2_from_raw(op.g2_raw(None)) is None
    # A1 before A0 (gnark-crypto E2 marshalling)
    assert op.g2_raw(q)[:32] == q[0].b.to_bytes(32, "big")


def _toy_groth16(rng, witness_ok=True):
    """x^3 + x + 5 = out as an R1CS over wires (1, out, x, v1 = x x, v2 = v1 x); out public. No commitment."""
    rows = [({2: 1}, {2: 1}, {3: 1}), ({3: 1}, {2: 1}, {4: 1}), ({4: 1, 2: 1, 0: 5}, {0: 1}, {1: 1})]
    m, n_pub, N = 5, 1, 4
    w = ob.root_of_unity(2)
    tau, alpha, beta, gamma, delta = (rng.randrange(1, R) for _ in range(5))
    inv = lambda v: pow(v, R - 2, R)

    def lagrange(j, x):
        num = (pow(x, N, R) - 1) * pow(w, j, R) % R
        return num * inv(N * (x - pow(w, j, R)) % R) % R

    A, B, C = [0] * m, [0] * m, [0] * m
    for j, (l, r_, o) in enumerate(rows):
        lj = lagrange(j, tau)
        for vec, le in ((A, l), (B, r_), (C, o)):
            for i, c in le.items():
                vec[i] = (vec[i] + c * lj) % R
    g1 = lambda s: ob.ec_mul(1, ob.G1_GEN, s)
    g2 = lambda s: ob.ec_mul(2, ob.G2_GEN, s)
    kk = [(beta * A[i] + alpha * B[i] + C[i]) % R for i in range(m)]
    vk = {"alpha1": g1(alpha), "beta2": g2(beta), "gamma2": g2(gamma), "delta2": g2(delta),
          "K": [g1(kk[i] * inv(gamma) % R) for i in range(n_pub + 1)]}
    zt = (pow(tau, N, R) - 1) % R
    # witness
    x = 3
    wv = [1, (x**3 + x + 5) % R, x, x * x % R, x**3 % R]
    if not witness_ok:
        wv[3] = (wv[3] + 1) % R
    # h = (A.B - C)/Z_H: evaluate on a coset like the prover does (only exact when the witness satisfies the rows)
    av = [sum(c * wv[i] for i, c in l.items()) % R for l, _, _ in rows] + [0]
    bv = [sum(c * wv[i] for i, c in r_.items()) % R for _, r_, _ in rows] + [0]
    cv = [sum(c * wv[i] for i, c in o.items()) % R for _, _, o in rows] + [0]
    ac, bc, cc = (ob.ntt_fast(ob.ntt_fast(v, inverse=True), coset=True) for v in (av, bv, cv))
    zc = inv((pow(ob.FR_GENERATOR, N, R) - 1) % R)
    h = ob.ntt_fast([(a * b - c) * zc % R for a, b, c in zip(ac, bc, cc)], inverse=True, coset=True)
    r_, s_ = rng.randrange(R), rng.randrange(R)
    sa = (alpha + sum(wv[i] * A[i] for i in range(m)) + r_ * delta) % R
    sb = (beta + sum(wv[i] * B[i] for i in range(m)) + s_ * delta) % R
    sk = sum(wv[i] * kk[i] for i in range(n_pub + 1, m)) * inv(delta) % R
    sz = sum(h[j] * pow(tau, j, R) for j in range(N - 1)) * zt % R * inv(delta) % R
    skrs = (sk + sz + s_ * sa + r_ * sb - r_ * s_ * delta) % R
    proof = {"Ar": g1(sa), "Bs": g2(sb), "Krs": g1(skrs), "commitments": [], "pok": None}
    return vk, proof, wv[1:1 + n_pub]


def test_groth16_verify_on_a_toy_circuit_done_in_python():
    rng = random.Random(11)
    vk, proof, public = _toy_groth16(rng)
    ok, why = op.groth16_verify(vk, proof, public)
    assert ok, why
    assert not op.groth16_verify(vk, proof, [public[0] + 1])[0]
    vk2, proof2, public2 = _toy_groth16(rng, witness_ok=False)
    assert not op.groth16_verify(vk2, proof2, public2)[0]
    # raw proof bytes round trip through the gnark layout
    buf = op.g1_raw(proof["Ar"]) + op.g2_raw(proof["Bs"]) + op.g1_raw(proof["Krs"]) + (0).to_bytes(4, "big") + op.g1_raw(None)
    back = op.parse_proof_raw(buf)
    assert back["Ar"] == proof["Ar"] and ob.point_key(2, back["Bs"]) == ob.point_key(2, proof["Bs"]) and back["commitments"] == []
    assert op.groth16_verify(vk, back, public)[0]


def test_kzg_opening_check():
    rng = random.Random(3)
    tau = rng.randrange(1, R)
    poly = [rng.randrange(R) for _ in range(6)]
    ev = lambda x: sum(c * pow(x, i, R) for i, c in enumerate(poly)) % R
    z = rng.randrange(R)
    y = ev(z)
    # quotient (p(X) - y)/(X - z) by synthetic division
    q, acc = [0] * 5, 0
    for i in range(5, 0, -1):
        acc = (poly[i] + acc * z) % R
        q[i - 1] = acc
    com = ob.ec_mul(1, ob.G1_GEN, ev(tau))
    h = ob.ec_mul(1, ob.G1_GEN, sum(c * pow(tau, i, R) for i, c in enumerate(q)) % R)
    g2_tau = ob.ec_mul(2, ob.G2_GEN, tau)
    assert op.kzg_verify(com, z, y, h, g2_tau)
    assert not op.kzg_verify(com, z, (y + 1) % R, h, g2_tau)
