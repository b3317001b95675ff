"""oracle/plonk_verify.py pinned without a GPU: a complete toy PLONK prover on Python integers (4-row circuit x y = z, z + 5 =
out with `out` public; textbook polynomial arithmetic, commitments computed directly from a known tau) emits a proof in the
library's byte format; the verifier accepts it and rejects tampered statements, evaluations, commitments and opening proofs.
The GPU prover (csrc/plonk.cu) is checked by the same verifier in tests/test_gpu_plonk.py."""
import hashlib
import random

from oracle import bn254 as ob
from oracle import pairing as opair
from oracle import plonk_verify as pv

R = ob.R
inv = lambda v: pow(v, R - 2, R)


def padd(a, b):
    n = max(len(a), len(b))
    return [((a[i] if i < len(a) else 0) + (b[i] if i < len(b) else 0)) % R for i in range(n)]


def pscale(a, k):
    return [x * k % R for x in a]


def pmul(a, b):
    out = [0] * (len(a) + len(b) - 1)
    for i, x in enumerate(a):
        for j, y in enumerate(b):
            out[i + j] = (out[i + j] + x * y) % R
    return out


def peval(a, x):
    acc = 0
    for c in reversed(a):
        acc = (acc * x + c) % R
    return acc


def pdiv_linear(a, z):
    """(a(X) - a(z)) / (X - z) by synthetic division"""
    q, acc = [0] * (len(a) - 1), 0
    for i in range(len(a) - 1, 0, -1):
        acc = (a[i] + acc * z) % R
        q[i - 1] = acc
    return q


def pdiv_zh(a, n):
    """a(X) / (X^n - 1), exact"""
    a = a[:]
    q = [0] * (len(a) - n)
    for i in range(len(a) - 1, n - 1, -1):
        q[i - n] = a[i]
        a[i - n] = (a[i - n] + a[i]) % R
        a[i] = 0
    assert all(x == 0 for x in a), "not divisible by Z_H"
    return q


def toy_proof(x, y, tamper_witness=False):
    logn, n = 2, 4
    w = ob.root_of_unity(logn)
    k1, k2 = 5, 25
    tau = 0x1234567890abcdef1234567 % R
    interp = lambda ev: ob.ntt_fast(list(ev), inverse=True)
    z_ = x * y % R
    out = (z_ + 5) % R
    if tamper_witness:
        z_ = (z_ + 1) % R
    var = {"one": 1, "x": x, "y": y, "z": z_, "out": out}
    # rows: public ONE, public out, x y - z = 0, z + 5 - out = 0; unused slots hold variable "one"
    A, B, C = ["one", "out", "x", "z"], ["one", "one", "y", "one"], ["one", "one", "z", "out"]
    ql, qr, qm, qo, qc = [1, 1, 0, 1], [0, 0, 0, 0], [0, 0, 1, 0], [0, 0, R - 1, R - 1], [0, 0, 0, 5]
    xs = [1, out]
    slots = A + B + C
    sigma = list(range(12))
    for name in set(slots):
        idx = [i for i, s in enumerate(slots) if s == name]
        for i, s in enumerate(idx):
            sigma[s] = idx[(i + 1) % len(idx)]
    kk = [1, k1, k2]
    s_ev = [[kk[sigma[c * n + i] // n] * pow(w, sigma[c * n + i] % n, R) % R for i in range(n)] for c in range(3)]
    a_ev, b_ev, c_ev = ([var[v] for v in col] for col in (A, B, C))
    polys = {k: interp(v) for k, v in dict(a=a_ev, b=b_ev, c=c_ev, ql=ql, qr=qr, qm=qm, qo=qo, qc=qc, qcp=[0] * n, s1=s_ev[0], s2=s_ev[1],
                                           s3=s_ev[2], p2=[0] * n, pi=[(-xs[0]) % R, (-xs[1]) % R, 0, 0]).items()}
    com = lambda f: ob.ec_mul(1, ob.G1_GEN, peval(f, tau)) if peval(f, tau) else None
    raw = lambda f: opair.g1_raw(com(f))
    fr = lambda v: (v % R).to_bytes(32, "big")
    vk_coms = [raw(polys[k]) for k in ("ql", "qr", "qm", "qo", "qc", "qcp", "s1", "s2", "s3")]
    vk = ((logn).to_bytes(4, "big") + len(xs).to_bytes(4, "big") + (0).to_bytes(4, "big") + fr(k1) + fr(k2) + fr(w) + b"".join(vk_coms)
          + opair.g2_raw(ob.ec_mul(2, ob.G2_GEN, tau)))
    digest = hashlib.sha256(b"".join(vk_coms) + n.to_bytes(4, "little") + len(xs).to_bytes(4, "little")).digest()
    H = opair.hash_to_fr
    c_a, c_b, c_c, c_p2 = raw(polys["a"]), raw(polys["b"]), raw(polys["c"]), raw(polys["p2"])
    beta = H(digest + b"".join(fr(v) for v in xs) + c_a + c_b + c_c + c_p2, b"gpw-plonk-beta")
    gamma = H(fr(beta), b"gpw-plonk-gamma")
    z_ev, acc = [], 1
    for i in range(n):
        z_ev.append(acc)
        wi = pow(w, i, R)
        num = (a_ev[i] + beta * wi + gamma) * (b_ev[i] + beta * k1 * wi + gamma) % R * (c_ev[i] + beta * k2 * wi + gamma) % R
        den = (a_ev[i] + beta * s_ev[0][i] + gamma) * (b_ev[i] + beta * s_ev[1][i] + gamma) % R * (c_ev[i] + beta * s_ev[2][i] + gamma) % R
        acc = acc * num % R * inv(den) % R
    zp = interp(z_ev)
    c_z = raw(zp)
    alpha = H(fr(gamma) + c_z, b"gpw-plonk-alpha")
    P = polys
    X = [0, 1]
    gate = padd(padd(padd(pmul(P["ql"], P["a"]), pmul(P["qr"], P["b"])), padd(pmul(P["qm"], pmul(P["a"], P["b"])), pmul(P["qo"], P["c"]))),
                padd(padd(P["qc"], P["pi"]), pmul(P["qcp"], P["p2"])))
    lin = lambda f, k: padd(padd(f, pscale(X, beta * k % R)), [gamma])
    p1 = pmul(pmul(zp, lin(P["a"], 1)), pmul(lin(P["b"], k1), lin(P["c"], k2)))
    zw = [c * pow(w, i, R) % R for i, c in enumerate(zp)]           # Z(w X)
    sl = lambda f, s: padd(padd(f, pscale(s, beta)), [gamma])
    p2 = pmul(pmul(zw, sl(P["a"], P["s1"])), pmul(sl(P["b"], P["s2"]), sl(P["c"], P["s3"])))
    l0 = [inv(n)] * n
    bound = pmul(padd(zp, [R - 1]), l0)
    numer = padd(gate, padd(pscale(padd(p1, pscale(p2, R - 1)), alpha), pscale(bound, alpha * alpha % R)))
    t = pdiv_zh(numer, n)                                           # raises if the witness does not satisfy the system
    t = t + [0] * (3 * n - len(t))
    assert all(c == 0 for c in t[3 * n:])
    tparts = [t[0:n], t[n:2 * n], t[2 * n:3 * n]]
    c_t = [raw(p) for p in tparts]
    zeta = H(fr(alpha) + b"".join(c_t), b"gpw-plonk-zeta")
    order = [P["a"], P["b"], P["c"], zp, P["p2"], P["ql"], P["qr"], P["qm"], P["qo"], P["qc"], P["qcp"], P["s1"], P["s2"], P["s3"]] + tparts
    evals = [peval(f, zeta) for f in order] + [peval(zp, zeta * w % R)]
    nu = H(fr(zeta) + b"".join(fr(e) for e in evals), b"gpw-plonk-nu")
    F, nk = [0], 1
    for f in order:
        F = padd(F, pscale(f, nk))
        nk = nk * nu % R
    wz = pdiv_linear(F, zeta)
    wzw = pdiv_linear(zp, zeta * w % R)
    proof = c_a + c_b + c_c + c_p2 + c_z + b"".join(c_t) + raw(wz) + raw(wzw) + b"".join(fr(e) for e in evals)
    return vk, proof, [out]


def test_toy_plonk_proof_verifies_and_tampering_is_rejected():
    vk_b, proof_b, public = toy_proof(7, 9)
    vk = pv.parse_vk(vk_b)
    ok, why = pv.verify(vk, pv.parse_proof(proof_b), public)
    assert ok, why
    assert not pv.verify(vk, pv.parse_proof(proof_b), [public[0] + 1])[0]
    for off in (10 * 64 + 31, 10 * 64 + 32 * 3 + 31, 10 * 64 + 32 * 17 + 31):      # a(zeta), z(zeta), z(zeta w)
        bad = bytearray(proof_b)
        bad[off] ^= 1
        assert not pv.verify(vk, pv.parse_proof(bytes(bad)), public)[0]
    bad = bytearray(proof_b)
    bad[0:64] = proof_b[64:128]                                                   # [a] replaced by [b]
    assert not pv.verify(vk, pv.parse_proof(bytes(bad)), public)[0]
    bad = bytearray(proof_b)
    bad[64 * 8:64 * 9] = opair.g1_raw(ob.ec_add(1, opair.g1_from_raw(proof_b[64 * 8:64 * 9]), ob.G1_GEN))   # W_zeta + G
    ok, why = pv.verify(vk, pv.parse_proof(bytes(bad)), public)
    assert not ok and "KZG" in why
    # another witness for the same circuit verifies too; a witness that violates a gate has no quotient
    vk_b2, proof_b2, public2 = toy_proof(123456789, R - 2)
    assert vk_b2 == vk_b and pv.verify(vk, pv.parse_proof(proof_b2), public2)[0]
    try:
        toy_proof(7, 9, tamper_witness=True)
        raise AssertionError("a non-satisfying witness produced a quotient")
    except AssertionError as e:
        assert "Z_H" in str(e)
