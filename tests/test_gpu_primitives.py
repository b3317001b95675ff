"""Parity tests proper: every CUDA kernel, called through the C ABI, against the oracle. Needs a B200."""
import os
import random

import numpy as np
import pytest

import gpw
from oracle import bn254 as ob
from oracle import goldilocks as ogl
from oracle.engine import Api
from oracle.poseidon import GoldilocksChip, BN254Chip
from oracle.verifier import verify_testdata

pytestmark = pytest.mark.gpu

GLP = ogl.P


@pytest.fixture(scope="module")
def ctx():
    c = gpw.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="module")
def step_trace(testdata_dir):
    api, chip = verify_testdata(os.path.join(testdata_dir, "step"))
    return api, chip


def test_field_selftest(ctx):
    # PTX even/odd multiply == portable CIOS on device == host, Fr and Fp, incl. 0, 1, (p-1)^2
    ctx.selftest_ff(1 << 16, seed=42)


# ---- K1: hints ------------------------------------------------------------------------------------
def test_gl_hints_random_and_edges(ctx):
    rng = random.Random(1)
    edge = [0, 1, 2, GLP - 1, GLP - 2, 1 << 32, (1 << 32) - 1, 1 << 63, 0xffffffff00000000]
    a = edge + [rng.randrange(GLP) for _ in range(5000)]
    b = list(reversed(edge)) + [rng.randrange(GLP) for _ in range(5000)]
    c = [GLP - 1] * len(edge) + [rng.randrange(GLP) for _ in range(5000)]
    q, r = ctx.gl_mul_add_hint(a, b, c)
    exp = [ogl.mul_add_hint(x, y, z) for x, y, z in zip(a, b, c)]
    assert list(map(int, q)) == [e[0] for e in exp] and list(map(int, r)) == [e[1] for e in exp]
    # goldilocks/base_test.go:108-114
    q, r = ctx.gl_mul_add_hint([1 << 63], [1 << 63], [3])
    assert int(r[0]) == 18446744068340842500
    xs = [0, GLP - 1, GLP, GLP + 1, (1 << 64) - 1, 1 << 64, (1 << 144) - 1, (1 << 192) + 12345, (1 << 208) - 1,
          ob.R - 1] + [rng.randrange(1 << rng.choice([64, 100, 128, 144, 192, 208, 253])) for _ in range(5000)]
    q4, r = ctx.gl_reduce_hint(gpw.ints_to_limbs(xs))
    assert gpw.limbs_to_ints(q4) == [x // GLP for x in xs]
    assert list(map(int, r)) == [x % GLP for x in xs]
    inv = ctx.gl_inverse_hint(a)
    assert list(map(int, inv)) == [ogl.inverse_hint(x) for x in a]
    hi, lo = ctx.gl_split_limbs_hint(a)
    assert [(int(h), int(l)) for h, l in zip(hi, lo)] == [ogl.split_limbs_hint(x) for x in a]


def test_gl_hints_reject_noncanonical(ctx):
    # SURVEY 8a trap 10: MulAddHint panics, SplitLimbsHint errors, InverseHint panics for x >= p
    with pytest.raises(gpw.GpwError) as e:
        ctx.gl_mul_add_hint([1, GLP], [1, 1], [0, 0])
    assert e.value.code == -5 and "input 1" in str(e.value)
    with pytest.raises(gpw.GpwError):
        ctx.gl_split_limbs_hint([GLP])
    with pytest.raises(gpw.GpwError):
        ctx.gl_inverse_hint([(1 << 64) - 1])
    q, r = ctx.gl_mul_add_hint([], [], [])
    assert q.size == 0


def test_gl_hints_on_real_step_trace(ctx, step_trace):
    # all 449k hint calls of testdata/step, in the reference's order, bit-exact
    api, _ = step_trace
    ma = [(i, o) for k, i, o in api.hints if k == "muladd"]
    q, r = ctx.gl_mul_add_hint([i[0] for i, _ in ma], [i[1] for i, _ in ma], [i[2] for i, _ in ma])
    assert len(ma) == 44186
    assert list(map(int, q)) == [o[0] for _, o in ma] and list(map(int, r)) == [o[1] for _, o in ma]
    rd = [(i, o) for k, i, o in api.hints if k == "reduce"]
    assert len(rd) == 151410
    q4, r = ctx.gl_reduce_hint(gpw.ints_to_limbs([i[0] for i, _ in rd]))
    assert gpw.limbs_to_ints(q4) == [o[0] for _, o in rd] and list(map(int, r)) == [o[1] for _, o in rd]
    sp = [(i, o) for k, i, o in api.hints if k == "split"]
    assert len(sp) == 251232
    hi, lo = ctx.gl_split_limbs_hint([i[0] for i, _ in sp])
    assert list(map(int, hi)) == [o[0] for _, o in sp] and list(map(int, lo)) == [o[1] for _, o in sp]
    iv = [(i, o) for k, i, o in api.hints if k == "inverse"]
    assert len(iv) == 1849
    assert list(map(int, ctx.gl_inverse_hint([i[0] for i, _ in iv]))) == [o[0] for _, o in iv]


# ---- K2: Poseidon-Goldilocks --------------------------------------------------------------------------
def test_poseidon_gl(ctx, kats):
    rng = random.Random(2)
    states = [[0] * 12] + [[rng.randrange(GLP) for _ in range(12)] for _ in range(200)] + [[GLP - 1] * 12]
    out = ctx.poseidon_gl(np.array(states, dtype=np.uint64))
    assert list(map(int, out[0])) == [int(x) for x in kats["poseidon_gl_perm_zero"]]
    chip = GoldilocksChip(Api(trace=False))
    for s, o in zip(states, out):
        assert list(map(int, o)) == chip.Poseidon(s)


# ---- K3: Poseidon-BN254, leaf hashing, Merkle paths ---------------------------------------------------------
def test_poseidon_bn254_kats_and_random(ctx, kats):
    rng = random.Random(3)
    ins = [[int(x) for x in c["in"]] for c in kats["poseidon_bn254"]]
    ins += [[rng.randrange(ob.R) for _ in range(4)] for _ in range(300)]
    flat = gpw.ints_to_limbs([v for s in ins for v in s]).reshape(-1, 16)
    out = ctx.poseidon_bn254(flat, mont=False)
    got = gpw.limbs_to_ints(out.reshape(-1, 4))
    chip = BN254Chip(Api(trace=False))
    for i, s in enumerate(ins):
        exp = chip.Poseidon(s)
        assert got[4 * i:4 * i + 4] == exp
    for i, c in enumerate(kats["poseidon_bn254"]):
        assert got[4 * i:4 * i + 4] == [int(x) for x in c["out"]]
    # Montgomery in/out gives the same permutation
    m_in = gpw.host_ff_to_mont(0, flat.reshape(-1, 4)).reshape(-1, 16)
    m_out = ctx.poseidon_bn254(m_in, mont=True)
    assert gpw.limbs_to_ints(gpw.host_ff_from_mont(0, m_out.reshape(-1, 4))) == got


def test_hash_or_noop(ctx):
    rng = random.Random(4)
    chip = BN254Chip(Api(trace=False))
    for leaf_len in (0, 1, 2, 3, 4, 9, 10, 16, 20, 32, 86, 136):
        leaves = [[rng.randrange(GLP) for _ in range(leaf_len)] for _ in range(5)]
        arr = np.array(leaves, dtype=np.uint64).reshape(5, leaf_len)
        got = gpw.limbs_to_ints(ctx.hash_or_noop_bn254(arr))
        assert got == [chip.HashOrNoop(l) for l in leaves], leaf_len


def test_all_merkle_paths_of_step_proof(ctx, testdata_dir):
    # every one of the 28 x 6 Merkle paths of testdata/step must land on its cap entry
    # (fri/fri.go:97-157; verifier/verifier_test.go end-to-end assertion)
    from oracle import types as ot
    d = os.path.join(testdata_dir, "step")
    common = ot.read_common_circuit_data(os.path.join(d, "common_circuit_data.json"))
    proof, pis = ot.read_proof_with_public_inputs(os.path.join(d, "proof_with_public_inputs.json"))
    vd = ot.read_verifier_only_circuit_data(os.path.join(d, "verifier_only_circuit_data.json"))
    api, chip = verify_testdata(d, trace=False)
    idxs = chip.challenges.FriChallenges.FriQueryIndices
    caps = [vd.ConstantSigmasCap, proof.WiresCap, proof.PlonkZsPartialProductsCap, proof.QuotientPolysCap]
    lde_bits = common.FriParams.LdeBits()
    by_depth = {}
    for q, qr in enumerate(proof.OpeningProof.QueryRoundProofs):
        x = idxs[q] % (1 << lde_bits)
        cap_idx = x >> (lde_bits - 4)
        for t, (leaf, sib) in enumerate(qr.EvalsProofs):
            by_depth.setdefault((len(sib), len(leaf)), []).append((leaf, sib, x, caps[t][cap_idx]))
        bits = x
        for s, (evals, sib) in enumerate(qr.Steps):
            bits >>= common.FriParams.ReductionArityBits[s]
            leaf = [v for e in evals for v in e]
            by_depth.setdefault((len(sib), len(leaf)), []).append(
                (leaf, sib, bits, proof.OpeningProof.CommitPhaseMerkleCaps[s][cap_idx]))
    total = 0
    for (depth, leaf_len), items in by_depth.items():
        leaves = np.array([it[0] for it in items], dtype=np.uint64)
        digests = ctx.hash_or_noop_bn254(leaves)
        sibs = gpw.ints_to_limbs([s for it in items for s in it[1]]).reshape(len(items), depth, 4)
        roots = ctx.merkle_paths_bn254(digests, sibs, [it[2] for it in items], depth)
        assert gpw.limbs_to_ints(roots) == [it[3] for it in items]
        total += len(items)
    assert total == 28 * 6


# ---- K9: MSM ------------------------------------------------------------------------------------------
def _msm_case(ctx, group, scalars, ks, window_bits=0, mont=False):
    """points = [k_i]G; MSM must equal [sum s_i k_i]G (exact group-element equality)."""
    gen = ob.G1_GEN if group == 1 else ob.G2_GEN
    # build the distinct multiples on the host: consecutive run then pick
    kmin, kmax = min(ks), max(ks)
    table = gpw.host_ec_generator_multiples(group, kmin, kmax - kmin + 1)
    pts = table[[k - kmin for k in ks]]
    sl = gpw.ints_to_limbs(scalars)
    if mont:
        sl = gpw.host_ff_to_mont(0, sl)
    out = ctx.msm(group, sl, pts, scalars_mont=mont, window_bits=window_bits)
    exp = ob.ec_mul(group, gen, sum(s * k for s, k in zip(scalars, ks)) % ob.R)
    assert gpw.points_to_ints(group, out)[0] == ob.point_key(group, exp)


@pytest.mark.parametrize("group", [1, 2])
def test_msm_small_vs_naive(ctx, group):
    rng = random.Random(10 + group)
    gen = ob.G1_GEN if group == 1 else ob.G2_GEN
    n = 37
    ks = [rng.randrange(1, 1000) for _ in range(n)]
    scalars = [rng.randrange(ob.R) for _ in range(n)]
    pts_py = [ob.ec_mul(group, gen, k) for k in ks]
    exp = ob.msm_naive(group, scalars, pts_py)
    pts = gpw.ints_to_points(group, [ob.point_key(group, p) for p in pts_py])
    for c in (0, 4, 7, 16):
        out = ctx.msm(group, gpw.ints_to_limbs(scalars), pts, window_bits=c)
        assert gpw.points_to_ints(group, out)[0] == ob.point_key(group, exp), c


@pytest.mark.parametrize("group", [1, 2])
def test_msm_edge_cases(ctx, group):
    R = ob.R
    words = 8 if group == 1 else 16
    # empty -> infinity
    out = ctx.msm(group, np.zeros((0, 4), np.uint64), np.zeros((0, words), np.uint64))
    assert not out.any()
    # all-zero scalars -> infinity; single point; scalar r-1 (= -P); max digits
    _msm_case(ctx, group, [0, 0, 0], [1, 2, 3])
    _msm_case(ctx, group, [1], [5])
    _msm_case(ctx, group, [R - 1], [5])
    _msm_case(ctx, group, [R - 1, 1], [5, 5])          # P + (-P) = infinity via the bucket path
    _msm_case(ctx, group, [1, 1, 1, 1], [3, 3, 3, 3])  # equal points in one bucket: doubling branch
    _msm_case(ctx, group, [(1 << 254) - 1 - (1 << 200), 0x8000, 0x7fff, 0x8001, 0xffff, 0x10000], [2, 3, 4, 5, 6, 7])
    # a point at infinity among the bases
    pts = gpw.host_ec_generator_multiples(group, 1, 3)
    pts[1] = 0
    out = ctx.msm(group, gpw.ints_to_limbs([5, 7, 11]), pts)
    gen = ob.G1_GEN if group == 1 else ob.G2_GEN
    assert gpw.points_to_ints(group, out)[0] == ob.point_key(group, ob.ec_mul(group, gen, 5 * 1 + 11 * 3))


@pytest.mark.parametrize("group,n", [(1, 1 << 14), (2, 1 << 12)])
def test_msm_medium_linearity(ctx, group, n):
    rng = random.Random(20 + group)
    ks = list(range(1, n + 1))
    # witness-shaped scalar mix: 35% < 2^16 (many 0/1), 45% < 2^64, 20% full width
    scalars = []
    for _ in range(n):
        u = rng.random()
        if u < 0.15:
            scalars.append(rng.randrange(2))
        elif u < 0.35:
            scalars.append(rng.randrange(1 << 16))
        elif u < 0.80:
            scalars.append(rng.randrange(1 << 64))
        else:
            scalars.append(rng.randrange(ob.R))
    _msm_case(ctx, group, scalars, ks)
    _msm_case(ctx, group, scalars, ks, mont=True)
    _msm_case(ctx, group, [rng.randrange(ob.R) for _ in range(n)], ks, window_bits=13)


def test_msm_hot_bucket(ctx):
    # one bucket holding 60k entries (all scalars equal 1): exercises the multi-task fix-up path
    n = 60000
    _msm_case(ctx, 1, [1] * n, list(range(1, n + 1)))
    _msm_case(ctx, 1, [3] * (n // 2) + [ob.R - 3] * (n // 2), list(range(1, n + 1)))


@pytest.mark.parametrize("group,n", [(1, 8192), (1, 16384), (1, 8192 * 3 + 77), (2, 8192 * 2 + 5)])
def test_msm_hot_bucket_whole_ctas(ctx, group, n):
    # a bucket that fills whole accumulate CTAs (128 threads x 64 entries): the in-CTA merge of the 128 partials,
    # for a bucket that is exactly one CTA, exactly two, and one that starts and ends inside neighbouring CTAs;
    # a few other scalars before and after it so the hot bucket is neither the first nor the last one
    ks = list(range(1, n + 1))
    _msm_case(ctx, group, [7] * n, ks, window_bits=16)
    lead, trail = [1, 2, 3, 5, 6], [9, 11, 0xffff, 1 << 40]
    _msm_case(ctx, group, lead + [7] * n + trail, list(range(1, n + len(lead) + len(trail) + 1)), window_bits=16)
    # the same base repeated inside the hot bucket: doubling / cancellation leave the straight-line path
    _msm_case(ctx, group, [7] * n, [1 + (i % 5) for i in range(n)], window_bits=16)


def test_msm_window_split_matches_full(ctx):
    # multi-GPU window split: partial results over disjoint window ranges must add up to the full MSM
    import torch
    rng = random.Random(31)
    n = 5000
    pts = gpw.host_ec_generator_multiples(1, 1, n)
    scalars = [rng.randrange(ob.R) for _ in range(n)]
    sl = gpw.ints_to_limbs(scalars)
    full = ctx.msm(1, sl, pts, window_bits=16)
    ds = torch.from_numpy(sl.view(np.int64)).cuda()
    dp = torch.from_numpy(pts.view(np.int64)).cuda()
    torch.cuda.synchronize()
    parts = [ctx.msm_dev(1, ds.data_ptr(), dp.data_ptr(), n, window_bits=16, win_lo=lo, win_hi=hi)
             for lo, hi in ((0, 5), (5, 11), (11, 16))]
    acc = parts[0]
    for p in parts[1:]:
        acc = gpw.host_ec_add(1, acc, p)
    assert (acc == full).all()


@pytest.mark.parametrize("window_bits", [7, 16, 22])
def test_msm_fixed_base_matches_windowed(ctx, window_bits):
    # fixed-base mode (precomputed 2^(c w) P_i, one bucket set) must give the same group element as the windowed MSM
    import torch
    rng = random.Random(40 + window_bits)
    n = 3000
    pts = gpw.host_ec_generator_multiples(1, 5, n)
    pts[17] = 0                                              # a base at infinity
    scalars = [rng.randrange(ob.R) if i % 4 else rng.randrange(1 << 20) for i in range(n)]
    scalars[0], scalars[1], scalars[2] = 0, 1, ob.R - 1
    sl = gpw.ints_to_limbs(scalars)
    ref = ctx.msm(1, sl, pts)
    ds = torch.from_numpy(sl.view(np.int64)).cuda()
    dp = torch.from_numpy(pts.view(np.int64)).cuda()
    W = ctx.msm_fixed_windows(window_bits)
    table = torch.zeros((W * n, 8), dtype=torch.int64, device="cuda")
    torch.cuda.synchronize()
    ctx.msm_g1_fixed_table(dp.data_ptr(), n, window_bits, table.data_ptr())
    out = ctx.msm_g1_fixed_dev(ds.data_ptr(), table.data_ptr(), n, window_bits)
    assert (out == ref).all()
    # table rows really are 2^(c w) P_i
    row = table[2 * n + 3].cpu().numpy().view(np.uint64)
    exp = ob.ec_mul(1, ob.G1_GEN, (8 << (2 * window_bits)) % ob.R)
    assert gpw.points_to_ints(1, row)[0] == ob.point_key(1, exp)


@pytest.mark.parametrize("rounds", [1, 3, 5])
def test_msm_batch_affine_rounds_bit_identical(ctx, rounds):
    # csrc/msm_affine.cuh (gnark-crypto's batch-affine bucket accumulation as a pairwise tree): the opt-in rounds must give
    # the same group element as the XYZZ-only path on the witness mix, on hot buckets that span many CTAs, on buckets holding
    # the same base several times (doubling), P + (-P) (cancellation), bases at infinity, and for G2
    rng = random.Random(50 + rounds)
    try:
        ctx.set_option("msm_affine_rounds", rounds)
        n = 40000
        ks = list(range(1, n + 1))
        mix = [rng.choice((rng.randrange(2), rng.randrange(1 << 16), rng.randrange(1 << 64), rng.randrange(ob.R))) for _ in range(n)]
        _msm_case(ctx, 1, mix, ks)
        _msm_case(ctx, 1, mix, ks, mont=True, window_bits=9)
        _msm_case(ctx, 1, [7] * n, ks, window_bits=16)                                   # one hot bucket
        _msm_case(ctx, 1, [7] * n, [1 + (i % 5) for i in range(n)], window_bits=16)      # equal points: doubling
        _msm_case(ctx, 1, [3] * (n // 2) + [ob.R - 3] * (n // 2), [1 + (i % 7) for i in range(n)])  # cancellations
        _msm_case(ctx, 2, mix[:6000], ks[:6000])
        _msm_case(ctx, 2, [5] * 3000, [1 + (i % 3) for i in range(3000)], window_bits=12)
        pts = gpw.host_ec_generator_multiples(1, 1, 5000)
        pts[::7] = 0                                                                      # bases at infinity
        sc = [rng.randrange(1 << 20) for _ in range(5000)]
        out = ctx.msm(1, gpw.ints_to_limbs(sc), pts, window_bits=8)
        exp = ob.ec_mul(1, ob.G1_GEN, sum(s * (i + 1) for i, s in enumerate(sc) if i % 7) % ob.R)
        assert gpw.points_to_ints(1, out)[0] == ob.point_key(1, exp)
    finally:
        ctx.set_option("msm_affine_rounds", 0)


def _dot_mod_r(scalars_u64x4, first_k):
    """sum_i scalars[i] * (first_k + i) mod r for (n, 4) little-endian u64 limbs, exactly, with numpy: 16-bit pieces keep
    every partial sum below 2^64 (piece < 2^16, multiplier < 2^24, n <= 2^23)."""
    n = scalars_u64x4.shape[0]
    ks = np.arange(first_k, first_k + n, dtype=np.uint64)
    assert first_k + n < (1 << 24) and n <= (1 << 23)
    total = 0
    for limb in range(4):
        col = scalars_u64x4[:, limb]
        for piece in range(4):
            part = (col >> np.uint64(16 * piece)) & np.uint64(0xffff)
            total += int((part * ks).sum(dtype=np.uint64)) << (64 * limb + 16 * piece)
    return total % ob.R


def test_msm_full_size_windowed_and_fixed_base(ctx):
    # BASELINE-scale MSM (n = 2^23 - 1 G1 points, the size of the Z MSM of testdata/step): exact check through the known
    # discrete logs - points [k]G, k = 1..n, so the result must be [sum s_k k]G - for the windowed MSM on a witness-shaped
    # scalar mix and for the fixed-base MSM (22-bit windows, one bucket set) on full-width scalars
    import torch
    n = (1 << 23) - 1
    pts = torch.empty((n, 8), dtype=torch.int64, device="cuda")
    ctx.generator_multiples_dev(1, 1, n, pts.data_ptr())
    rng = np.random.default_rng(11)
    s = rng.integers(0, 1 << 62, size=(n, 4), dtype=np.uint64)
    s[:, 3] &= np.uint64((1 << 59) - 1)                    # < 2^251 < r
    full = s.copy()
    u = rng.random(n)
    s[u < 0.80, 1:] = 0                                     # 80 % at most 64 bits ...
    s[u < 0.35, 0] &= np.uint64(0xffff)                     # ... 35 % 16-bit limbs ...
    s[u < 0.15, 0] &= np.uint64(1)                          # ... 15 % bits (the hot bucket)
    for scalars, fixed in ((s, False), (full, True)):
        ds = torch.from_numpy(scalars.view(np.int64)).cuda()
        torch.cuda.synchronize()
        if fixed:
            W = ctx.msm_fixed_windows(22)
            table = torch.empty((W * n, 8), dtype=torch.int64, device="cuda")
            ctx.msm_g1_fixed_table(pts.data_ptr(), n, 22, table.data_ptr())
            out = ctx.msm_g1_fixed_dev(ds.data_ptr(), table.data_ptr(), n, 22)
            del table
        else:
            out = ctx.msm_dev(1, ds.data_ptr(), pts.data_ptr(), n)
        exp = ob.ec_mul(1, ob.G1_GEN, _dot_mod_r(scalars, 1))
        assert gpw.points_to_ints(1, out)[0] == ob.point_key(1, exp), "fixed-base" if fixed else "windowed"


# ---- K8: NTT ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("logn", [0, 1, 2, 3, 5, 8, 9, 10])
def test_ntt_small_vs_definition(ctx, logn):
    rng = random.Random(40 + logn)
    n = 1 << logn
    a = [rng.randrange(ob.R) for _ in range(n)]
    am = gpw.host_ff_to_mont(0, gpw.ints_to_limbs(a))
    ref = ob.dft_naive if logn <= 8 else ob.ntt_fast
    for inverse in (False, True):
        for coset in (False, True):
            exp = ref(a, inverse=inverse, coset=coset)
            got = gpw.limbs_to_ints(gpw.host_ff_from_mont(0, ctx.ntt(am, inverse=inverse, coset=coset)))
            assert got == exp, (logn, inverse, coset)
            # DIF leaves bit-reversed output; DIT consumes bit-reversed input
            got_br = gpw.limbs_to_ints(gpw.host_ff_from_mont(0, ctx.ntt(am, inverse=inverse, coset=coset, out_bitrev=True)))
            assert got_br == ob.bit_reverse_list(exp)
            a_br = gpw.host_ff_to_mont(0, gpw.ints_to_limbs(ob.bit_reverse_list(a)))
            got_dit = gpw.limbs_to_ints(gpw.host_ff_from_mont(0, ctx.ntt(a_br, inverse=inverse, coset=coset, in_bitrev=True)))
            assert got_dit == exp


@pytest.mark.parametrize("logn", [14, 17])
def test_ntt_medium_vs_fast_oracle(ctx, logn):
    rng = random.Random(50 + logn)
    n = 1 << logn
    a = [rng.randrange(ob.R) for _ in range(n)]
    am = gpw.host_ff_to_mont(0, gpw.ints_to_limbs(a))
    got = gpw.limbs_to_ints(gpw.host_ff_from_mont(0, ctx.ntt(am, coset=True)))
    assert got == ob.ntt_fast(a, coset=True)


def test_ntt_large_roundtrip_and_linearity(ctx):
    # size-independent properties at a BASELINE-scale size (2^22 here; 2^23 in bench): iNTT(NTT(a)) = a on
    # the coset, and NTT(delta_1) = (w^k)
    logn = 22
    n = 1 << logn
    rng = np.random.default_rng(5)
    a = rng.integers(0, 1 << 62, size=(n, 4), dtype=np.uint64)
    a[:, 3] &= np.uint64((1 << 60) - 1)   # < 2^252 < r: valid Montgomery residues
    f = ctx.ntt(a, coset=True, out_bitrev=True)
    b = ctx.ntt(f, inverse=True, coset=True, in_bitrev=True)
    assert (a == b).all()
    d = np.zeros((n, 4), dtype=np.uint64)
    d[1] = gpw.host_ff_to_mont(0, gpw.ints_to_limbs([1]))[0]
    spec = ctx.ntt(d)
    w = ob.root_of_unity(logn)
    ks = [0, 1, 2, 12345, n // 2, n - 1]
    got = gpw.limbs_to_ints(gpw.host_ff_from_mont(0, spec[ks]))
    assert got == [pow(w, k, ob.R) for k in ks]


def test_sharded_msm_shares_add_up_and_single_rank_communicator(ctx):
    # csrc/comm.cu on ONE GPU: (1) for both splits the shares of N = 2, 3, 8 and 20 virtual ranks (20 > the 16 windows:
    # empty shares) add up to the whole MSM, G1 and G2; (2) a real NCCL communicator of one rank runs the collective path
    # end to end (ncclCommInitRank, all-gather on the context's stream, device-side sum) and returns the same point
    import torch
    rng = random.Random(77)
    for group, n in ((1, 6000), (2, 1500)):
        pts = gpw.host_ec_generator_multiples(group, 1, n)
        sc = [rng.choice((rng.randrange(ob.R), rng.randrange(1 << 64), rng.randrange(2))) for _ in range(n)]
        sl = gpw.ints_to_limbs(sc)
        full = ctx.msm(group, sl, pts, window_bits=16)
        ds = torch.from_numpy(sl.view(np.int64)).cuda()
        dp = torch.from_numpy(pts.view(np.int64)).cuda()
        torch.cuda.synchronize()
        for split in (1, 2):
            for N in (2, 3, 8, 20):
                acc = np.zeros(8 if group == 1 else 16, dtype=np.uint64)
                for r in range(N):
                    acc = gpw.host_ec_add(group, acc, ctx.msm_sharded_partial(group, ds.data_ptr(), dp.data_ptr(), n, split, r, N,
                                                                              window_bits=16))
                assert (acc == full).all(), (group, split, N)
    c2 = gpw.Context(0)
    try:
        with pytest.raises(gpw.GpwError) as e:
            c2.msm_sharded(1, ds.data_ptr(), dp.data_ptr(), 10)
        assert e.value.code == -7                       # GPW_ENCCL: no communicator yet
        c2.comm_init(1, 0, gpw.comm_unique_id())
        info = c2.comm_info()
        assert info["ranks"] == 1 and info["rank"] == 0 and info["nccl_version"] > 20000
        pts = gpw.host_ec_generator_multiples(1, 1, 6000)
        sl = gpw.ints_to_limbs([rng.randrange(ob.R) for _ in range(6000)])
        ds = torch.from_numpy(sl.view(np.int64)).cuda()
        dp = torch.from_numpy(pts.view(np.int64)).cuda()
        torch.cuda.synchronize()
        full = c2.msm(1, sl, pts, window_bits=16)
        for split in (0, 1, 2):
            assert (c2.msm_sharded(1, ds.data_ptr(), dp.data_ptr(), 6000, window_bits=16, split=split) == full).all()
        c2.comm_destroy()
        assert c2.comm_info()["ranks"] == 0
    finally:
        c2.close()


def test_sharded_msm_nccl_all_gpus():
    # the in-library sharded MSM over every GPU of the box (NCCL over NVLink); the single-GPU test above covers the logic
    import subprocess, sys, torch
    ngpu = torch.cuda.device_count()
    if ngpu < 2:
        pytest.skip("needs >= 2 GPUs (the one-rank communicator path is covered above)")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(ngpu),
                          "--master-addr", "127.0.0.1", "--master-port", "29517", os.path.join(root, "tools", "sharded_msm_check.py"), "18"],
                         capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.count("bit-identical") == 4
