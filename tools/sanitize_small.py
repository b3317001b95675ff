#!/usr/bin/env python3
"""Small workload for compute-sanitizer (memcheck / racecheck): NTTs of 2^14 (one TMA-staged pass + one register-staged pass per
transform), a windowed and a fixed-base MSM, and complete wrap proofs of two gadget circuits (spine with a Poseidon-Goldilocks
macro chunk, wide levels, R1CS evaluation, computeH, deferred MSMs).
usage: compute-sanitizer --tool memcheck|racecheck python tools/sanitize_small.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "gnark-plonky2-verifier_b200"))
import gpw  # noqa: E402

ctx = gpw.Context(0)
ctx.selftest_ff(512, seed=3)
logn = 14
g = torch.Generator(device="cuda").manual_seed(5)
a = torch.randint(0, 1 << 62, (1 << logn, 4), dtype=torch.int64, device="cuda", generator=g)
a[:, 3] &= (1 << 59) - 1
ref = a.clone()
for kw in (dict(inverse=False, out_bitrev=True), dict(inverse=True, in_bitrev=True),
           dict(inverse=False, coset=True, out_bitrev=True), dict(inverse=True, coset=True, in_bitrev=True)):
    ctx.ntt_dev(a.data_ptr(), logn, **kw)
ctx.sync()
# forward then inverse (plain and coset) is the identity on Montgomery residues up to the representative: compare canonically
back = gpw.limbs_to_ints(gpw.host_ff_from_mont(0, a.cpu().numpy().view(np.uint64)[:64]))
want = gpw.limbs_to_ints(gpw.host_ff_from_mont(0, ref.cpu().numpy().view(np.uint64)[:64]))
assert back == want, "NTT round trip"
n = 3000
pts = gpw.host_ec_generator_multiples(1, 1, n)
sc = gpw.ints_to_limbs([(i * 0x9E3779B97F4A7C15 + 1) % gpw.R_MOD if i % 3 else i % 65536 for i in range(n)])
ctx.msm(1, sc, pts)
for name in ("poseidon_gl", "qe_mul_div"):
    circ = gpw.Circuit.compile_gadget(ctx, name)
    key = gpw.WrapKey(ctx, circ, seed=11)
    n_in = circ.info["public"] + circ.info["secret"]
    if name == "qe_mul_div":
        from oracle import goldilocks as ogl
        from oracle.engine import Api
        ch = ogl.Chip(Api(trace=False))
        x, y = (3, 5), (7, 11)
        m = ch.MulExtension(x, y)
        d, _ = ch.DivExtension(x, y)
        inputs = circ.inputs_from_ints(list(m) + list(d), list(x) + list(y))
    else:
        from oracle.engine import Api
        from oracle.poseidon import GoldilocksChip
        st = list(range(1, 13))
        out = GoldilocksChip(Api(trace=False)).Poseidon(list(st))
        inputs = circ.inputs_from_ints([int(v) for v in out], st)
    p = key.prove(inputs, 3, 4)
    assert p["n_unsatisfied"] == 0, name
    key.close()
    circ.close()
ctx.close()
print("sanitize_small: ok")
