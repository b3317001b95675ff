#pragma once
// Batch-affine bucket accumulation: the first rounds of every MSM's bucket sums.
//
// gnark-crypto's MultiExp accumulates its buckets in AFFINE coordinates with a shared inversion ("batch affine", SURVEY
// A.3 item 6): P1 + P2 = (lambda^2 - x1 - x2, lambda (x1 - x3) - y1), lambda = (y2 - y1) / (x2 - x1), costs 3 field
// multiplications plus 3 for its share of one Montgomery-trick inversion - 6 instead of the 8 multiplications + 2 squarings
// + dual product (~9.5 multiplications' worth of IMAD.WIDE) of the extended-Jacobian mixed addition in k_msm_accumulate,
// and the kernel is bound by exactly that count (profiles/r01_msm_accumulate_ncu.md: IMAD pipe 88 %).
//
// On the GPU the additions of one bucket are made independent by a PAIRWISE TREE over the bucket-sorted entry list:
// round r turns a list in which bucket b owns k entries into one in which it owns ceil(k / 2) - neighbours (2i, 2i+1)
// are added, an odd last entry is carried over. The lists' bucket offsets off_r[b] = sum_{b' < b} ceil(k_b' / 2^r) follow
// from the sort's offsets by scans, so every output slot finds its two inputs without atomics, hot buckets (millions of
// entries for the bit wires of a witness) need no special casing, and from round 1 on the inputs are contiguous 64-byte
// points instead of gathers. After MSM_AFFINE_ROUNDS rounds (15/16 of all additions) the existing XYZZ kernel finishes
// the short remaining runs and writes the buckets.
//
// One thread owns PAIR_P consecutive output slots; a CTA shares ONE inversion per PAIR_P * 128 pairs:
//   inputs    gathered through a per-thread cp.async ring in shared memory (4 slots deep): the 64-byte gathers of the
//             next three slots are in flight while the current one is multiplied - the kernel is latency bound without it
//   forward   den_j = x2 - x1 (2 y1 for a doubling, 1 for copies / cancellations), prefix products in local memory
//   combine   the 128 per-thread totals are inverted together: warp-shuffle prefix / suffix products, one
//             shift-and-subtract inversion (ff.cuh inv_euclid: ALU pipe, no multiplications) by thread 0
//   backward  1 / den_j = u * prefix_j, u *= den_j, then the three multiplications of the addition itself.
#include "common.cuh"
#include "ec.cuh"

namespace gpw {

constexpr int PAIR_THREADS = 128;

template <class F>
struct PairCfg;
template <>
struct PairCfg<Fp> {
  static constexpr int P = 32, MIN_CTAS = 3;
};
template <>
struct PairCfg<Fp2> {
  static constexpr int P = 16, MIN_CTAS = 2;
};

__device__ __forceinline__ Fp shfl_up_f(const Fp& a, int d) {
  Fp r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.l[i] = __shfl_up_sync(0xffffffffu, a.l[i], d);
  return r;
}
__device__ __forceinline__ Fp shfl_down_f(const Fp& a, int d) {
  Fp r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.l[i] = __shfl_down_sync(0xffffffffu, a.l[i], d);
  return r;
}
__device__ __forceinline__ Fp2 shfl_up_f(const Fp2& a, int d) { return {shfl_up_f(a.c0, d), shfl_up_f(a.c1, d)}; }
__device__ __forceinline__ Fp2 shfl_down_f(const Fp2& a, int d) { return {shfl_down_f(a.c0, d), shfl_down_f(a.c1, d)}; }

__device__ __forceinline__ Fp inv_batch_total(const Fp& a) { return inv_euclid(a); }
__device__ __forceinline__ Fp2 inv_batch_total(const Fp2& a) {
  Fp n = add(sqr(a.c0), sqr(a.c1));
  Fp ni = inv_euclid(n);
  return {mul(a.c0, ni), neg(mul(a.c1, ni))};
}

template <class P>
__device__ __forceinline__ Fe<P> fsel_f(bool c, const Fe<P>& a, const Fe<P>& b) {
  Fe<P> r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.l[i] = c ? a.l[i] : b.l[i];
  return r;
}
__device__ __forceinline__ Fp2 fsel_f(bool c, const Fp2& a, const Fp2& b) { return {fsel_f(c, a.c0, b.c0), fsel_f(c, a.c1, b.c1)}; }

// counts_r[b] = ceil((offsets0[b+1] - offsets0[b]) / 2^r)
static __global__ void __launch_bounds__(256) k_msm_round_counts(const uint32_t* __restrict__ offsets0, uint32_t B, int r,
                                                                 uint32_t* __restrict__ counts) {
  const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const uint32_t k = offsets0[b + 1] - offsets0[b];
  counts[b] = (k + (1u << r) - 1u) >> r;
}

enum : uint32_t { PK_ADD = 0, PK_DBL = 1, PK_COPY1 = 2, PK_COPY2 = 3, PK_INF = 4 };

// denominator whose inverse the addition needs, and what kind of "addition" it is
template <class F>
__device__ __forceinline__ uint32_t pair_kind(const Affine<F>& p1, const Affine<F>& p2, bool valid, bool has2, F& den) {
  const F one = F::one();
  const F dx = sub(p2.x, p1.x);
  const bool i1 = p1.is_inf(), i2 = p2.is_inf();
  uint32_t kind = PK_ADD;
  den = dx;
  if (!valid || !has2 || i2) {
    kind = PK_COPY1;
    den = one;
  } else if (i1) {
    kind = PK_COPY2;
    den = one;
  } else if (dx.is_zero()) {
    if (p1.y == p2.y) {
      kind = PK_DBL;
      den = dbl(p1.y);
    } else {
      kind = PK_INF;
      den = one;
    }
  }
  return kind;
}

// ---- per-thread cp.async ring: the inputs of the next RING - 1 slots are in flight while the current one is used ------
// Layout ring[slot][chunk][thread] of 16-byte chunks (conflict-free LDS.128 / cp.async destinations): the two points of
// the slot and, on the way back, the slot's prefix product. Every thread reads only what it copied itself, so
// cp.async.wait_group is all the synchronisation needed.
template <class F>
struct PairRing {
  static constexpr int CH = (int)(sizeof(Affine<F>) / 16);  // chunks per point
  static constexpr int CHF = (int)(sizeof(F) / 16);         // chunks per field element (the prefix product)
  static constexpr int SLOT = 2 * CH + CHF;
  static constexpr int RING = sizeof(F) == sizeof(Fp) ? 3 : 2;  // 60 KB (3 CTAs per SM) / 80 KB (2 CTAs per SM)
  static constexpr size_t BYTES = (size_t)RING * SLOT * PAIR_THREADS * 16;
};

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

// Nothing of a thread's P slots lives in local memory: the bucket walk is redone (forwards, then backwards) when the
// copies of a slot are issued, per-slot flags are bit masks, and the prefix products go to a global scratch array
// (`pre`, coalesced: slot-major inside the CTA's block) and come back through the ring.
template <class F, bool FIRST>
__global__ void __launch_bounds__(PAIR_THREADS, PairCfg<F>::MIN_CTAS)
    k_msm_pair_round(const Affine<F>* __restrict__ in, const uint32_t* __restrict__ sorted, const uint32_t* __restrict__ off_in,
                     const uint32_t* __restrict__ off_out, uint32_t B, Affine<F>* __restrict__ out, F* __restrict__ pre) {
  constexpr int P = PairCfg<F>::P;
  constexpr int CH = PairRing<F>::CH, CHF = PairRing<F>::CHF, SLOT = PairRing<F>::SLOT, RING = PairRing<F>::RING;
  static_assert(P <= 32, "per-slot flags are 32-bit masks");
  extern __shared__ uint4 pair_smem[];
  __shared__ F sh_tot[PAIR_THREADS / 32];
  __shared__ F sh_k[PAIR_THREADS / 32];
  const uint32_t M_out = off_out[B];
  const uint64_t cta_o0 = (uint64_t)blockIdx.x * PAIR_THREADS * P;
  if (cta_o0 >= M_out) return;
  const uint32_t tid = threadIdx.x, lane = tid & 31u, wid = tid >> 5;
  const uint64_t o0_64 = cta_o0 + (uint64_t)tid * P;
  const bool active = o0_64 < M_out;
  const uint32_t o0 = active ? (uint32_t)o0_64 : M_out - 1u;
  const uint32_t n_valid = active ? min((uint32_t)P, M_out - o0) : 0u;
  F* const pre_cta = pre + (size_t)blockIdx.x * PAIR_THREADS * P + tid;  // slot j of this thread: pre_cta[j * PAIR_THREADS]
  // bucket of the first slot: largest b with off_out[b] <= o0
  uint32_t lo = 0, hi = B;
  while (hi - lo > 1) {
    const uint32_t mid = (lo + hi) >> 1;
    if (off_out[mid] <= o0) lo = mid; else hi = mid;
  }
  uint32_t b = lo;
  uint32_t ob_start = off_out[b], ob_end = off_out[b + 1], ib_start = off_in[b], ib_end = off_in[b + 1];
  uint32_t has2_mask = 0, neg1_mask = 0, neg2_mask = 0;

  // start the copies of slot j's inputs into ring slot j % RING. dir = +1: slots are issued in increasing order (the walk
  // moves to later buckets), -1: in decreasing order.
  auto issue = [&](int j, int dir) {
    const bool valid = (uint32_t)j < n_valid;
    const uint32_t o = o0 + (valid ? (uint32_t)j : 0u);
    if (valid) {
      if (dir > 0) {
        while (o >= ob_end) {  // next non-empty bucket
          b++;
          ob_start = ob_end;
          ob_end = off_out[b + 1];
          ib_start = ib_end;
          ib_end = off_in[b + 1];
        }
      } else {
        while (o < ob_start) {  // previous non-empty bucket
          b--;
          ob_end = ob_start;
          ob_start = off_out[b];
          ib_end = ib_start;
          ib_start = off_in[b];
        }
      }
    }
    // (slots past the end of the list read entry 0, which exists whenever the list is not empty, and are never stored)
    const uint32_t a = valid ? ib_start + 2u * (o - ob_start) : 0u;
    const bool has2 = valid && (a + 1u < ib_end);
    uint32_t i1 = a, i2 = a + 1u;
    if (FIRST) {
      const uint32_t e1 = sorted[a];
      i1 = e1 & 0x7fffffffu;
      if (dir > 0) neg1_mask |= (e1 >> 31) << j;
      if (has2) {
        const uint32_t e2 = sorted[a + 1];
        i2 = e2 & 0x7fffffffu;
        if (dir > 0) neg2_mask |= (e2 >> 31) << j;
      }
    }
    if (dir > 0) has2_mask |= (has2 ? 1u : 0u) << j;
    uint4* dst = pair_smem + (size_t)(j % RING) * SLOT * PAIR_THREADS + tid;
    const uint4* g1 = reinterpret_cast<const uint4*>(in + i1);
#pragma unroll
    for (int c = 0; c < CH; c++) cp_async16(dst + c * PAIR_THREADS, g1 + c);
    if (has2) {
      const uint4* g2 = reinterpret_cast<const uint4*>(in + i2);
#pragma unroll
      for (int c = 0; c < CH; c++) cp_async16(dst + (CH + c) * PAIR_THREADS, g2 + c);
    }
    if (dir < 0) {
      const uint4* gp = reinterpret_cast<const uint4*>(pre_cta + (size_t)j * PAIR_THREADS);
#pragma unroll
      for (int c = 0; c < CHF; c++) cp_async16(dst + (2 * CH + c) * PAIR_THREADS, gp + c);
    }
  };
  auto fetch = [&](int j, Affine<F>& p1, Affine<F>& p2) {  // slot j's inputs out of the ring (signs applied)
    const uint4* src = pair_smem + (size_t)(j % RING) * SLOT * PAIR_THREADS + tid;
    uint4* d1 = reinterpret_cast<uint4*>(&p1);
    uint4* d2 = reinterpret_cast<uint4*>(&p2);
#pragma unroll
    for (int c = 0; c < CH; c++) d1[c] = src[c * PAIR_THREADS];
    const bool has2 = (has2_mask >> j) & 1u;
    if (has2) {
#pragma unroll
      for (int c = 0; c < CH; c++) d2[c] = src[(CH + c) * PAIR_THREADS];
    }
    if (FIRST) {
      if ((neg1_mask >> j) & 1u) p1.y = neg(p1.y);
      if (has2 && ((neg2_mask >> j) & 1u)) p2.y = neg(p2.y);
    }
    if (!has2) p2 = p1;
  };
  // ---- forward: denominators and their prefix products ---------------------------------------------------------------------
  F run = F::one();
#pragma unroll
  for (int j = 0; j < RING - 1; j++) {
    issue(j, +1);
    cp_async_commit();
  }
#pragma unroll 1
  for (int j = 0; j < P; j++) {
    if (j + RING - 1 < P) issue(j + RING - 1, +1);
    cp_async_commit();
    cp_async_wait<RING - 1>();
    const bool valid = (uint32_t)j < n_valid;
    Affine<F> p1, p2;
    fetch(j, p1, p2);
    F den;
    pair_kind(p1, p2, valid, (has2_mask >> j) & 1u, den);
    st_struct(pre_cta + (size_t)j * PAIR_THREADS, run);
    run = mul(run, den);
  }
  // ---- the CTA's 128 totals are inverted together ---------------------------------------------------------------------
  F incl = run;  // inclusive prefix product over the lanes of the warp
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const F y = shfl_up_f(incl, d);
    const F t = mul(incl, y);
    incl = fsel_f((int)lane >= d, t, incl);
  }
  F sfx = run;  // inclusive suffix product
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const F y = shfl_down_f(sfx, d);
    const F t = mul(sfx, y);
    sfx = fsel_f((int)lane + d < 32, t, sfx);
  }
  if (lane == 31) sh_tot[wid] = incl;
  // the backward pass's first inputs travel while the inversion runs (a thread's own prefix stores are visible to its
  // own later copies: same thread, program order)
#pragma unroll
  for (int j = P - 1; j > P - RING; j--) {
    issue(j, -1);
    cp_async_commit();
  }
  __syncthreads();
  if (tid == 0) {
    constexpr int NW = PAIR_THREADS / 32;
    F tot[NW];
#pragma unroll
    for (int w = 0; w < NW; w++) tot[w] = sh_tot[w];
    F g = tot[0];
#pragma unroll
    for (int w = 1; w < NW; w++) g = mul(g, tot[w]);
    const F ginv = inv_batch_total(g);
    // k_w = ginv * prod_{w' != w} tot_w'
    F left = F::one();
#pragma unroll
    for (int w = 0; w < NW; w++) {
      F right = F::one();
#pragma unroll
      for (int w2 = NW - 1; w2 > w; w2--) right = mul(right, tot[w2]);
      sh_k[w] = mul(ginv, mul(left, right));
      left = mul(left, tot[w]);
    }
  }
  __syncthreads();
  // 1 / (this thread's total) = k_w * (product of the lanes before) * (product of the lanes after)
  F ex_pre = shfl_up_f(incl, 1), ex_sfx = shfl_down_f(sfx, 1);
  ex_pre = fsel_f(lane == 0, F::one(), ex_pre);
  ex_sfx = fsel_f(lane == 31, F::one(), ex_sfx);
  F u = mul(sh_k[wid], mul(ex_pre, ex_sfx));
  // ---- backward: the additions ---------------------------------------------------------------------------------------------
#pragma unroll 1
  for (int j = P - 1; j >= 0; j--) {
    if (j - (RING - 1) >= 0) issue(j - (RING - 1), -1);
    cp_async_commit();
    cp_async_wait<RING - 1>();
    const bool valid = (uint32_t)j < n_valid;
    const bool has2 = (has2_mask >> j) & 1u;
    Affine<F> p1, p2;
    fetch(j, p1, p2);
    F pj;
    {
      const uint4* src = pair_smem + (size_t)(j % RING) * SLOT * PAIR_THREADS + tid;
      uint4* d = reinterpret_cast<uint4*>(&pj);
#pragma unroll
      for (int c = 0; c < CHF; c++) d[c] = src[(2 * CH + c) * PAIR_THREADS];
    }
    F den;
    const uint32_t kind = pair_kind(p1, p2, valid, has2, den);
    const F dinv = mul(u, pj);
    u = mul(u, den);
    F num = sub(p2.y, p1.y);
    if (kind == PK_DBL) {  // rare: equal points in one bucket
      const F xx = sqr(p1.x);
      num = add(dbl(xx), xx);
    }
    const F lam = mul(num, dinv);
    const F x3 = sub(sub(sqr(lam), p1.x), p2.x);
    const F y3 = sub(mul(lam, sub(p1.x, x3)), p1.y);
    Affine<F> r;
    const bool is_add = kind <= PK_DBL;
    r.x = fsel_f(is_add, x3, kind == PK_COPY2 ? p2.x : p1.x);
    r.y = fsel_f(is_add, y3, kind == PK_COPY2 ? p2.y : p1.y);
    if (kind == PK_INF) r = Affine<F>{F::zero(), F::zero()};
    if (valid) st_struct(out + o0 + j, r);
  }
}

}  // namespace gpw
