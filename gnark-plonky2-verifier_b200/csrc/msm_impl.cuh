#pragma once
// K9 - Pippenger multi-scalar multiplication over BN254 G1 / G2 for sm_100a.
//
// Replaces gnark-crypto's G1Jac.MultiExp / G2Jac.MultiExp (un-vendored dependency of the reference,
// reached from groth16.Prove at benchmark.go:249). Output is the unique group element, so parity
// with any correct implementation is exact.
//
// Pipeline (all on one stream, no host round trips until the final 16 window sums):
//   1. k_msm_digits<0> signed c-bit digits of every scalar -> histogram over (window, |digit|) buckets
//   2. k_msm_scan      exclusive prefix sum of the histogram (bucket start offsets)
//   3. k_msm_digits<1> counting-sort scatter of (point index | sign) by bucket (both passes warp-aggregate their atomics)
//   4. k_msm_accumulate  THE hot kernel: the sorted entry array is cut into uniform tasks of
//                      MSM_TASK entries; one thread walks one task doing XYZZ += affine mixed adds
//                      (8 multiplications + the dual-product Y3: ~1290 IMAD.WIDE) with 64 B / 128 B gathers of the
//                      affine points, in ONE flat loop so the warp stays converged. Uniform tasks make the work per
//                      thread identical even for the heavily skewed scalar distribution of a gnark witness (most
//                      wires are 0/1 or < 2^64), where thread-per-bucket schemes serialise on the hot buckets.
//   5. k_msm_fixup(_big)  buckets that span several tasks: their partials follow from the bucket offsets alone
//   6. k_msm_window_partial (two levels) / k_msm_window_final   sum_b (b+1) * B[w][b] per window
//   7. host: Horner over <= 16 window sums (270 group ops) and one inversion to affine.
// Zero digits are skipped entirely, so a scalar < 2^64 costs 4-5 adds instead of 16.
// Variants: fixed-base mode (precomputed 2^(c w) P_i, one bucket set, msm_dev_impl's fixed_windows) and shared sorts
// (several MSMs over the same scalars, sort_tag / reuse_sort).
#include "common.cuh"
#include "ec.cuh"

namespace gpw {

constexpr int MSM_TASK = 64;         // sorted entries per accumulate thread

template <class T>
__device__ __forceinline__ T ld_struct(const T* p) {
  static_assert(sizeof(T) % 16 == 0, "16-byte multiple");
  T r;
  const uint4* s = reinterpret_cast<const uint4*>(p);
  uint4* d = reinterpret_cast<uint4*>(&r);
#pragma unroll
  for (int i = 0; i < (int)(sizeof(T) / 16); i++) d[i] = __ldg(s + i);
  return r;
}

template <class T>
__device__ __forceinline__ void st_struct(T* p, const T& v) {
  const uint4* s = reinterpret_cast<const uint4*>(&v);
  uint4* d = reinterpret_cast<uint4*>(p);
#pragma unroll
  for (int i = 0; i < (int)(sizeof(T) / 16); i++) d[i] = s[i];
}

__device__ __forceinline__ uint32_t get_bits(const uint32_t* s, int lo, int c) {
  int w = lo >> 5, off = lo & 31;
  if (w >= 8) return 0;
  uint64_t v = s[w];
  if (w + 1 < 8) v |= (uint64_t)s[w + 1] << 32;
  return (uint32_t)(v >> off) & ((1u << c) - 1u);
}

}  // namespace gpw
#include "msm_affine.cuh"
namespace gpw {

// Signed-digit extraction shared by the histogram and the scatter pass. Every lane walks every window (uniform trip
// count) so that the warp can aggregate its atomics: lanes that hit the same bucket - bit wires and small constants make
// (window 0, digit 1) and its neighbours receive millions of entries - are found with match.any and served by ONE
// atomic of the group's size instead of a same-address atomic per lane (which the L2 serialises).
// fixed != 0 (fixed-base mode, see msm_dev_impl): all windows share ONE bucket set and digit w of scalar i refers to
// the precomputed point 2^(c w) P_i stored at index w n + i.
template <bool SCATTER>
static __global__ void __launch_bounds__(256) k_msm_digits(const Fr* __restrict__ scalars, size_t n, int mont, int c, int nwin,
                                                           int win_lo, int win_hi, int fixed, uint32_t* __restrict__ counters,
                                                           uint32_t* __restrict__ sorted) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t lane = threadIdx.x & 31u;
  Fr s = Fr::zero();
  if (i < n) {
    s = ld_struct(scalars + i);
    if (mont) s = from_mont(s);
  }
  const uint32_t half = 1u << (c - 1);
  const int wend = win_hi < nwin ? win_hi : nwin;
  uint32_t carry = 0;
  for (int w = 0; w < wend; w++) {
    const uint32_t raw = get_bits(s.l, w * c, c) + carry;
    const bool negv = raw > half;
    const uint32_t mag = negv ? ((1u << c) - raw) : raw;
    carry = negv ? 1u : 0u;
    const bool has = mag != 0 && w >= win_lo;
    if (!__any_sync(0xffffffffu, has)) continue;
    const uint32_t key = has ? (fixed ? 0u : (uint32_t)(w - win_lo) * half) + (mag - 1u) : (0xffffffffu - lane);
    const uint32_t peers = __match_any_sync(0xffffffffu, key);
    if (has) {
      const uint32_t leader = (uint32_t)__ffs(peers) - 1u;
      const uint32_t cnt = (uint32_t)__popc(peers);
      if (!SCATTER) {
        if (lane == leader) atomicAdd(&counters[key], cnt);
      } else {
        uint32_t base = 0;
        if (lane == leader) base = atomicAdd(&counters[key], cnt);
        base = __shfl_sync(peers, base, leader);
        const uint32_t rank = (uint32_t)__popc(peers & ((1u << lane) - 1u));
        sorted[base + rank] = (uint32_t)(fixed ? (size_t)w * n + i : i) | (negv ? 0x80000000u : 0u);
      }
    }
  }
}

// exclusive scan of B counters -> offsets[0..B], cursor[0..B): per-block sums, scan of the block sums,
// per-block exclusive scan + block offset. SCAN_BLOCK counters per block.
constexpr uint32_t SCAN_THREADS = 256;
constexpr uint32_t SCAN_PER_THREAD = 8;
constexpr uint32_t SCAN_BLOCK = SCAN_THREADS * SCAN_PER_THREAD;

static __device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* sh, uint32_t& total) {
  const uint32_t tid = threadIdx.x, lane = tid & 31u, wid = tid >> 5;
  uint32_t x = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    uint32_t y = __shfl_up_sync(0xffffffffu, x, d);
    if ((int)lane >= d) x += y;
  }
  if (lane == 31) sh[wid] = x;
  __syncthreads();
  if (wid == 0) {
    uint32_t w = lane < (blockDim.x >> 5) ? sh[lane] : 0u;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      uint32_t y = __shfl_up_sync(0xffffffffu, w, d);
      if ((int)lane >= d) w += y;
    }
    sh[lane] = w;  // inclusive scan of warp totals
  }
  __syncthreads();
  uint32_t warp_off = wid ? sh[wid - 1] : 0u;
  total = sh[(blockDim.x >> 5) - 1];
  return warp_off + x - v;
}

static __global__ void __launch_bounds__(SCAN_THREADS) k_msm_scan_sums(const uint32_t* __restrict__ counts, uint32_t B,
                                                                        uint32_t* __restrict__ block_sums) {
  __shared__ uint32_t sh[32];
  uint32_t base = blockIdx.x * SCAN_BLOCK + threadIdx.x * SCAN_PER_THREAD;
  uint32_t s = 0;
#pragma unroll
  for (uint32_t k = 0; k < SCAN_PER_THREAD; k++)
    if (base + k < B) s += counts[base + k];
  uint32_t total;
  block_exclusive_scan(s, sh, total);
  if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

// single block: exclusive scan of nblocks block sums in place (nblocks <= SCAN_BLOCK * 64)
static __global__ void __launch_bounds__(SCAN_THREADS) k_msm_scan_blocks(uint32_t* __restrict__ block_sums, uint32_t nblocks,
                                                                          uint32_t* __restrict__ total_out) {
  __shared__ uint32_t sh[32];
  __shared__ uint32_t carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (uint32_t base = 0; base < nblocks; base += SCAN_THREADS) {
    uint32_t i = base + threadIdx.x;
    uint32_t v = i < nblocks ? block_sums[i] : 0u;
    uint32_t total;
    uint32_t ex = block_exclusive_scan(v, sh, total);
    uint32_t c = carry;
    if (i < nblocks) block_sums[i] = c + ex;
    __syncthreads();
    if (threadIdx.x == 0) carry = c + total;
    __syncthreads();
  }
  if (threadIdx.x == 0) *total_out = carry;
}

static __global__ void __launch_bounds__(SCAN_THREADS) k_msm_scan_final(const uint32_t* __restrict__ counts, uint32_t B,
                                                                         const uint32_t* __restrict__ block_sums,
                                                                         uint32_t* __restrict__ offsets,
                                                                         uint32_t* __restrict__ cursor) {
  __shared__ uint32_t sh[32];
  uint32_t base = blockIdx.x * SCAN_BLOCK + threadIdx.x * SCAN_PER_THREAD;
  uint32_t v[SCAN_PER_THREAD];
  uint32_t s = 0;
#pragma unroll
  for (uint32_t k = 0; k < SCAN_PER_THREAD; k++) {
    v[k] = (base + k < B) ? counts[base + k] : 0u;
    s += v[k];
  }
  uint32_t total;
  uint32_t run = block_sums[blockIdx.x] + block_exclusive_scan(s, sh, total);
#pragma unroll
  for (uint32_t k = 0; k < SCAN_PER_THREAD; k++) {
    if (base + k < B) {
      offsets[base + k] = run;
      if (cursor) cursor[base + k] = run;
    }
    run += v[k];
  }
}

// select helpers: branch-free so that the lanes of a warp stay converged through the accumulate loop
template <class P>
__device__ __forceinline__ Fe<P> fsel(bool c, const Fe<P>& a, const Fe<P>& b) {
  Fe<P> r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.l[i] = c ? a.l[i] : b.l[i];
  return r;
}
__device__ __forceinline__ Fp2 fsel(bool c, const Fp2& a, const Fp2& b) { return {fsel(c, a.c0, b.c0), fsel(c, a.c1, b.c1)}; }

// acc += (q.x, negate ? -q.y : q.y), written for SIMT: the common cases (ordinary addition, first point of a run,
// base at infinity) run the same instruction stream and are resolved by selects; only the doubling / cancellation
// case (acc == +-q, never hit by honest keys but reachable from tests) leaves the straight-line path.
template <class F>
__device__ __forceinline__ void add_mixed_flat(XYZZ<F>& acc, const Affine<F>& q, bool negate) {
  const bool q_inf = q.is_inf();
  const bool a_inf = acc.is_inf();
  const F y2 = fsel(negate, neg(q.y), q.y);
  const F U2 = mul(q.x, acc.ZZ);
  const F S2 = mul(y2, acc.ZZZ);
  const F Pv = sub(U2, acc.X);
  const F R = sub(S2, acc.Y);
  if (!q_inf && !a_inf && Pv.is_zero()) {
    if (R.is_zero()) acc = dbl_affine(Affine<F>{q.x, y2});
    else acc = XYZZ<F>::inf();
    return;
  }
  const F PP = sqr(Pv);
  const F PPP = mul(Pv, PP);
  const F Q = mul(acc.X, PP);
  const F X3 = sub(sub(sqr(R), PPP), dbl(Q));
  const F Y3 = mul_sub2(R, sub(Q, X3), acc.Y, PPP);  // G1: both products under one Montgomery reduction
  const F Z2 = mul(acc.ZZ, PP);
  const F Z3 = mul(acc.ZZZ, PPP);
  // a_inf: the sum is q itself; q_inf: acc unchanged
  const F one = F::one();
  acc.X = fsel(q_inf, acc.X, fsel(a_inf, q.x, X3));
  acc.Y = fsel(q_inf, acc.Y, fsel(a_inf, y2, Y3));
  acc.ZZ = fsel(q_inf, acc.ZZ, fsel(a_inf, one, Z2));
  acc.ZZZ = fsel(q_inf, acc.ZZZ, fsel(a_inf, one, Z3));
}

constexpr int MSM_ACC_THREADS = 128;
#ifndef MSM_ACC_MIN_CTAS
#define MSM_ACC_MIN_CTAS 3  /* 4 (a 128-register cap) was measured: no faster, the kernel is IMAD-pipe bound, not occupancy bound */
#endif
#ifndef MSM_ACC_MIN_CTAS_G2
#define MSM_ACC_MIN_CTAS_G2 2  /* G2: 254 registers; 3 CTAs per SM (a 168-register cap, 624 bytes of spills) was measured: 10.0 instead of 9.06 ms for B2 */
#endif

// One thread per task of MSM_TASK consecutive sorted entries, walked by ONE flat loop so that all lanes of a warp
// execute the same mixed addition in lock step (the earlier nested per-run loops left ~55 % of the lanes idle:
// smsp__thread_inst_executed_per_inst_executed 14.4, profiles/r01_msm_accumulate_ncu.md). A run that ends inside
// the task is flushed by a short divergent store. CTAs whose 8192 entries all belong to one (hot) bucket - bit
// wires make bucket (window 0, digit 1) millions of entries long - merge their 128 partials in shared memory, so
// the fix-up pass sees one partial per CTA instead of one per thread.
template <class F>
__global__ void __launch_bounds__(MSM_ACC_THREADS, sizeof(F) == sizeof(Fp) ? MSM_ACC_MIN_CTAS : MSM_ACC_MIN_CTAS_G2)
    k_msm_accumulate(const Affine<F>* __restrict__ points, const uint32_t* __restrict__ sorted,
                     const uint32_t* __restrict__ offsets, uint32_t B, XYZZ<F>* __restrict__ buckets,
                     XYZZ<F>* __restrict__ head, XYZZ<F>* __restrict__ tail, uint32_t* __restrict__ tail_key,
                     uint32_t* __restrict__ tail_list, uint32_t* __restrict__ ntail) {
  extern __shared__ uint4 acc_smem[];
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t M = offsets[B];
  const uint64_t cta_p0 = (uint64_t)blockIdx.x * blockDim.x * MSM_TASK;
  if (cta_p0 >= M) return;  // whole CTA past the end
  const uint64_t cta_p1 = cta_p0 + (uint64_t)blockDim.x * MSM_TASK;
  const uint64_t p0 = (uint64_t)t * MSM_TASK;
  const bool active = p0 < M;
  const uint32_t pos0 = active ? (uint32_t)p0 : M - 1;
  const uint32_t pos1 = active ? min(pos0 + (uint32_t)MSM_TASK, M) : M - 1;
  // bucket containing pos0: largest b with offsets[b] <= pos0 (skips empty buckets sharing the offset)
  uint32_t lo = 0, hi = B;
  while (hi - lo > 1) {
    uint32_t mid = (lo + hi) >> 1;
    if (offsets[mid] <= pos0) lo = mid; else hi = mid;
  }
  uint32_t b = lo;
  uint32_t bstart = offsets[b], bend = offsets[b + 1];
  // the same value in every thread: if one thread's bucket covers the CTA's whole range, all threads are in it
  const bool cta_uniform = cta_p1 <= M && bstart <= cta_p0 && bend >= cta_p1;

  XYZZ<F> acc = XYZZ<F>::inf();
  uint32_t pos = pos0;
  // sorted == nullptr: the entries ARE the points (the list left by the batch-affine rounds), in order, no signs
  uint32_t e_next = sorted ? sorted[pos0] : pos0;
  Affine<F> p_next = ld_struct(points + (e_next & 0x7fffffffu));
#pragma unroll 1
  while (pos < pos1) {
    if (pos == bend) {  // the previous entry closed bucket b (never taken in a uniform CTA)
      st_struct(bstart < pos0 ? head + t : buckets + b, acc);
      acc = XYZZ<F>::inf();
      b++;
      while (offsets[b + 1] <= pos) b++;
      bstart = pos;
      bend = offsets[b + 1];
    }
    const uint32_t e = e_next;
    const Affine<F> p = p_next;
    pos++;
    if (pos < pos1) {  // prefetch the next point while this add runs (dropping the prefetch for G2, whose addition
                       // is register bound, was measured: no faster)
      e_next = sorted ? sorted[pos] : pos;
      p_next = ld_struct(points + (e_next & 0x7fffffffu));
    }
    add_mixed_flat(acc, p, (e >> 31) != 0);
  }

  if (cta_uniform) {
    XYZZ<F>* sm = reinterpret_cast<XYZZ<F>*>(acc_smem);
    const uint32_t tid = threadIdx.x;
    sm[tid] = acc;
    __syncthreads();
    for (uint32_t d = blockDim.x >> 1; d >= 1; d >>= 1) {
      if (tid < d) {
        XYZZ<F> o = sm[tid + d];
        add_full(acc, o);
        sm[tid] = acc;
      }
      __syncthreads();
    }
    // thread 0 carries the CTA's sum: the fix-up pass knows from the bucket's offsets which CTAs merged (those lying
    // wholly inside the bucket) and reads one partial per such CTA - slot of its thread 0 - skipping the other 127
    if (tid != 0) return;
    const bool starts_here = bstart == (uint32_t)cta_p0, ends_here = bend == (uint32_t)cta_p1;
    if (starts_here && ends_here) {  // the bucket is exactly this CTA
      st_struct(buckets + b, acc);
    } else if (starts_here) {
      st_struct(tail + t, acc);
      tail_key[t] = b;
      tail_list[atomicAdd(ntail, 1u)] = t;
    } else {
      st_struct(head + t, acc);
    }
    return;
  }
  if (!active) return;
  // last run of the task (ends at pos1 or beyond)
  if (bstart < pos0) {  // bucket began in an earlier task
    st_struct(head + t, acc);
  } else if (bend > pos1) {  // bucket continues into later tasks
    st_struct(tail + t, acc);
    tail_key[t] = b;
    tail_list[atomicAdd(ntail, 1u)] = t;
  } else {
    st_struct(buckets + b, acc);
  }
}

// The partial sums of a bucket that spans several tasks: tail[t0] (the task the bucket starts in) plus one "item" per
// later task up to the task holding the bucket's last entry - except that CTAs lying wholly inside the bucket merged
// their 128 partials into the slot of their first thread. Everything follows from the bucket's offsets, so the
// accumulate kernel does not have to publish per-task keys.
struct SpanItems {
  uint32_t t0, pre, mid, first_u, qb, total;
  __device__ __forceinline__ SpanItems(uint32_t t0_, uint32_t bstart, uint32_t bend) : t0(t0_) {
    constexpr uint32_t CTA_E = (uint32_t)MSM_ACC_THREADS * MSM_TASK;
    const uint32_t t_last = (bend - 1u) / MSM_TASK;
    const uint32_t qa = (bstart + CTA_E - 1u) / CTA_E;
    qb = bend / CTA_E;
    if (qa < qb) {
      first_u = qa;
      if (t0 == qa * MSM_ACC_THREADS) {  // the bucket starts exactly at CTA qa: its merged sum IS tail[t0]
        first_u = qa + 1u;
        pre = 0;
      } else {
        pre = qa * MSM_ACC_THREADS - (t0 + 1u);
      }
      mid = qb - first_u;
      const uint32_t post = t_last + 1u - qb * MSM_ACC_THREADS;  // t_last >= qb * 128 - 1 always
      total = pre + mid + post;
    } else {
      first_u = qb = 0;
      pre = t_last - t0;
      mid = 0;
      total = pre;
    }
  }
  // task slot (index into head[]) of item j
  __device__ __forceinline__ uint32_t slot(uint32_t j) const {
    if (j < pre) return t0 + 1u + j;
    if (j < pre + mid) return (first_u + (j - pre)) * MSM_ACC_THREADS;
    return qb * MSM_ACC_THREADS + (j - pre - mid);
  }
};

constexpr uint32_t FIXUP_SERIAL_MAX = 24;  // spans up to this many tasks are summed by one thread

// Buckets that span several tasks. One thread per bucket; the rare long spans (hot buckets of a skewed scalar
// distribution) are deferred to the CTA kernel below.
template <class F>
__global__ void __launch_bounds__(128)
    k_msm_fixup(const XYZZ<F>* __restrict__ head, const XYZZ<F>* __restrict__ tail, const uint32_t* __restrict__ offsets,
                const uint32_t* __restrict__ tail_key, const uint32_t* __restrict__ tail_list,
                const uint32_t* __restrict__ ntail, uint32_t* __restrict__ big_list, uint32_t* __restrict__ nbig,
                XYZZ<F>* __restrict__ buckets) {
  const uint32_t nt = *ntail;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < nt; i += gridDim.x * blockDim.x) {
    const uint32_t t0 = tail_list[i];
    const uint32_t key = tail_key[t0];
    const SpanItems it(t0, offsets[key], offsets[key + 1]);
    if (it.total > FIXUP_SERIAL_MAX) {
      big_list[atomicAdd(nbig, 1u)] = t0;
      continue;
    }
    XYZZ<F> acc = ld_struct(tail + t0);
    for (uint32_t j = 0; j < it.total; j++) {
      XYZZ<F> h = ld_struct(head + it.slot(j));
      add_full(acc, h);
    }
    st_struct(buckets + key, acc);
  }
}

// one CTA per long-span bucket: strided serial sums per thread, then a shared-memory tree
template <class F>
__global__ void __launch_bounds__(256)
    k_msm_fixup_big(const XYZZ<F>* __restrict__ head, const XYZZ<F>* __restrict__ tail, const uint32_t* __restrict__ offsets,
                    const uint32_t* __restrict__ tail_key, const uint32_t* __restrict__ big_list,
                    const uint32_t* __restrict__ nbig, XYZZ<F>* __restrict__ buckets) {
  extern __shared__ uint4 fix_smem[];
  XYZZ<F>* sm = reinterpret_cast<XYZZ<F>*>(fix_smem);
  const uint32_t nt = *nbig;
  for (uint32_t i = blockIdx.x; i < nt; i += gridDim.x) {
    const uint32_t t0 = big_list[i];
    const uint32_t key = tail_key[t0];
    const SpanItems it(t0, offsets[key], offsets[key + 1]);
    XYZZ<F> acc = XYZZ<F>::inf();
    if (threadIdx.x == 0) acc = ld_struct(tail + t0);
    for (uint32_t j = threadIdx.x; j < it.total; j += blockDim.x) {
      XYZZ<F> h = ld_struct(head + it.slot(j));
      add_full(acc, h);
    }
    st_struct(sm + threadIdx.x, acc);
    __syncthreads();
    for (uint32_t d = blockDim.x >> 1; d >= 1; d >>= 1) {
      if (threadIdx.x < d) {
        XYZZ<F> o = sm[threadIdx.x + d];
        add_full(acc, o);
        sm[threadIdx.x] = acc;
      }
      __syncthreads();
    }
    if (threadIdx.x == 0) st_struct(buckets + key, acc);
    __syncthreads();
  }
}

// Window reduction sum_b (b + 1) B[w][b], two levels deep. Level 0 (chunk_sums != nullptr): one thread per chunk of
// `chunk` buckets computes with the running-sum trick T = sum_j (j + 1) B[lo + j] -> partials and the plain chunk sum
// S = sum_j B[lo + j] -> chunk_sums; what is still owed, lo * S with lo = chunk * ci, is chunk * sum_ci ci * S_ci: the
// same problem on 16x fewer elements with weights ci instead of ci + 1. Level 1 (chunk_sums == nullptr, weight_off = 0)
// runs this kernel again on the chunk sums and pays the small scalar multiplication by its chunk offset there - on
// 1/256 of the elements - instead of once per level-0 chunk, which used to be ~45 % of the reduction's work.
// partial = sum_j (lo + j + weight_off) E[lo + j] when chunk_sums == nullptr, sum_j (j + weight_off) E[lo + j] otherwise.
template <class F>
__global__ void __launch_bounds__(128)
    k_msm_window_partial(const XYZZ<F>* __restrict__ buckets, const uint32_t* __restrict__ offsets, uint32_t half,
                         uint32_t chunk, uint32_t nwin, uint32_t weight_off, XYZZ<F>* __restrict__ partials,
                         XYZZ<F>* __restrict__ chunk_sums) {
  const uint32_t gid = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t nchunks = half / chunk;
  const uint32_t w = gid / nchunks, ci = gid % nchunks;
  if (w >= nwin) return;
  const uint32_t lo = ci * chunk;
  XYZZ<F> acc = XYZZ<F>::inf(), sum = XYZZ<F>::inf();
  // chunks without a single entry (the upper windows of an MSM over 16-bit limbs or 64-bit values) cost nothing
  if (offsets && offsets[w * half + lo] == offsets[w * half + lo + chunk]) {
    st_struct(partials + gid, sum);
    if (chunk_sums) st_struct(chunk_sums + gid, acc);
    return;
  }
  for (int b = (int)(lo + chunk) - 1; b >= (int)lo; b--) {
    XYZZ<F> bk = buckets[(size_t)w * half + b];
    add_full(acc, bk);
    add_full(sum, acc);
  }
  // sum = sum_j (j + 1) E[lo + j], acc = sum_j E[lo + j]
  if (weight_off == 0 && !acc.is_inf()) {
    XYZZ<F> t = neg(acc);
    add_full(sum, t);
  }
  if (chunk_sums) {
    st_struct(chunk_sums + gid, acc);
  } else if (lo > 0 && !acc.is_inf()) {
    XYZZ<F> t = mul_small(acc, lo);
    add_full(sum, t);
  }
  st_struct(partials + gid, sum);
}

// out[w * gridDim.x + g] = sum of partials[w][g * per_cta .. (g + 1) * per_cta)   (grid: (groups, windows))
constexpr uint32_t WINDOW_FINAL_PER_CTA = 2048;
template <class F>
__global__ void __launch_bounds__(128)
    k_msm_window_final(const XYZZ<F>* __restrict__ partials, uint32_t nchunks, uint32_t per_cta, XYZZ<F>* __restrict__ window_sums) {
  extern __shared__ uint4 smem_raw[];
  XYZZ<F>* sm = reinterpret_cast<XYZZ<F>*>(smem_raw);
  const uint32_t w = blockIdx.y, tid = threadIdx.x;
  const uint32_t lo = blockIdx.x * per_cta, hi = min(nchunks, lo + per_cta);
  XYZZ<F> acc = XYZZ<F>::inf();
  for (uint32_t i = lo + tid; i < hi; i += blockDim.x) {
    XYZZ<F> p = partials[(size_t)w * nchunks + i];
    add_full(acc, p);
  }
  sm[tid] = acc;
  __syncthreads();
  for (uint32_t d = blockDim.x >> 1; d >= 1; d >>= 1) {
    if (tid < d) {
      XYZZ<F> o = sm[tid + d];
      add_full(acc, o);
      sm[tid] = acc;
    }
    __syncthreads();
  }
  if (tid == 0) st_struct(window_sums + (size_t)w * gridDim.x + blockIdx.x, acc);
}

static int choose_window(size_t n) {
  int lg = 0;
  while ((1ull << (lg + 1)) <= n) lg++;
  int c = lg - 3;
  if (c > 16) c = 16;
  if (c < 4) c = 4;
  return c;
}

// Second half of an MSM: waits for its window sums (pinned memory) and folds them on the host,
// R = sum_w 2^(c w) (T_w + chunk U_w) with T_w / U_w the level-0 / level-1 sums of window w (k_msm_window_partial).
template <class F>
static int msm_finish_impl(gpw_ctx* ctx, gpw_ctx::MsmPending& P) {
  if (!P.open) return GPW_OK;
  P.open = false;
  GPW_CUDA(cudaEventSynchronize(P.ev[3]));
  GPW_CUDA(cudaEventElapsedTime(&P.acc_ms, P.ev[1], P.ev[2]));
  GPW_CUDA(cudaEventElapsedTime(&P.total_ms, P.ev[0], P.ev[3]));
  const XYZZ<F>* hw = (const XYZZ<F>*)P.hw;
  const int nw = P.nw, c = P.c;
  P.digits = *P.Mp;
  ctx->msm_acc_ms = P.acc_ms;
  ctx->msm_total_ms = P.total_ms;
  ctx->msm_digits = P.digits;
  {
    const int g = sizeof(Affine<F>) == 64 ? 0 : 1;
    ctx->msm_acc_ms_sum[g] += P.acc_ms;
    ctx->msm_total_ms_sum[g] += P.total_ms;
    ctx->msm_points_sum[g] += P.n;
    ctx->msm_digits_sum[g] += P.digits;
    ctx->msm_calls[g] += 1;
  }
  XYZZ<F> R = XYZZ<F>::inf();
  for (int w = nw - 1; w >= 0; w--) {
    for (int k = 0; k < c; k++) R = dbl(R);
    // window sum = level-0 sum + chunk * level-1 sum
    XYZZ<F> l1 = hw[nw + w];
    for (uint32_t k = 1; k < P.chunk; k <<= 1) l1 = dbl(l1);
    add_full(R, hw[w]);
    add_full(R, l1);
  }
  for (int k = 0; k < c * P.win_lo; k++) R = dbl(R);  // (fixed-base mode: one bucket set, win_lo = 0 -> R = hw[0])
  Affine<F> a = to_affine(R);
  memcpy(P.out, &a, sizeof(a));
  return GPW_OK;
}

// fixed_windows == 0: windowed Pippenger over `points` (n bases).
// fixed_windows == W > 0 (fixed-base mode): `points` is a table of W n bases, entry w n + i = 2^(c w) P_i
// (gpw_msm_g1_fixed_table); W must equal the number of c-bit windows of a scalar. Every digit of every scalar then
// lands in ONE set of 2^(c-1) buckets, so the bucket reduction is paid once instead of once per window and wide
// windows become affordable: c = 22 needs 12 additions per full-width scalar instead of the 16 of c = 16.
// sort_tag / reuse_sort: MSMs over the SAME scalars (G1 and G2 sides of B, the commitment and its proof of knowledge,
// the A and K bases of the log-derivative quotients) share one digit decomposition + bucket sort: the first call sorts
// into the scratch named by sort_tag, the following ones pass reuse_sort = true (same n, c, window range, mode).
template <class F>
static int msm_dev_impl(gpw_ctx* ctx, const Fr* scalars, const Affine<F>* points, size_t n, int mont, int c,
                        int win_lo, int win_hi, uint64_t* out_affine, const char* tag, int fixed_windows = 0,
                        const char* sort_tag = nullptr, bool reuse_sort = false) {
  constexpr int OUT_WORDS = (int)(sizeof(Affine<F>) / 8);
  if (n >= (1ull << 31)) {
    set_error("msm: n=%zu too large (max 2^31-1)", n);
    return GPW_EINVAL;
  }
  if (c == 0) c = choose_window(n ? n : 1);
  if (c < 2 || c > (fixed_windows ? 24 : 16)) {
    set_error("msm: window_bits=%d out of range [2,%d]", c, fixed_windows ? 24 : 16);
    return GPW_EINVAL;
  }
  const int nwin = (254 + c) / c;  // ceil(255 / c): room for the final signed-digit carry
  if (fixed_windows && (fixed_windows != nwin || win_lo != 0 || (win_hi != 0 && win_hi != nwin))) {
    set_error("msm: fixed-base table must hold all %d windows of %d bits (got %d)", nwin, c, fixed_windows);
    return GPW_EINVAL;
  }
  if (win_lo == 0 && win_hi == 0) win_hi = nwin;
  if (win_lo < 0 || win_hi > nwin || win_lo >= win_hi) {
    set_error("msm: bad window range [%d,%d) of %d", win_lo, win_hi, nwin);
    return GPW_EINVAL;
  }
  if (n == 0) {
    for (int i = 0; i < OUT_WORDS; i++) out_affine[i] = 0;
    return GPW_OK;
  }
  GPW_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  const int nw = fixed_windows ? 1 : win_hi - win_lo;  // bucket sets
  const uint32_t half = 1u << (c - 1);
  const uint32_t B = (uint32_t)nw * half;
  const uint64_t max_entries = (uint64_t)n * (win_hi - win_lo);
  if (fixed_windows && max_entries >= (1ull << 31)) {
    set_error("msm: fixed-base table index n * windows = %llu exceeds 2^31", (unsigned long long)max_entries);
    return GPW_EINVAL;
  }
  if (max_entries >= (1ull << 32)) {
    set_error("msm: n * windows = %llu exceeds 2^32 entries; split the call", (unsigned long long)max_entries);
    return GPW_EINVAL;
  }
  const uint32_t ntasks = (uint32_t)((max_entries + MSM_TASK - 1) / MSM_TASK);
  // buckets per thread of the window reduction's first level (power of two; GPW_MSM_CHUNK overrides for experiments)
  static const uint32_t chunk_pref = [] {
    const char* e = getenv("GPW_MSM_CHUNK");
    const uint32_t v = e ? (uint32_t)atoi(e) : 16u;
    return (v >= 2 && v <= 64 && (v & (v - 1)) == 0) ? v : 16u;
  }();
  const uint32_t chunk = half < chunk_pref ? half : chunk_pref;
  const uint32_t nchunks = half / chunk;

  uint32_t *counts, *offsets, *cursor, *sorted, *tail_key, *tail_list, *ntail, *nbig, *big_list, *block_sums;
  XYZZ<F>*buckets, *head, *tail, *partials, *wsums;
  // deferred mode (common.cuh): the tail of the previous MSM may still read its scratch while this one starts - two sets
  const bool defer = ctx->msm_defer;
  const int parity = defer ? (ctx->msm_parity & 1) : 0;
  if (defer && ctx->n_pend >= gpw_ctx::MAX_PENDING) {
    set_error("msm: too many deferred MSMs");
    return GPW_EINVAL;
  }
  gpw_ctx::MsmPending& PD = ctx->pend[defer ? ctx->n_pend : gpw_ctx::MAX_PENDING - 1];
  std::string T(tag);
  if (defer) T += parity ? ".b" : ".a";
  const std::string ST(sort_tag ? std::string(sort_tag) : T);
  GPW_TRY(ctx->get_scratch((ST + ".counts").c_str(), (size_t)(B + 1) * 4 * 3 + 64, (void**)&counts));
  offsets = counts + (B + 1);
  cursor = offsets + (B + 1);
  const uint32_t nscan_blocks = (B + SCAN_BLOCK - 1) / SCAN_BLOCK;
  GPW_TRY(ctx->get_scratch((ST + ".scanblk").c_str(), (size_t)nscan_blocks * 4 + 16, (void**)&block_sums));
  GPW_TRY(ctx->get_scratch((ST + ".sorted").c_str(), (size_t)max_entries * 4 + 16, (void**)&sorted));
  GPW_TRY(ctx->get_scratch((T + ".keys").c_str(), (size_t)ntasks * 4 * 3 + 16, (void**)&tail_key));
  tail_list = tail_key + ntasks;
  big_list = tail_list + ntasks;
  ntail = big_list + ntasks;  // per-run counters live with the run's own scratch (a shared sort is read-only once built)
  nbig = ntail + 1;
  GPW_TRY(ctx->get_scratch((T + ".buckets").c_str(), (size_t)B * sizeof(XYZZ<F>), (void**)&buckets));
  GPW_TRY(ctx->get_scratch((T + ".head").c_str(), (size_t)ntasks * sizeof(XYZZ<F>), (void**)&head));
  GPW_TRY(ctx->get_scratch((T + ".tail").c_str(), (size_t)ntasks * sizeof(XYZZ<F>), (void**)&tail));
  const uint32_t ngroups = (nchunks + WINDOW_FINAL_PER_CTA - 1) / WINDOW_FINAL_PER_CTA;
  if (ngroups > WINDOW_FINAL_PER_CTA) {
    set_error("msm: too many buckets per window");
    return GPW_EINVAL;
  }
  // level 1 of the window reduction works on the nchunks chunk sums of each window
  const uint32_t chunk2 = nchunks < 16u ? nchunks : 16u;
  const uint32_t nchunks2 = nchunks / chunk2;
  const uint32_t ngroups2 = (nchunks2 + WINDOW_FINAL_PER_CTA - 1) / WINDOW_FINAL_PER_CTA;  // <= ngroups
  XYZZ<F>*chunk_sums, *partials2;
  GPW_TRY(ctx->get_scratch((T + ".partials").c_str(), (size_t)nw * (2 * (size_t)nchunks + nchunks2) * sizeof(XYZZ<F>), (void**)&partials));
  chunk_sums = partials + (size_t)nw * nchunks;
  partials2 = chunk_sums + (size_t)nw * nchunks;
  // wsums: [0, nw) level-0 sums T_w | [nw, 2 nw) level-1 sums | intermediates of the two-round sums
  GPW_TRY(ctx->get_scratch((T + ".wsums").c_str(), (size_t)nw * (2 + ngroups + ngroups2) * sizeof(XYZZ<F>), (void**)&wsums));

  // Batch-affine rounds (msm_affine.cuh) before the XYZZ accumulation. OFF by default: measured on B200 (profiles/
  // r02_batch_affine.md) the pairwise-tree rounds are bit-exact but slower than the XYZZ kernel - 6.4 instead of 9.5
  // multiplications per addition, yet 2 140 instead of ~1 700 instructions, one dependent multiplication chain per thread
  // (XYZZ has 2-3 independent ones) and three passes over the operands. gpw_ctx_set_option(ctx, "msm_affine_rounds", r) or
  // GPW_MSM_AFFINE_ROUNDS=r turn them on (r = 1..8) for experiments and for the parity tests of that path.
  int rounds = ctx->msm_affine_rounds;
  {
    static const char* env = getenv("GPW_MSM_AFFINE_ROUNDS");
    if (env && rounds < 0) rounds = atoi(env);
    rounds = std::max(0, std::min(8, rounds));
    if (max_entries < 2 * (uint64_t)MSM_TASK) rounds = 0;
  }
  uint32_t *offr = nullptr, *rcounts = nullptr;
  Affine<F>*affA = nullptr, *affB = nullptr;
  F* affPre = nullptr;  // prefix products of one round's denominators (one per output slot, rounded up to whole CTAs)
  if (rounds > 0) {
    GPW_TRY(ctx->get_scratch((ST + ".offr").c_str(), (size_t)rounds * (B + 1) * 4 + 16, (void**)&offr));
    GPW_TRY(ctx->get_scratch("msm.rcounts", (size_t)(B + 1) * 4 + 16, (void**)&rcounts));
    const uint64_t cap1 = (max_entries + 1) / 2 + B, cap2 = (cap1 + 1) / 2 + B;
    const char* grp = sizeof(Affine<F>) == 64 ? "msm.aff1" : "msm.aff2";
    GPW_TRY(ctx->get_scratch((std::string(grp) + "A").c_str(), (size_t)cap1 * sizeof(Affine<F>) + 16, (void**)&affA));
    if (rounds > 1) GPW_TRY(ctx->get_scratch((std::string(grp) + "B").c_str(), (size_t)cap2 * sizeof(Affine<F>) + 16, (void**)&affB));
    const size_t pre_slots = (size_t)div_up(cap1, (size_t)PAIR_THREADS * PairCfg<F>::P) * PAIR_THREADS * PairCfg<F>::P;
    GPW_TRY(ctx->get_scratch((std::string(grp) + "P").c_str(), pre_slots * sizeof(F) + 16, (void**)&affPre));
  }
  // a shared sort may only be reused by an MSM of exactly the shape it was built for
  if (sort_tag) {
    gpw_ctx::SortDesc want{n, c, win_lo, win_hi, fixed_windows, rounds, (const void*)scalars};
    gpw_ctx::SortDesc& have = ctx->sort_desc[ST];
    if (reuse_sort) {
      if (!(have == want)) {
        set_error("msm: shared sort '%s' was built for a different MSM (n, window, range, mode, rounds or scalars differ)", sort_tag);
        return GPW_EINVAL;
      }
    } else {
      have = want;
    }
  }

  if (defer && ctx->slot_used[parity]) GPW_CUDA(cudaStreamWaitEvent(st, ctx->slot_done[parity], 0));  // scratch set is free again
  if (defer && rounds > 0 && ctx->slot_used[parity ^ 1])  // the batch-affine scratch is single-buffered: no overlap with it on
    GPW_CUDA(cudaStreamWaitEvent(st, ctx->slot_done[parity ^ 1], 0));
  GPW_CUDA(cudaEventRecord(PD.ev[0], st));
  GPW_CUDA(cudaMemsetAsync(buckets, 0, (size_t)B * sizeof(XYZZ<F>), st));
  GPW_CUDA(cudaMemsetAsync(ntail, 0, 8, st));  // ntail, nbig
  if (!reuse_sort) {
    GPW_CUDA(cudaMemsetAsync(counts, 0, (size_t)(B + 1) * 4 * 3 + 64, st));
    const int TPB = 256;
    k_msm_digits<false><<<div_up(n, TPB), TPB, 0, st>>>(scalars, n, mont, c, nwin, win_lo, win_hi, fixed_windows, counts, nullptr);
    GPW_CHECK_LAUNCH();
    k_msm_scan_sums<<<nscan_blocks, SCAN_THREADS, 0, st>>>(counts, B, block_sums);
    GPW_CHECK_LAUNCH();
    k_msm_scan_blocks<<<1, SCAN_THREADS, 0, st>>>(block_sums, nscan_blocks, offsets + B);
    GPW_CHECK_LAUNCH();
    k_msm_scan_final<<<nscan_blocks, SCAN_THREADS, 0, st>>>(counts, B, block_sums, offsets, cursor);
    GPW_CHECK_LAUNCH();
    k_msm_digits<true><<<div_up(n, TPB), TPB, 0, st>>>(scalars, n, mont, c, nwin, win_lo, win_hi, fixed_windows, cursor, sorted);
    GPW_CHECK_LAUNCH();
    ctx->launches += 5;
  }
  GPW_CUDA(cudaEventRecord(PD.ev[1], st));
  // ---- batch-affine rounds (msm_affine.cuh): pairwise tree over the sorted list, then XYZZ for what is left --------------
  const Affine<F>* acc_points = points;
  const uint32_t* acc_sorted = sorted;
  const uint32_t* acc_offsets = offsets;
  uint32_t acc_ntasks = ntasks;
  if (rounds > 0) {
    constexpr int PP = PairCfg<F>::P;
    uint64_t cap = max_entries;
    const Affine<F>* in = points;
    const uint32_t* off_prev = offsets;
    for (int r = 1; r <= rounds; r++) {
      uint32_t* off_r = offr + (size_t)(r - 1) * (B + 1);
      if (!reuse_sort) {
        k_msm_round_counts<<<div_up(B, 256), 256, 0, st>>>(offsets, B, r, rcounts);
        GPW_CHECK_LAUNCH();
        k_msm_scan_sums<<<nscan_blocks, SCAN_THREADS, 0, st>>>(rcounts, B, block_sums);
        GPW_CHECK_LAUNCH();
        k_msm_scan_blocks<<<1, SCAN_THREADS, 0, st>>>(block_sums, nscan_blocks, off_r + B);
        GPW_CHECK_LAUNCH();
        k_msm_scan_final<<<nscan_blocks, SCAN_THREADS, 0, st>>>(rcounts, B, block_sums, off_r, nullptr);
        GPW_CHECK_LAUNCH();
        ctx->launches += 4;
      }
      cap = (cap + 1) / 2 + B;  // bucket b keeps ceil(k_b / 2) entries
      Affine<F>* out = (r & 1) ? affA : affB;
      const int grid = div_up(cap, (size_t)PAIR_THREADS * PP);
      static bool smem_set = false;  // (per template instance)
      if (!smem_set) {
        GPW_CUDA(cudaFuncSetAttribute(k_msm_pair_round<F, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PairRing<F>::BYTES));
        GPW_CUDA(cudaFuncSetAttribute(k_msm_pair_round<F, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PairRing<F>::BYTES));
        smem_set = true;
      }
      if (r == 1) k_msm_pair_round<F, true><<<grid, PAIR_THREADS, PairRing<F>::BYTES, st>>>(in, sorted, off_prev, off_r, B, out, affPre);
      else k_msm_pair_round<F, false><<<grid, PAIR_THREADS, PairRing<F>::BYTES, st>>>(in, nullptr, off_prev, off_r, B, out, affPre);
      GPW_CHECK_LAUNCH();
      ctx->launches += 1;
      in = out;
      off_prev = off_r;
    }
    acc_points = in;
    acc_sorted = nullptr;
    acc_offsets = off_prev;
    acc_ntasks = (uint32_t)((cap + MSM_TASK - 1) / MSM_TASK);
  }
  k_msm_accumulate<F><<<div_up(acc_ntasks, MSM_ACC_THREADS), MSM_ACC_THREADS, MSM_ACC_THREADS * sizeof(XYZZ<F>), st>>>(
      acc_points, acc_sorted, acc_offsets, B, buckets, head, tail, tail_key, tail_list, ntail);
  GPW_CHECK_LAUNCH();
  GPW_CUDA(cudaEventRecord(PD.ev[2], st));
  // development aid (GPW_DEBUG_MSM=1, synchronous MSMs only): split of the tail into fix-up / window reduction / sums
  static const bool dbg_env = getenv("GPW_DEBUG_MSM") != nullptr;
  const bool dbg_phases = dbg_env && !defer;
  cudaEvent_t dbg_ev[3] = {nullptr, nullptr, nullptr};
  if (dbg_phases)
    for (auto& e : dbg_ev) cudaEventCreate(&e);
  // the tail: short latency-bound grids. Deferred: on the context's high-priority stream, beside the next MSM.
  static const bool tail_stream = !getenv("GPW_MSM_TAIL_STREAM") || atoi(getenv("GPW_MSM_TAIL_STREAM")) != 0;
  const cudaStream_t gs = (defer && ctx->stream_hi && tail_stream) ? ctx->stream_hi : st;
  if (gs != st) {
    GPW_CUDA(cudaEventRecord(ctx->ev_hop, st));
    GPW_CUDA(cudaStreamWaitEvent(gs, ctx->ev_hop, 0));
  }
  k_msm_fixup<F><<<ctx->sm_count * 8, 128, 0, gs>>>(head, tail, acc_offsets, tail_key, tail_list, ntail, big_list, nbig, buckets);
  GPW_CHECK_LAUNCH();
  {
    const int fb_threads = sizeof(XYZZ<F>) > 128 ? 128 : 256;
    k_msm_fixup_big<F><<<ctx->sm_count * 2, fb_threads, fb_threads * sizeof(XYZZ<F>), gs>>>(head, tail, acc_offsets, tail_key, big_list,
                                                                                          nbig, buckets);
  }
  GPW_CHECK_LAUNCH();
  if (dbg_phases) cudaEventRecord(dbg_ev[0], gs);
  k_msm_window_partial<F><<<div_up((size_t)nw * nchunks, 128), 128, 0, gs>>>(buckets, offsets, half, chunk, (uint32_t)nw, 1u, partials,
                                                                             chunk_sums);
  GPW_CHECK_LAUNCH();
  k_msm_window_partial<F><<<div_up((size_t)nw * nchunks2, 128), 128, 0, gs>>>(chunk_sums, nullptr, nchunks, chunk2, (uint32_t)nw, 0u,
                                                                              partials2, nullptr);
  GPW_CHECK_LAUNCH();
  if (dbg_phases) cudaEventRecord(dbg_ev[1], gs);
  // per-window sums of an array of `cnt` partials per window -> dst[0 .. nw)  (two rounds above 2048 partials)
  auto sum_partials = [&](const XYZZ<F>* src, uint32_t cnt, uint32_t groups, XYZZ<F>* tmp, XYZZ<F>* dst) -> int {
    if (groups == 1) {
      k_msm_window_final<F><<<dim3(1, nw), 128, 128 * sizeof(XYZZ<F>), gs>>>(src, cnt, WINDOW_FINAL_PER_CTA, dst);
      GPW_CHECK_LAUNCH();
      ctx->launches += 1;
    } else {
      k_msm_window_final<F><<<dim3(groups, nw), 128, 128 * sizeof(XYZZ<F>), gs>>>(src, cnt, WINDOW_FINAL_PER_CTA, tmp);
      GPW_CHECK_LAUNCH();
      k_msm_window_final<F><<<dim3(1, nw), 128, 128 * sizeof(XYZZ<F>), gs>>>(tmp, groups, WINDOW_FINAL_PER_CTA, dst);
      GPW_CHECK_LAUNCH();
      ctx->launches += 2;
    }
    return GPW_OK;
  };
  GPW_TRY(sum_partials(partials, nchunks, ngroups, wsums + 2 * nw, wsums));
  GPW_TRY(sum_partials(partials2, nchunks2, ngroups2, wsums + (size_t)nw * (2 + ngroups), wsums + nw));
  ctx->launches += 5;
  const XYZZ<F>* hw = (const XYZZ<F>*)ctx->pin_take((size_t)2 * nw * sizeof(XYZZ<F>));
  const uint32_t* Mp = (const uint32_t*)ctx->pin_take(4);
  GPW_CUDA(cudaMemcpyAsync((void*)hw, wsums, (size_t)2 * nw * sizeof(XYZZ<F>), cudaMemcpyDeviceToHost, gs));
  GPW_CUDA(cudaMemcpyAsync((void*)Mp, offsets + B, 4, cudaMemcpyDeviceToHost, gs));
  GPW_CUDA(cudaEventRecord(PD.ev[3], gs));
  PD.group = sizeof(Affine<F>) == 64 ? 1 : 2;
  PD.hw = hw;
  PD.Mp = Mp;
  PD.nw = nw;
  PD.c = c;
  PD.win_lo = win_lo;
  PD.chunk = chunk;
  PD.n = n;
  PD.out = out_affine;
  PD.open = true;
  if (defer) {
    GPW_CUDA(cudaEventRecord(ctx->slot_done[parity], gs));
    ctx->slot_used[parity] = true;
    ctx->msm_parity ^= 1;
    ctx->n_pend++;
    return GPW_OK;
  }
  GPW_TRY(msm_finish_impl<F>(ctx, PD));
  if (dbg_phases) {
    float t_sort, t_fix, t_part, t_sum;
    cudaEventElapsedTime(&t_sort, PD.ev[0], PD.ev[1]);
    cudaEventElapsedTime(&t_fix, PD.ev[2], dbg_ev[0]);
    cudaEventElapsedTime(&t_part, dbg_ev[0], dbg_ev[1]);
    cudaEventElapsedTime(&t_sum, dbg_ev[1], PD.ev[3]);
    fprintf(stderr, "[gpw msm] %-6s n=%8zu c=%2d: sort %.2f | accumulate %.2f | fix-up %.2f | window partial %.2f | sums + copy %.2f ms\n", tag, n, c,
            t_sort, ctx->msm_acc_ms, t_fix, t_part, t_sum);
    for (auto& e : dbg_ev) cudaEventDestroy(e);
  }
  return GPW_OK;
}

// table[w n + i] = 2^(c w) P_i, w < W (one thread per base: c doublings and one inversion per entry; setup only)
template <class F>
__global__ void __launch_bounds__(128)
    k_msm_fixed_table(const Affine<F>* __restrict__ points, size_t n, int c, int W, Affine<F>* __restrict__ table) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const Affine<F> p = ld_struct(points + i);
  st_struct(table + i, p);
  XYZZ<F> cur = XYZZ<F>::from_affine(p);
#pragma unroll 1
  for (int w = 1; w < W; w++) {
#pragma unroll 1
    for (int k = 0; k < c; k++) cur = dbl(cur);
    const Affine<F> a = to_affine(cur);
    st_struct(table + (size_t)w * n + i, a);
    cur = XYZZ<F>::from_affine(a);  // keeps the next doublings' operands short-lived and the result canonical
  }
}

template <class F>
static int msm_fixed_table_impl(gpw_ctx* ctx, const Affine<F>* points, size_t n, int c, int W, Affine<F>* table) {
  if (!ctx || (!points && n) || (!table && n) || c < 2 || c > 24 || W != (254 + c) / c) {
    set_error("msm_fixed_table: bad argument (W must be ceil(255 / window_bits))");
    return GPW_EINVAL;
  }
  if (!n) return GPW_OK;
  GPW_CUDA(cudaSetDevice(ctx->device));
  k_msm_fixed_table<F><<<div_up(n, 128), 128, 0, ctx->stream>>>(points, n, c, W, table);
  GPW_CHECK_LAUNCH();
  ctx->launches += 1;
  return GPW_OK;
}

template <class F>
static int msm_host_impl(gpw_ctx* ctx, const uint64_t* scalars, const uint64_t* points, size_t n, int mont, int c,
                         uint64_t* out, const char* tag) {
  if (!ctx || (!scalars && n) || (!points && n) || !out) {
    set_error("msm: null argument");
    return GPW_EINVAL;
  }
  GPW_CUDA(cudaSetDevice(ctx->device));
  Fr* ds = nullptr;
  Affine<F>* dp = nullptr;
  std::string T(tag);
  GPW_TRY(ctx->get_scratch((T + ".in_scalars").c_str(), n * sizeof(Fr) + 16, (void**)&ds));
  GPW_TRY(ctx->get_scratch((T + ".in_points").c_str(), n * sizeof(Affine<F>) + 16, (void**)&dp));
  GPW_CUDA(cudaMemcpyAsync(ds, scalars, n * sizeof(Fr), cudaMemcpyHostToDevice, ctx->stream));
  GPW_CUDA(cudaMemcpyAsync(dp, points, n * sizeof(Affine<F>), cudaMemcpyHostToDevice, ctx->stream));
  return msm_dev_impl<F>(ctx, ds, dp, n, mont, c, 0, 0, out, tag);
}

}  // namespace gpw
