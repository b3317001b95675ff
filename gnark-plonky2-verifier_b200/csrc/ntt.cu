// K8 - radix-2 NTT over BN254 Fr for sm_100a.
//
// Replaces gnark-crypto fr/fft Domain.FFT / FFTInverse (+ OnCoset), un-vendored dependency of the
// reference, as used by groth16's computeH (7 transforms per proof). Conventions follow
// gnark-crypto: 2-adicity 28, w_N = w_28^(2^(28-logN)), coset generator 5.
//
// One kernel serves every pass of both algorithms: a pass owns `k` consecutive index bits
// [lobits, lobits + k); a CTA stages 2^k x C elements (C = 4 neighbouring columns, so every global
// access is a full 128 B line made of 32 B-sector-sized elements) in shared memory, runs the k
// butterfly stages there and writes back - one HBM read + one write per pass, 3 passes for 2^23.
//   DIF (natural -> bit-reversed): passes walk the index bits from the top, stages high bit -> low bit,
//        butterfly (u + v, (u - v) w)
//   DIT (bit-reversed -> natural): passes walk from the bottom, stages low bit -> high bit,
//        butterfly (u + v w, u - v w)
// The twiddle for the pair at global bit g with j = index mod 2^g is w^(j << (L - 1 - g)), read from a
// table of w^e, e < N/2 (inverse transforms use w^-e = -w^(N/2 - e), the sign folded into the
// butterfly). Coset / 1/N scaling is fused into the first pass's load or the last pass's store.
// The arithmetic is integer-pipe bound (11.5 Montgomery multiplies per element), not HBM bound.
#include "common.cuh"
#include "ff.cuh"
#include "tma.cuh"

namespace gpw {

constexpr int NTT_MAX_K = 8;
constexpr int NTT_COLS = 4;

__device__ __forceinline__ Fr ld_fr(const Fr* p) {
  Fr r;
  const uint4* s = reinterpret_cast<const uint4*>(p);
  uint4* d = reinterpret_cast<uint4*>(&r);
  d[0] = s[0];
  d[1] = s[1];
  return r;
}
__device__ __forceinline__ void st_fr(Fr* p, const Fr& v) {
  const uint4* s = reinterpret_cast<const uint4*>(&v);
  uint4* d = reinterpret_cast<uint4*>(p);
  d[0] = s[0];
  d[1] = s[1];
}

__device__ __forceinline__ uint32_t bitrev(uint32_t x, int bits) { return bits ? (__brev(x) >> (32 - bits)) : 0u; }

struct NttPass {
  int L;        // log2 N
  int lobits;   // index bits below the tile bits
  int k;        // stages in this pass
  int cols;     // columns per CTA (power of two <= NTT_COLS)
  int col_in_lo;  // 1: columns are consecutive `lo` values; 0: consecutive tiles (lobits == 0)
  int dit;      // 0 = DIF stage order / butterfly, 1 = DIT
  int inverse;  // use w^-e
  const Fr* scale_in;   // optional elementwise table applied on load  (index: position or bitrev(position))
  const Fr* scale_out;  // optional elementwise table applied on store
  int scale_in_bitrev, scale_out_bitrev;
  int has_scale_const;  // multiply every output by scale_const on store (the 1/N of a plain inverse transform)
  int tma;              // tile rows staged by TMA bulk copies (column-in-lo tiles of NTT_COLS columns: 128-byte rows)
  Fr scale_const;
};

// Shared-memory tile of 32-byte elements, addressed in 16-byte chunks. SWIZZLED layout: half h of element e lives at chunk
// 2e + (h ^ ((e >> 2) & 1)) - the 8 lanes of a quarter warp reading 8 consecutive elements with LDS.128 then touch 8
// different 16-byte bank groups on each of their two loads (elements 0..3 read chunks 0,2,4,6, elements 4..7 chunks
// 9,11,13,15); with the plain layout (half h at 2e + h) every such access is a 2-way conflict. The swap is done in the
// ADDRESS (each lane fetches its logical low half from wherever it lives), so no register is moved. The plain layout is
// what a TMA bulk copy leaves behind: the first butterfly stage of a TMA-staged tile reads plain and writes swizzled.
struct NttTile {
  uint4* ch;
  template <bool SW>
  __device__ __forceinline__ Fr ld(uint32_t e) const {
    const uint32_t s = SW ? ((e >> 2) & 1u) : 0u;
    Fr r;
    uint4* d = reinterpret_cast<uint4*>(&r);
    d[0] = ch[2 * e + s];
    d[1] = ch[2 * e + (s ^ 1u)];
    return r;
  }
  template <bool SW>
  __device__ __forceinline__ void st(uint32_t e, const Fr& v) const {
    const uint32_t s = SW ? ((e >> 2) & 1u) : 0u;
    const uint4* p = reinterpret_cast<const uint4*>(&v);
    ch[2 * e + s] = p[0];
    ch[2 * e + (s ^ 1u)] = p[1];
  }
};

// One stage for NB butterflies of a thread: all operands (and twiddles) are loaded first, then the NB independent
// Montgomery multiplications run back to back (they interleave in the IMAD pipe), then everything is stored - loads and
// stores through the same shared-memory pointers would otherwise keep the compiler from overlapping the butterflies.
// LDSW = false: first stage of a TMA-staged tile (plain layout as the bulk copies left it; the elementwise input scale of
// a coset transform, which the register path applies while loading, is applied here).
template <int NB, bool LDSW>
__device__ __forceinline__ void ntt_butterflies(const NttTile& sm, const Fr* __restrict__ tw, const NttPass& P, uint32_t bf0,
                                                uint32_t bf_stride, int lb, uint32_t lo_base, uint32_t hi, uint32_t halfN) {
  const uint32_t cols = (uint32_t)P.cols;
  const uint32_t mask = (1u << lb) - 1u;
  const int shift = P.L - 1 - (lb + P.lobits);
  uint32_t i0[NB], i1[NB];
  bool triv[NB];
  Fr u[NB], v[NB], w[NB];
#pragma unroll
  for (int i = 0; i < NB; i++) {
    const uint32_t bf = bf0 + (uint32_t)i * bf_stride;
    const uint32_t c = bf % cols, pi = bf / cols;
    const uint32_t t0 = ((pi & ~mask) << 1) | (pi & mask);
    const uint32_t lo = P.col_in_lo ? (lo_base + c) : 0u;
    const uint32_t e = ((((t0 & mask) << P.lobits) | lo)) << shift;  // < N/2
    i0[i] = t0 * cols + c;
    i1[i] = (t0 | (1u << lb)) * cols + c;
    triv[i] = e == 0;
    // inverse transforms use w^-e = -w^(N/2-e): the sign is folded into the butterfly below
    w[i] = ld_fr(tw + (triv[i] ? 0u : (P.inverse ? halfN - e : e)));
    u[i] = sm.ld<LDSW>(i0[i]);
    v[i] = sm.ld<LDSW>(i1[i]);
    if (!LDSW && P.scale_in) {  // (TMA tiles are column-in-lo tiles)
      const uint32_t g0 = (hi << (P.lobits + P.k)) | (t0 << P.lobits) | (lo_base + c);
      const uint32_t g1 = g0 | (1u << (lb + P.lobits));
      u[i] = mul(u[i], ld_fr(P.scale_in + (P.scale_in_bitrev ? bitrev(g0, P.L) : g0)));
      v[i] = mul(v[i], ld_fr(P.scale_in + (P.scale_in_bitrev ? bitrev(g1, P.L) : g1)));
    }
  }
#pragma unroll
  for (int i = 0; i < NB; i++) {
    Fr r0, r1;
    if (!P.dit) {
      r0 = add(u[i], v[i]);
      const Fr d = P.inverse && !triv[i] ? sub(v[i], u[i]) : sub(u[i], v[i]);
      r1 = triv[i] ? d : mul(d, w[i]);
    } else {
      const Fr vw = triv[i] ? v[i] : mul(v[i], w[i]);
      const bool flip = P.inverse && !triv[i];
      r0 = flip ? sub(u[i], vw) : add(u[i], vw);
      r1 = flip ? add(u[i], vw) : sub(u[i], vw);
    }
    u[i] = r0;
    v[i] = r1;
  }
#pragma unroll
  for (int i = 0; i < NB; i++) {
    sm.st<true>(i0[i], u[i]);
    sm.st<true>(i1[i], v[i]);
  }
}

__global__ void __launch_bounds__(128, 7) k_ntt_pass(Fr* __restrict__ data, const Fr* __restrict__ tw, NttPass P) {
  extern __shared__ uint4 smem_raw[];
  __shared__ uint64_t tile_bar;
  const int k = P.k, cols = P.cols;
  const uint32_t tile = 1u << k;
  const uint32_t nelem = tile * cols;
  const NttTile sm{smem_raw};
  const uint32_t tid = threadIdx.x;
  // CTA -> (hi, lo_base) or (tile_base)
  uint64_t cta = blockIdx.x;
  uint32_t lo_base, hi;
  if (P.col_in_lo) {
    uint32_t lo_groups = (1u << P.lobits) / cols;
    lo_base = (uint32_t)(cta % lo_groups) * cols;
    hi = (uint32_t)(cta / lo_groups);
  } else {
    lo_base = 0;
    hi = (uint32_t)cta * cols;
  }
  auto gindex = [&](uint32_t t, uint32_t c) -> uint32_t {
    if (P.col_in_lo) return (hi << (P.lobits + k)) | (t << P.lobits) | (lo_base + c);
    return ((hi + c) << k) | t;
  };
  if (P.tma) {
    // TMA staging: row t of the tile (NTT_COLS neighbouring columns = 128 contiguous bytes of global memory) is one bulk
    // copy into its 128 bytes of shared memory; the 2^k copies of the CTA complete on one mbarrier by byte count. No
    // thread moves data through registers; the tile arrives in the plain layout and the first stage re-swizzles it.
    if (tid == 0) {
      mbar_init(&tile_bar, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) mbar_expect_tx(&tile_bar, nelem * (uint32_t)sizeof(Fr));
    __syncthreads();  // the byte count is registered before any copy can complete on the barrier
    for (uint32_t t = tid; t < tile; t += blockDim.x)
      tma_bulk_load(smem_raw + (size_t)t * (2 * NTT_COLS), data + gindex(t, 0), NTT_COLS * (uint32_t)sizeof(Fr), &tile_bar);
    mbar_wait(&tile_bar, 0);
  } else {
    for (uint32_t e = tid; e < nelem; e += blockDim.x) {
      uint32_t c = e % cols, t = e / cols;
      uint32_t gi = gindex(t, c);
      Fr v = ld_fr(data + gi);
      if (P.scale_in) {
        uint32_t si = P.scale_in_bitrev ? bitrev(gi, P.L) : gi;
        v = mul(v, ld_fr(P.scale_in + si));
      }
      sm.st<true>(e, v);
    }
    __syncthreads();
  }
  const uint32_t halfN = 1u << (P.L - 1);
  const uint32_t nbf = nelem / 2;
  for (int q = 0; q < k; q++) {
    const int lb = P.dit ? q : (k - 1 - q);  // local pair bit
    // one butterfly at a time per thread: thread-level parallelism (7 CTAs of 128 threads per SM at 64 registers) beats
    // unrolling 4 butterflies per thread (148 registers, 3 CTAs per SM: measured 2.3 ms instead of 1.9 ms for 2^23)
    if (q == 0 && P.tma) {
#pragma unroll 1
      for (uint32_t bf = tid; bf < nbf; bf += blockDim.x) ntt_butterflies<1, false>(sm, tw, P, bf, 0, lb, lo_base, hi, halfN);
    } else {
#pragma unroll 1
      for (uint32_t bf = tid; bf < nbf; bf += blockDim.x) ntt_butterflies<1, true>(sm, tw, P, bf, 0, lb, lo_base, hi, halfN);
    }
    __syncthreads();
  }
  for (uint32_t e = tid; e < nelem; e += blockDim.x) {
    uint32_t c = e % cols, t = e / cols;
    uint32_t gi = gindex(t, c);
    Fr v = sm.ld<true>(e);
    if (P.scale_out) {
      uint32_t si = P.scale_out_bitrev ? bitrev(gi, P.L) : gi;
      v = mul(v, ld_fr(P.scale_out + si));
    }
    if (P.has_scale_const) v = mul(v, P.scale_const);
    st_fr(data + gi, v);
  }
}

__global__ void k_bitrev_permute(Fr* __restrict__ data, int L) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (1u << L)) return;
  uint32_t j = bitrev(i, L);
  if (i < j) {
    Fr a = ld_fr(data + i), b = ld_fr(data + j);
    st_fr(data + i, b);
    st_fr(data + j, a);
  }
}

// table[i] = base^i * c0 for i < n: each thread seeds base^(i0) by square-and-multiply, then walks
__global__ void k_power_table(Fr* __restrict__ table, uint32_t n, Fr base, Fr c0, uint32_t per_thread) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  uint64_t i0 = (uint64_t)t * per_thread;
  if (i0 >= n) return;
  uint32_t ew[1] = {(uint32_t)i0};
  Fr cur = mul(pow_words(base, ew, 1), c0);
  uint32_t end = (uint32_t)min((uint64_t)n, i0 + per_thread);
  for (uint32_t i = (uint32_t)i0; i < end; i++) {
    st_fr(table + i, cur);
    cur = mul(cur, base);
  }
}

// ---- host side --------------------------------------------------------------------------------
static Fr fr_from_u64(uint64_t v) {
  Fr a = Fr::zero();
  a.l[0] = (uint32_t)v;
  a.l[1] = (uint32_t)(v >> 32);
  return to_mont(a);
}

static Fr fr_pow_u64(Fr a, uint64_t e) {
  uint32_t w[2] = {(uint32_t)e, (uint32_t)(e >> 32)};
  return pow_words(a, w, 2);
}

// w_28 = 5^((r-1)/2^28)   
static Fr root_of_unity(int logn) {
  // (r - 1) / 2^28 as 8 LE words
  Fr m = modulus<FrParams>();
  uint32_t e[8];
  for (int i = 0; i < 8; i++) e[i] = m.l[i];
  e[0] -= 1;
  // shift right by 28
  uint32_t s[8];
  for (int i = 0; i < 8; i++) {
    uint64_t v = e[i];
    if (i + 1 < 8) v |= (uint64_t)e[i + 1] << 32;
    s[i] = (uint32_t)(v >> 28);
  }
  Fr w = pow_words(fr_from_u64(5), s, 8);
  for (int i = 0; i < 28 - logn; i++) w = sqr(w);
  return w;
}

static int ensure_tables(gpw_ctx* ctx, int L, NttTables** out) {
  auto it = ctx->ntt.find(L);
  if (it != ctx->ntt.end()) {
    *out = &it->second;
    return GPW_OK;
  }
  NttTables t;
  const uint32_t N = 1u << L;
  const uint32_t halfN = N >> 1;
  cudaStream_t st = ctx->stream;
  GPW_CUDA(cudaMalloc(&t.tw, (size_t)(halfN ? halfN : 1) * sizeof(Fr)));
  GPW_CUDA(cudaMalloc(&t.coset, (size_t)N * sizeof(Fr)));
  GPW_CUDA(cudaMalloc(&t.coset_inv, (size_t)N * sizeof(Fr)));
  Fr w = root_of_unity(L);
  Fr g = fr_from_u64(5);
  Fr ginv = inv(g);
  Fr ninv = inv(fr_from_u64(N));
  const uint32_t per = 64;
  if (halfN) {
    k_power_table<<<div_up(div_up(halfN, per), 128), 128, 0, st>>>((Fr*)t.tw, halfN, w, Fr::one(), per);
    GPW_CHECK_LAUNCH();
  }
  k_power_table<<<div_up(div_up(N, per), 128), 128, 0, st>>>((Fr*)t.coset, N, g, Fr::one(), per);
  GPW_CHECK_LAUNCH();
  k_power_table<<<div_up(div_up(N, per), 128), 128, 0, st>>>((Fr*)t.coset_inv, N, ginv, ninv, per);
  GPW_CHECK_LAUNCH();
  ctx->launches += 3;
  ctx->ntt[L] = t;
  *out = &ctx->ntt[L];
  return GPW_OK;
}

static int launch_pass(gpw_ctx* ctx, Fr* data, const NttTables* tb, NttPass P) {
  const uint32_t N = 1u << P.L;
  int cols = NTT_COLS;
  if (P.lobits == 0) {
    P.col_in_lo = 0;
    while ((uint64_t)cols << P.k > N) cols >>= 1;
  } else {
    P.col_in_lo = 1;
    while (cols > (1 << P.lobits)) cols >>= 1;
  }
  P.cols = cols;
  static const bool tma_on = !getenv("GPW_NTT_TMA") || atoi(getenv("GPW_NTT_TMA")) != 0;
  P.tma = (tma_on && P.col_in_lo && cols == NTT_COLS && P.k >= 1) ? 1 : 0;
  const uint32_t nelem = (1u << P.k) * cols;
  const uint32_t ctas = N / nelem;
  int threads = (int)(nelem / 2);
  static const int max_threads = getenv("GPW_NTT_THREADS") ? atoi(getenv("GPW_NTT_THREADS")) : 128;
  if (threads > max_threads) threads = max_threads;
  if (threads > 128) threads = 128;  // launch bounds of k_ntt_pass
  if (threads < 32) threads = 32;
  k_ntt_pass<<<ctas, threads, nelem * sizeof(Fr), ctx->stream>>>(data, (const Fr*)tb->tw, P);
  GPW_CHECK_LAUNCH();
  ctx->launches += 1;
  return GPW_OK;
}

static int ntt_dev_impl(gpw_ctx* ctx, Fr* data, int L, int inverse, int coset, int in_bitrev, int out_bitrev) {
  if (L < 0 || L > 27) {
    set_error("ntt: logn=%d out of range [0,27]", L);
    return GPW_EINVAL;
  }
  GPW_CUDA(cudaSetDevice(ctx->device));
  NttTables* tb;
  GPW_TRY(ensure_tables(ctx, L, &tb));
  const uint32_t N = 1u << L;
  cudaStream_t st = ctx->stream;
  if (L == 0) {
    return GPW_OK;  // size-1 transform is the identity (coset scale g^0 = 1, 1/N = 1)
  }
  if (in_bitrev && out_bitrev) {  // rare: make the input natural first
    k_bitrev_permute<<<div_up(N, 256), 256, 0, st>>>(data, L);
    GPW_CHECK_LAUNCH();
    ctx->launches += 1;
    in_bitrev = 0;
  }
  const bool dit = in_bitrev != 0;
  // split L stages into passes of <= NTT_MAX_K, avoiding a pass with lobits == 1 (needs >= 2 columns in lo)
  std::vector<int> ks;
  if (L > 3 * NTT_MAX_K) {
    // more than three passes: balance them (2^25: 7, 6, 6, 6) - the greedy split would end in a 2-stage pass (8, 8, 7, 2)
    // whose 16-element tiles pay a whole trip through HBM for two butterfly stages in nearly empty CTAs
    const int np = (L + NTT_MAX_K - 1) / NTT_MAX_K;
    for (int i = 0; i < np; i++) ks.push_back(L / np + (i < L % np ? 1 : 0));
  } else {
    int rem = L;
    while (rem > 0) {
      int k = rem > NTT_MAX_K ? NTT_MAX_K : rem;
      if (rem - k == 1) k -= 1;  // never leave a single bit for the last pass
      ks.push_back(k);
      rem -= k;
    }
  }
  // scaling tables: forward coset pre-scales by g^j (natural index j); inverse post-scales by g^-j / N
  // (or by 1/N alone).
  const Fr* pre = nullptr;
  const Fr* post = nullptr;
  if (!inverse && coset) pre = (const Fr*)tb->coset;
  if (inverse && coset) post = (const Fr*)tb->coset_inv;
  int bits_done = 0;
  for (size_t pi = 0; pi < ks.size(); pi++) {
    NttPass P{};
    P.L = L;
    P.k = ks[pi];
    P.dit = dit ? 1 : 0;
    P.inverse = inverse;
    // DIF walks from the top bit down, DIT from bit 0 up
    P.lobits = dit ? bits_done : (L - bits_done - P.k);
    if (pi == 0 && pre) {
      P.scale_in = pre;
      P.scale_in_bitrev = in_bitrev;  // data index is bitrev(j) when the input is bit-reversed
    }
    if (pi + 1 == ks.size() && post) {
      P.scale_out = post;
      P.scale_out_bitrev = dit ? 0 : 1;  // DIF leaves output bit-reversed: position p holds coefficient bitrev(p)
    }
    if (pi + 1 == ks.size() && inverse && !coset) {  // 1/N fused into the last pass's store
      P.has_scale_const = 1;
      P.scale_const = inv(fr_from_u64(N));
    }
    GPW_TRY(launch_pass(ctx, data, tb, P));
    bits_done += P.k;
  }
  const bool is_bitrev_now = !dit;
  if (is_bitrev_now != (out_bitrev != 0)) {
    k_bitrev_permute<<<div_up(N, 256), 256, 0, st>>>(data, L);
    GPW_CHECK_LAUNCH();
    ctx->launches += 1;
  }
  return GPW_OK;
}

}  // namespace gpw

using namespace gpw;

// Lets `to` (another context of the same device, e.g. a proving lane) use `from`'s twiddle / coset tables for 2^logn
// instead of building its own 0.6 GB copy. `from` must outlive `to`.
extern "C" int gpw_ntt_share_tables(gpw_ctx* from, gpw_ctx* to, int logn) {
  if (!from || !to || from->device != to->device || logn < 0 || logn > 27) {
    set_error("ntt_share_tables: bad argument");
    return GPW_EINVAL;
  }
  if (from == to || to->ntt.count(logn)) return GPW_OK;
  GPW_CUDA(cudaSetDevice(from->device));
  NttTables* tb;
  GPW_TRY(ensure_tables(from, logn, &tb));
  GPW_CUDA(cudaStreamSynchronize(from->stream));
  NttTables copy = *tb;
  copy.shared = true;
  to->ntt[logn] = copy;
  return GPW_OK;
}

extern "C" int gpw_ntt_fr_dev(gpw_ctx* ctx, uint64_t data_dev, int logn, int inverse, int coset, int in_bitrev,
                              int out_bitrev) {
  if (!ctx || !data_dev) {
    set_error("ntt: null argument");
    return GPW_EINVAL;
  }
  return ntt_dev_impl(ctx, (Fr*)data_dev, logn, inverse, coset, in_bitrev, out_bitrev);
}

extern "C" int gpw_ntt_fr(gpw_ctx* ctx, uint64_t* data, int logn, int inverse, int coset, int in_bitrev,
                          int out_bitrev) {
  if (!ctx || !data) {
    set_error("ntt: null argument");
    return GPW_EINVAL;
  }
  if (logn < 0 || logn > 27) {
    set_error("ntt: logn=%d out of range [0,27]", logn);
    return GPW_EINVAL;
  }
  GPW_CUDA(cudaSetDevice(ctx->device));
  size_t bytes = ((size_t)1 << logn) * sizeof(Fr);
  Fr* d;
  GPW_TRY(ctx->get_scratch("ntt.io", bytes, (void**)&d));
  GPW_CUDA(cudaMemcpyAsync(d, data, bytes, cudaMemcpyHostToDevice, ctx->stream));
  GPW_TRY(ntt_dev_impl(ctx, d, logn, inverse, coset, in_bitrev, out_bitrev));
  GPW_CUDA(cudaMemcpyAsync(data, d, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  GPW_CUDA(cudaStreamSynchronize(ctx->stream));
  return GPW_OK;
}
