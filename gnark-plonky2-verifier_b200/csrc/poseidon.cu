// K2 / K3 - Poseidon over Goldilocks (width 12) and over BN254 Fr (t = 4), Merkle-path walking and
// leaf hashing, plus K1 - the batched Goldilocks hints. One thread per independent unit; round
// constants live in __constant__ memory (every lane of a warp reads the same constant -> broadcast).
//
// Reference: poseidon/bn254.go:39-208, poseidon/goldilocks.go:30-331, fri/fri.go:97-144,
// goldilocks/base.go:223-359. Constant tables: poseidon_constants.inc (generated).
#include "common.cuh"
#include "ff.cuh"
#include "gl.cuh"
#include "poseidon_constants.inc"

namespace gpw {

__constant__ uint64_t c_bn_C[88 * 4];
__constant__ uint64_t c_bn_S[392 * 4];
__constant__ uint64_t c_bn_M[16 * 4];
__constant__ uint64_t c_bn_P[16 * 4];
__constant__ uint64_t c_gl_rc[360];
__constant__ uint64_t c_gl_circ[12];
__constant__ uint64_t c_gl_diag[12];
__constant__ uint64_t c_gl_first[12];
__constant__ uint64_t c_gl_partial_rc[22];
__constant__ uint64_t c_gl_vs[242];
__constant__ uint64_t c_gl_w_hats[242];
__constant__ uint64_t c_gl_init[121];

int load_poseidon_constants(gpw_ctx* ctx) {
  if (ctx->poseidon_consts_loaded) return GPW_OK;
  GPW_CUDA(cudaMemcpyToSymbol(c_bn_C, GPW_BN_C_MONT, sizeof(GPW_BN_C_MONT)));
  GPW_CUDA(cudaMemcpyToSymbol(c_bn_S, GPW_BN_S_MONT, sizeof(GPW_BN_S_MONT)));
  GPW_CUDA(cudaMemcpyToSymbol(c_bn_M, GPW_BN_M_MONT, sizeof(GPW_BN_M_MONT)));
  GPW_CUDA(cudaMemcpyToSymbol(c_bn_P, GPW_BN_P_MONT, sizeof(GPW_BN_P_MONT)));
  GPW_CUDA(cudaMemcpyToSymbol(c_gl_rc, GPW_GL_ALL_ROUND_CONSTANTS, sizeof(GPW_GL_ALL_ROUND_CONSTANTS)));
  GPW_CUDA(cudaMemcpyToSymbol(c_gl_circ, GPW_GL_MDS_CIRC, sizeof(GPW_GL_MDS_CIRC)));
  GPW_CUDA(cudaMemcpyToSymbol(c_gl_diag, GPW_GL_MDS_DIAG, sizeof(GPW_GL_MDS_DIAG)));
  GPW_CUDA(cudaMemcpyToSymbol(c_gl_first, GPW_GL_FAST_PARTIAL_FIRST_ROUND_CONSTANT,
                              sizeof(GPW_GL_FAST_PARTIAL_FIRST_ROUND_CONSTANT)));
  GPW_CUDA(cudaMemcpyToSymbol(c_gl_partial_rc, GPW_GL_FAST_PARTIAL_ROUND_CONSTANTS,
                              sizeof(GPW_GL_FAST_PARTIAL_ROUND_CONSTANTS)));
  GPW_CUDA(cudaMemcpyToSymbol(c_gl_vs, GPW_GL_FAST_PARTIAL_ROUND_VS, sizeof(GPW_GL_FAST_PARTIAL_ROUND_VS)));
  GPW_CUDA(cudaMemcpyToSymbol(c_gl_w_hats, GPW_GL_FAST_PARTIAL_ROUND_W_HATS, sizeof(GPW_GL_FAST_PARTIAL_ROUND_W_HATS)));
  GPW_CUDA(cudaMemcpyToSymbol(c_gl_init, GPW_GL_FAST_PARTIAL_ROUND_INITIAL_MATRIX,
                              sizeof(GPW_GL_FAST_PARTIAL_ROUND_INITIAL_MATRIX)));
  ctx->poseidon_consts_loaded = true;
  return GPW_OK;
}

// ---- Poseidon-BN254 -----------------------------------------------------------------------------
__device__ __forceinline__ Fr cfr(const uint64_t* tab, int idx) {
  Fr r;
  const uint32_t* s = reinterpret_cast<const uint32_t*>(tab + 4 * idx);
#pragma unroll
  for (int i = 0; i < 8; i++) r.l[i] = s[i];
  return r;
}

__device__ __forceinline__ Fr exp5(const Fr& x) {
  Fr x2 = sqr(x);
  Fr x4 = sqr(x2);
  return mul(x4, x);
}

// result[i] = sum_j m[j][i] * state[j]   (bn254.go:194-208)
__device__ __forceinline__ void bn_mix(Fr st[4], const uint64_t* m) {
  Fr out[4];
#pragma unroll
  for (int i = 0; i < 4; i++) {
    Fr acc = mul(cfr(m, 0 * 4 + i), st[0]);
#pragma unroll
    for (int j = 1; j < 4; j++) acc = add(acc, mul(cfr(m, j * 4 + i), st[j]));
    out[i] = acc;
  }
#pragma unroll
  for (int i = 0; i < 4; i++) st[i] = out[i];
}

__device__ __forceinline__ void bn_ark(Fr st[4], int it) {
#pragma unroll
  for (int i = 0; i < 4; i++) st[i] = add(st[i], cfr(c_bn_C, it + i));
}

// bn254.go:39-45, 130-170. State in Montgomery form.
__device__ void poseidon_bn254_perm(Fr st[4]) {
  bn_ark(st, 0);
#pragma unroll 1
  for (int i = 0; i < 3; i++) {
#pragma unroll
    for (int k = 0; k < 4; k++) st[k] = exp5(st[k]);
    bn_ark(st, (i + 1) * 4);
    bn_mix(st, c_bn_M);
  }
#pragma unroll
  for (int k = 0; k < 4; k++) st[k] = exp5(st[k]);
  bn_ark(st, 16);
  bn_mix(st, c_bn_P);
#pragma unroll 1
  for (int i = 0; i < 56; i++) {
    st[0] = exp5(st[0]);
    st[0] = add(st[0], cfr(c_bn_C, 20 + i));
    Fr n0 = mul(cfr(c_bn_S, 7 * i), st[0]);
#pragma unroll
    for (int j = 1; j < 4; j++) n0 = add(n0, mul(cfr(c_bn_S, 7 * i + j), st[j]));
#pragma unroll
    for (int k = 1; k < 4; k++) st[k] = add(st[k], mul(st[0], cfr(c_bn_S, 7 * i + 4 + k - 1)));
    st[0] = n0;
  }
#pragma unroll 1
  for (int i = 0; i < 3; i++) {
#pragma unroll
    for (int k = 0; k < 4; k++) st[k] = exp5(st[k]);
    bn_ark(st, 20 + 56 + i * 4);
    bn_mix(st, c_bn_M);
  }
#pragma unroll
  for (int k = 0; k < 4; k++) st[k] = exp5(st[k]);
  bn_mix(st, c_bn_M);
}

__device__ __forceinline__ Fr ld_fr_g(const uint64_t* p) {
  Fr r;
  const uint4* s = reinterpret_cast<const uint4*>(p);
  uint4* d = reinterpret_cast<uint4*>(&r);
  d[0] = s[0];
  d[1] = s[1];
  return r;
}
__device__ __forceinline__ void st_fr_g(uint64_t* p, const Fr& v) {
  const uint4* s = reinterpret_cast<const uint4*>(&v);
  uint4* d = reinterpret_cast<uint4*>(p);
  d[0] = s[0];
  d[1] = s[1];
}

__global__ void __launch_bounds__(128) k_poseidon_bn254(const uint64_t* __restrict__ in, uint64_t* __restrict__ out,
                                                        size_t n, int mont) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Fr st[4];
#pragma unroll
  for (int k = 0; k < 4; k++) {
    st[k] = ld_fr_g(in + (i * 4 + k) * 4);
    if (!mont) st[k] = to_mont(st[k]);
  }
  poseidon_bn254_perm(st);
#pragma unroll
  for (int k = 0; k < 4; k++) {
    if (!mont) st[k] = from_mont(st[k]);
    st_fr_g(out + (i * 4 + k) * 4, st[k]);
  }
}

// fri.go:97-116: walk one Merkle path. Canonical in/out.
__global__ void __launch_bounds__(128)
    k_merkle_paths_bn254(const uint64_t* __restrict__ leaf, const uint64_t* __restrict__ siblings,
                         const uint64_t* __restrict__ index_bits, size_t n, int depth, uint64_t* __restrict__ roots) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Fr cur = to_mont(ld_fr_g(leaf + i * 4));
  uint64_t bits = index_bits[i];
#pragma unroll 1
  for (int d = 0; d < depth; d++) {
    Fr sib = to_mont(ld_fr_g(siblings + (i * depth + d) * 4));
    bool bit = (bits >> d) & 1ull;
    Fr st[4];
    st[0] = Fr::zero();
    st[1] = Fr::zero();
    st[2] = bit ? sib : cur;   // Select(bit, sibling, cur): sibling goes LEFT when bit = 1 (fri.go:111)
    st[3] = bit ? cur : sib;
    poseidon_bn254_perm(st);
    cur = st[0];
  }
  st_fr_g(roots + i * 4, from_mont(cur));
}

// bn254.go:47-94 HashOrNoop / HashNoPad. leaves: n x leaf_len canonical Goldilocks u64.
__global__ void __launch_bounds__(128) k_hash_or_noop_bn254(const uint64_t* __restrict__ leaves, size_t n,
                                                            int leaf_len, uint64_t* __restrict__ digests) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint64_t* lf = leaves + i * (size_t)leaf_len;
  if (leaf_len <= 3) {
    uint64_t o[4] = {0, 0, 0, 0};
    for (int k = 0; k < leaf_len; k++) o[k] = lf[k];
    for (int k = 0; k < 4; k++) digests[i * 4 + k] = o[k];
    return;
  }
  Fr st[4] = {Fr::zero(), Fr::zero(), Fr::zero(), Fr::zero()};
#pragma unroll 1
  for (int base = 0; base < leaf_len; base += 9) {
    int end = min(leaf_len, base + 9);
    int slot = 1;
    for (int j = base; j < end; j += 3, slot++) {
      Fr v = Fr::zero();
      for (int k = 0; k < 3 && j + k < end; k++) {
        uint64_t x = lf[j + k];
        v.l[2 * k] = (uint32_t)x;
        v.l[2 * k + 1] = (uint32_t)(x >> 32);
      }
      Fr m = to_mont(v);
      // slots not covered by a short final chunk keep their previous value (overwrite-mode sponge)
      if (slot == 1) st[1] = m; else if (slot == 2) st[2] = m; else st[3] = m;
    }
    poseidon_bn254_perm(st);
  }
  st_fr_g(digests + i * 4, from_mont(st[0]));
}

// ---- Poseidon-Goldilocks (goldilocks.go:30-331), plain field arithmetic ---------------------------
__device__ __forceinline__ uint64_t gl_sbox(uint64_t x) {
  uint64_t x2 = gl::mul(x, x);
  uint64_t x3 = gl::mul(x, x2);
  uint64_t x6 = gl::mul(x3, x3);
  return gl::mul(x, x6);
}

// sum of <= 13 products of canonical values by small/medium constants, reduced once (mirrors the lazy
// MulAddNoReduce chain + one Reduce of mdsRowShf; only the remainder is produced here)
__device__ __forceinline__ void acc_mul(uint64_t& lo, uint64_t& hi, uint64_t& top, uint64_t a, uint64_t b) {
  uint64_t l, h;
  gl::mul64(a, b, l, h);
  uint64_t s = lo + l;
  uint64_t c = s < l ? 1u : 0u;
  lo = s;
  uint64_t s2 = hi + h;
  uint64_t c2 = s2 < h ? 1u : 0u;
  uint64_t s3 = s2 + c;
  c2 += s3 < s2 ? 1u : 0u;
  hi = s3;
  top += c2;
}
__device__ __forceinline__ uint64_t acc_reduce(uint64_t lo, uint64_t hi, uint64_t top) {
  // (top * 2^128 + hi * 2^64 + lo) mod p, top small
  uint64_t t = gl::reduce128(hi, top);
  return gl::reduce128(lo, t);
}

__device__ void poseidon_gl_perm(uint64_t st[12]) {
  int rc = 0;
  // first full rounds
#pragma unroll 1
  for (int r = 0; r < 4; r++) {
    uint64_t v[12];
#pragma unroll
    for (int i = 0; i < 12; i++) v[i] = gl_sbox(gl::add(st[i], c_gl_rc[i + 12 * rc]));
#pragma unroll 1
    for (int row = 0; row < 12; row++) {
      uint64_t lo = 0, hi = 0, top = 0;
      for (int i = 0; i < 12; i++) acc_mul(lo, hi, top, v[(i + row) % 12], c_gl_circ[i]);
      acc_mul(lo, hi, top, v[row], c_gl_diag[row]);
      st[row] = acc_reduce(lo, hi, top);
    }
    rc++;
  }
  // partial rounds
  {
    uint64_t v[12];
#pragma unroll
    for (int i = 0; i < 12; i++) v[i] = gl::add(st[i], c_gl_first[i]);
    st[0] = v[0];
#pragma unroll 1
    for (int d = 1; d < 12; d++) {
      uint64_t lo = 0, hi = 0, top = 0;
      for (int r = 1; r < 12; r++) acc_mul(lo, hi, top, v[r], c_gl_init[(r - 1) * 11 + (d - 1)]);
      st[d] = acc_reduce(lo, hi, top);
    }
  }
#pragma unroll 1
  for (int r = 0; r < 22; r++) {
    uint64_t s0 = gl::add(gl_sbox(st[0]), c_gl_partial_rc[r]);
    uint64_t lo = 0, hi = 0, top = 0;
    for (int i = 1; i < 12; i++) acc_mul(lo, hi, top, st[i], c_gl_w_hats[r * 11 + i - 1]);
    acc_mul(lo, hi, top, s0, 25ull);  // MDS0TO0
    uint64_t d = acc_reduce(lo, hi, top);
#pragma unroll
    for (int i = 1; i < 12; i++) st[i] = gl::add(gl::mul(s0, c_gl_vs[r * 11 + i - 1]), st[i]);
    st[0] = d;
  }
  rc += 22;
#pragma unroll 1
  for (int r = 0; r < 4; r++) {
    uint64_t v[12];
#pragma unroll
    for (int i = 0; i < 12; i++) v[i] = gl_sbox(gl::add(st[i], c_gl_rc[i + 12 * rc]));
#pragma unroll 1
    for (int row = 0; row < 12; row++) {
      uint64_t lo = 0, hi = 0, top = 0;
      for (int i = 0; i < 12; i++) acc_mul(lo, hi, top, v[(i + row) % 12], c_gl_circ[i]);
      acc_mul(lo, hi, top, v[row], c_gl_diag[row]);
      st[row] = acc_reduce(lo, hi, top);
    }
    rc++;
  }
}

__global__ void __launch_bounds__(128) k_poseidon_gl(const uint64_t* __restrict__ in, uint64_t* __restrict__ out,
                                                     size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint64_t st[12];
#pragma unroll
  for (int k = 0; k < 12; k++) st[k] = in[i * 12 + k];
  poseidon_gl_perm(st);
#pragma unroll
  for (int k = 0; k < 12; k++) out[i * 12 + k] = st[k];
}

// ---- K1: hints -----------------------------------------------------------------------------------
__global__ void k_gl_mul_add(const uint64_t* a, const uint64_t* b, const uint64_t* c, size_t n, uint64_t* q,
                             uint64_t* r, unsigned long long* bad) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint64_t x = a[i], y = b[i], z = c[i];
  if (x >= gl::P || y >= gl::P || z >= gl::P) {  // base.go:228-232 panics
    atomicMin(bad, (unsigned long long)i);
    return;
  }
  uint64_t qq, rr;
  gl::mul_add_hint(x, y, z, qq, rr);
  q[i] = qq;
  r[i] = rr;
}

__global__ void k_gl_reduce(const uint64_t* x4, size_t n, uint64_t* q4, uint64_t* r) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint64_t x[4] = {x4[4 * i], x4[4 * i + 1], x4[4 * i + 2], x4[4 * i + 3]};
  uint64_t q[4], rr;
  gl::reduce_hint(x, q, rr);
  for (int k = 0; k < 4; k++) q4[4 * i + k] = q[k];
  r[i] = rr;
}

__global__ void k_gl_inverse(const uint64_t* x, size_t n, uint64_t* out, unsigned long long* bad) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint64_t v = x[i];
  if (v >= gl::P) {  // base.go:322-324
    atomicMin(bad, (unsigned long long)i);
    return;
  }
  out[i] = gl::inverse(v);
}

__global__ void k_gl_split(const uint64_t* x, size_t n, uint64_t* hi, uint64_t* lo, unsigned long long* bad) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint64_t v = x[i];
  if (v >= gl::P) {  // base.go:347-349
    atomicMin(bad, (unsigned long long)i);
    return;
  }
  uint64_t h, l;
  gl::split_limbs_hint(v, h, l);
  hi[i] = h;
  lo[i] = l;
}

// ---- host wrappers --------------------------------------------------------------------------------
struct DevBuf {
  gpw_ctx* ctx;
  DevBuf(gpw_ctx* c) : ctx(c) {}
  int up(const char* name, const void* host, size_t bytes, void** dev) {
    GPW_TRY(ctx->get_scratch(name, bytes + 16, dev));
    if (bytes) GPW_CUDA(cudaMemcpyAsync(*dev, host, bytes, cudaMemcpyHostToDevice, ctx->stream));
    return GPW_OK;
  }
  int alloc(const char* name, size_t bytes, void** dev) { return ctx->get_scratch(name, bytes + 16, dev); }
  int down(void* host, const void* dev, size_t bytes) {
    if (bytes) GPW_CUDA(cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    return GPW_OK;
  }
};

static int check_bad(gpw_ctx* ctx, unsigned long long* dbad, const char* what) {
  unsigned long long bad = 0;
  GPW_CUDA(cudaMemcpyAsync(&bad, dbad, 8, cudaMemcpyDeviceToHost, ctx->stream));
  GPW_CUDA(cudaStreamSynchronize(ctx->stream));
  if (bad != ~0ull) {
    set_error("%s: input %llu is not in the field", what, bad);
    return GPW_EHINT;
  }
  return GPW_OK;
}

}  // namespace gpw

using namespace gpw;

extern "C" int gpw_poseidon_bn254_dev(gpw_ctx* ctx, uint64_t in_dev, uint64_t out_dev, size_t n, int mont) {
  if (!ctx || ((!in_dev || !out_dev) && n)) {
    set_error("poseidon_bn254: null argument");
    return GPW_EINVAL;
  }
  GPW_CUDA(cudaSetDevice(ctx->device));
  GPW_TRY(load_poseidon_constants(ctx));
  if (!n) return GPW_OK;
  k_poseidon_bn254<<<div_up(n, 128), 128, 0, ctx->stream>>>((const uint64_t*)in_dev, (uint64_t*)out_dev, n, mont);
  GPW_CHECK_LAUNCH();
  ctx->launches++;
  return GPW_OK;
}

extern "C" int gpw_poseidon_bn254(gpw_ctx* ctx, const uint64_t* in, uint64_t* out, size_t n, int mont) {
  if (!ctx || ((!in || !out) && n)) {
    set_error("poseidon_bn254: null argument");
    return GPW_EINVAL;
  }
  GPW_CUDA(cudaSetDevice(ctx->device));
  DevBuf B(ctx);
  void *di, *dout;
  GPW_TRY(B.up("pbn.in", in, n * 128, &di));
  GPW_TRY(B.alloc("pbn.out", n * 128, &dout));
  GPW_TRY(gpw_poseidon_bn254_dev(ctx, (uint64_t)di, (uint64_t)dout, n, mont));
  GPW_TRY(B.down(out, dout, n * 128));
  GPW_CUDA(cudaStreamSynchronize(ctx->stream));
  return GPW_OK;
}

extern "C" int gpw_merkle_paths_bn254(gpw_ctx* ctx, const uint64_t* leaf_digests, const uint64_t* siblings,
                                      const uint64_t* index_bits, size_t n, int depth, uint64_t* roots_out) {
  if (!ctx || ((!leaf_digests || !index_bits || !roots_out || (!siblings && depth)) && n) || depth < 0 || depth > 64) {
    set_error("merkle_paths: bad argument");
    return GPW_EINVAL;
  }
  GPW_CUDA(cudaSetDevice(ctx->device));
  GPW_TRY(load_poseidon_constants(ctx));
  if (!n) return GPW_OK;
  DevBuf B(ctx);
  void *dl, *ds, *db, *dr;
  GPW_TRY(B.up("mk.leaf", leaf_digests, n * 32, &dl));
  GPW_TRY(B.up("mk.sib", siblings, n * (size_t)depth * 32, &ds));
  GPW_TRY(B.up("mk.bits", index_bits, n * 8, &db));
  GPW_TRY(B.alloc("mk.roots", n * 32, &dr));
  k_merkle_paths_bn254<<<div_up(n, 128), 128, 0, ctx->stream>>>((const uint64_t*)dl, (const uint64_t*)ds,
                                                               (const uint64_t*)db, n, depth, (uint64_t*)dr);
  GPW_CHECK_LAUNCH();
  ctx->launches++;
  GPW_TRY(B.down(roots_out, dr, n * 32));
  GPW_CUDA(cudaStreamSynchronize(ctx->stream));
  return GPW_OK;
}

extern "C" int gpw_hash_or_noop_bn254(gpw_ctx* ctx, const uint64_t* leaves, size_t n, int leaf_len,
                                      uint64_t* digests_out) {
  if (!ctx || leaf_len < 0 || (n && ((!leaves && leaf_len) || !digests_out))) {
    set_error("hash_or_noop: bad argument");
    return GPW_EINVAL;
  }
  GPW_CUDA(cudaSetDevice(ctx->device));
  GPW_TRY(load_poseidon_constants(ctx));
  if (!n) return GPW_OK;
  DevBuf B(ctx);
  void *dl, *dd;
  GPW_TRY(B.up("hn.leaves", leaves, n * (size_t)leaf_len * 8, &dl));
  GPW_TRY(B.alloc("hn.dig", n * 32, &dd));
  k_hash_or_noop_bn254<<<div_up(n, 128), 128, 0, ctx->stream>>>((const uint64_t*)dl, n, leaf_len, (uint64_t*)dd);
  GPW_CHECK_LAUNCH();
  ctx->launches++;
  GPW_TRY(B.down(digests_out, dd, n * 32));
  GPW_CUDA(cudaStreamSynchronize(ctx->stream));
  return GPW_OK;
}

extern "C" int gpw_poseidon_gl(gpw_ctx* ctx, const uint64_t* in, uint64_t* out, size_t n) {
  if (!ctx || ((!in || !out) && n)) {
    set_error("poseidon_gl: null argument");
    return GPW_EINVAL;
  }
  GPW_CUDA(cudaSetDevice(ctx->device));
  GPW_TRY(load_poseidon_constants(ctx));
  if (!n) return GPW_OK;
  DevBuf B(ctx);
  void *di, *dout;
  GPW_TRY(B.up("pgl.in", in, n * 96, &di));
  GPW_TRY(B.alloc("pgl.out", n * 96, &dout));
  k_poseidon_gl<<<div_up(n, 128), 128, 0, ctx->stream>>>((const uint64_t*)di, (uint64_t*)dout, n);
  GPW_CHECK_LAUNCH();
  ctx->launches++;
  GPW_TRY(B.down(out, dout, n * 96));
  GPW_CUDA(cudaStreamSynchronize(ctx->stream));
  return GPW_OK;
}

extern "C" int gpw_gl_mul_add_hint(gpw_ctx* ctx, const uint64_t* a, const uint64_t* b, const uint64_t* c, size_t n,
                                   uint64_t* q, uint64_t* r) {
  if (!ctx || ((!a || !b || !c || !q || !r) && n)) {
    set_error("mul_add_hint: null argument");
    return GPW_EINVAL;
  }
  GPW_CUDA(cudaSetDevice(ctx->device));
  if (!n) return GPW_OK;
  DevBuf B(ctx);
  void *da, *db, *dc, *dq, *dr, *dbad;
  GPW_TRY(B.up("h.a", a, n * 8, &da));
  GPW_TRY(B.up("h.b", b, n * 8, &db));
  GPW_TRY(B.up("h.c", c, n * 8, &dc));
  GPW_TRY(B.alloc("h.q", n * 8, &dq));
  GPW_TRY(B.alloc("h.r", n * 8, &dr));
  GPW_TRY(B.alloc("h.bad", 8, &dbad));
  GPW_CUDA(cudaMemsetAsync(dbad, 0xff, 8, ctx->stream));
  k_gl_mul_add<<<div_up(n, 256), 256, 0, ctx->stream>>>((uint64_t*)da, (uint64_t*)db, (uint64_t*)dc, n, (uint64_t*)dq,
                                                        (uint64_t*)dr, (unsigned long long*)dbad);
  GPW_CHECK_LAUNCH();
  ctx->launches++;
  GPW_TRY(B.down(q, dq, n * 8));
  GPW_TRY(B.down(r, dr, n * 8));
  return check_bad(ctx, (unsigned long long*)dbad, "MulAddHint");
}

extern "C" int gpw_gl_reduce_hint(gpw_ctx* ctx, const uint64_t* x4, size_t n, uint64_t* q4, uint64_t* r) {
  if (!ctx || ((!x4 || !q4 || !r) && n)) {
    set_error("reduce_hint: null argument");
    return GPW_EINVAL;
  }
  GPW_CUDA(cudaSetDevice(ctx->device));
  if (!n) return GPW_OK;
  DevBuf B(ctx);
  void *dx, *dq, *dr;
  GPW_TRY(B.up("h.a", x4, n * 32, &dx));
  GPW_TRY(B.alloc("h.q", n * 32, &dq));
  GPW_TRY(B.alloc("h.r", n * 8, &dr));
  k_gl_reduce<<<div_up(n, 256), 256, 0, ctx->stream>>>((uint64_t*)dx, n, (uint64_t*)dq, (uint64_t*)dr);
  GPW_CHECK_LAUNCH();
  ctx->launches++;
  GPW_TRY(B.down(q4, dq, n * 32));
  GPW_TRY(B.down(r, dr, n * 8));
  GPW_CUDA(cudaStreamSynchronize(ctx->stream));
  return GPW_OK;
}

extern "C" int gpw_gl_inverse_hint(gpw_ctx* ctx, const uint64_t* x, size_t n, uint64_t* inv_out) {
  if (!ctx || ((!x || !inv_out) && n)) {
    set_error("inverse_hint: null argument");
    return GPW_EINVAL;
  }
  GPW_CUDA(cudaSetDevice(ctx->device));
  if (!n) return GPW_OK;
  DevBuf B(ctx);
  void *dx, *dout, *dbad;
  GPW_TRY(B.up("h.a", x, n * 8, &dx));
  GPW_TRY(B.alloc("h.q", n * 8, &dout));
  GPW_TRY(B.alloc("h.bad", 8, &dbad));
  GPW_CUDA(cudaMemsetAsync(dbad, 0xff, 8, ctx->stream));
  k_gl_inverse<<<div_up(n, 256), 256, 0, ctx->stream>>>((uint64_t*)dx, n, (uint64_t*)dout, (unsigned long long*)dbad);
  GPW_CHECK_LAUNCH();
  ctx->launches++;
  GPW_TRY(B.down(inv_out, dout, n * 8));
  return check_bad(ctx, (unsigned long long*)dbad, "InverseHint");
}

extern "C" int gpw_gl_split_limbs_hint(gpw_ctx* ctx, const uint64_t* x, size_t n, uint64_t* hi, uint64_t* lo) {
  if (!ctx || ((!x || !hi || !lo) && n)) {
    set_error("split_limbs_hint: null argument");
    return GPW_EINVAL;
  }
  GPW_CUDA(cudaSetDevice(ctx->device));
  if (!n) return GPW_OK;
  DevBuf B(ctx);
  void *dx, *dh, *dl, *dbad;
  GPW_TRY(B.up("h.a", x, n * 8, &dx));
  GPW_TRY(B.alloc("h.q", n * 8, &dh));
  GPW_TRY(B.alloc("h.r", n * 8, &dl));
  GPW_TRY(B.alloc("h.bad", 8, &dbad));
  GPW_CUDA(cudaMemsetAsync(dbad, 0xff, 8, ctx->stream));
  k_gl_split<<<div_up(n, 256), 256, 0, ctx->stream>>>((uint64_t*)dx, n, (uint64_t*)dh, (uint64_t*)dl,
                                                      (unsigned long long*)dbad);
  GPW_CHECK_LAUNCH();
  ctx->launches++;
  GPW_TRY(B.down(hi, dh, n * 8));
  GPW_TRY(B.down(lo, dl, n * 8));
  return check_bad(ctx, (unsigned long long*)dbad, "SplitLimbsHint");
}
