// Groth16 prover glue kernels over BN254: computeH's pointwise step and synthetic-base generation.
// (The MSMs and NTTs themselves are msm_*.cu / ntt.cu.)  Replaces the corresponding steps of gnark's
// backend/groth16 (bn254) Prove, reached from benchmark.go:249.
#include "common.cuh"
#include "ec.cuh"
#include "host_ec.cuh"

namespace gpw {

__device__ __forceinline__ Fr ldfr(const Fr* p) {
  Fr r;
  const uint4* s = reinterpret_cast<const uint4*>(p);
  uint4* d = reinterpret_cast<uint4*>(&r);
  d[0] = s[0];
  d[1] = s[1];
  return r;
}
__device__ __forceinline__ void stfr(Fr* p, const Fr& v) {
  const uint4* s = reinterpret_cast<const uint4*>(&v);
  uint4* d = reinterpret_cast<uint4*>(p);
  d[0] = s[0];
  d[1] = s[1];
}

// a[i] = (a[i] * b[i] - c[i]) * k      (computeH: (A.B - C) / Z_H on the coset, Z_H constant there)
__global__ void k_h_pointwise(Fr* __restrict__ a, const Fr* __restrict__ b, const Fr* __restrict__ c, size_t n, Fr k) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  stfr(a + i, mul(sub(mul(ldfr(a + i), ldfr(b + i)), ldfr(c + i)), k));
}

// a[i] *= b[i]
__global__ void k_fr_mul_inplace(Fr* __restrict__ a, const Fr* __restrict__ b, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  stfr(a + i, mul(ldfr(a + i), ldfr(b + i)));
}

// computeH, last step: v and c hold polynomial coefficients in BIT-REVERSED order; h = (v - c) * k is left in v in natural
// order (the subtraction, the scaling and the bit-reversal permutation in one pass)
__global__ void k_h_finish_bitrev(Fr* __restrict__ v, const Fr* __restrict__ c, int L, Fr k) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (1u << L)) return;
  const uint32_t j = L ? (__brev(i) >> (32 - L)) : 0u;
  if (i > j) return;
  const Fr xi = mul(sub(ldfr(v + i), ldfr(c + i)), k);
  if (i == j) {
    stfr(v + i, xi);
    return;
  }
  const Fr xj = mul(sub(ldfr(v + j), ldfr(c + j)), k);
  stfr(v + i, xj);
  stfr(v + j, xi);
}

__global__ void k_fr_convert(Fr* __restrict__ a, size_t n, int to) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Fr v = ldfr(a + i);
  stfr(a + i, to ? to_mont(v) : from_mont(v));
}

// out[i] = [k0 + i] G as affine points. Setup-only (synthetic proving keys / test bases), so each point
// simply pays its own Fermat inversion.
template <class F, int PER>
__global__ void __launch_bounds__(128) k_gen_multiples(Affine<F> g, uint64_t k0, size_t n, Affine<F>* __restrict__ out) {
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t i0 = t * PER;
  if (i0 >= n) return;
  uint64_t k = k0 + i0;
  XYZZ<F> cur = XYZZ<F>::inf();
  for (int b = 63; b >= 0; b--) {
    cur = dbl(cur);
    if ((k >> b) & 1ull) add_mixed(cur, g, false);
  }
  int cnt = (int)min((size_t)PER, n - i0);
#pragma unroll 1
  for (int j = 0; j < cnt; j++) {
    out[i0 + j] = to_affine(cur);
    add_mixed(cur, g, false);
  }
}

}  // namespace gpw

using namespace gpw;

extern "C" int gpw_fr_h_pointwise_dev(gpw_ctx* ctx, uint64_t a_dev, uint64_t b_dev, uint64_t c_dev, size_t n,
                                      const uint64_t* k_mont) {
  if (!ctx || !k_mont || ((!a_dev || !b_dev || !c_dev) && n)) {
    set_error("h_pointwise: null argument");
    return GPW_EINVAL;
  }
  GPW_CUDA(cudaSetDevice(ctx->device));
  if (!n) return GPW_OK;
  Fr k;
  memcpy(&k, k_mont, 32);
  k_h_pointwise<<<div_up(n, 256), 256, 0, ctx->stream>>>((Fr*)a_dev, (const Fr*)b_dev, (const Fr*)c_dev, n, k);
  GPW_CHECK_LAUNCH();
  ctx->launches++;
  return GPW_OK;
}

extern "C" int gpw_fr_convert_dev(gpw_ctx* ctx, uint64_t a_dev, size_t n, int to_mont_flag) {
  if (!ctx || (!a_dev && n)) {
    set_error("fr_convert: null argument");
    return GPW_EINVAL;
  }
  GPW_CUDA(cudaSetDevice(ctx->device));
  if (!n) return GPW_OK;
  k_fr_convert<<<div_up(n, 256), 256, 0, ctx->stream>>>((Fr*)a_dev, n, to_mont_flag);
  GPW_CHECK_LAUNCH();
  ctx->launches++;
  return GPW_OK;
}

extern "C" int gpw_ec_generator_multiples_dev(gpw_ctx* ctx, int group, uint64_t k0, size_t n, uint64_t out_dev) {
  if (!ctx || (!out_dev && n)) {
    set_error("generator_multiples: null argument");
    return GPW_EINVAL;
  }
  GPW_CUDA(cudaSetDevice(ctx->device));
  if (!n) return GPW_OK;
  constexpr int PER = 32;
  int blocks = div_up(div_up(n, PER), 128);
  if (group == 1) {
    G1Affine g{fp_from_u64(1), fp_from_u64(2)};
    k_gen_multiples<Fp, PER><<<blocks, 128, 0, ctx->stream>>>(g, k0, n, (G1Affine*)out_dev);
  } else if (group == 2) {
    G2Affine g{{fp_from_dec("10857046999023057135944570762232829481370756359578518086990519993285655852781"),
                fp_from_dec("11559732032986387107991004021392285783925812861821192530917403151452391805634")},
               {fp_from_dec("8495653923123431417604973247489272438418190587263600148770280649306958101930"),
                fp_from_dec("4082367875863433681332203403145435568316851327593401208105741076214120093531")}};
    k_gen_multiples<Fp2, PER><<<blocks, 128, 0, ctx->stream>>>(g, k0, n, (G2Affine*)out_dev);
  } else {
    set_error("generator_multiples: group must be 1 or 2");
    return GPW_EINVAL;
  }
  GPW_CHECK_LAUNCH();
  ctx->launches++;
  return GPW_OK;
}

// ---------------------------------------------------------------------------------------------------
// Groth16 proving key + prover core (gnark backend/groth16 bn254: ProvingKey, Prove)
// ---------------------------------------------------------------------------------------------------
struct gpw_pk {
  gpw_ctx* ctx = nullptr;
  size_t m = 0;        // number of wires (incl. the constant-one wire)
  size_t n_pub = 0;    // wires [0, n_pub) are public (excluded from K)
  int logN = 0;        // FFT domain
  G1Affine *A = nullptr, *B1 = nullptr, *K = nullptr, *Z = nullptr;  // device
  G2Affine* B2 = nullptr;                                            // device
  G1Affine alpha1, beta1, delta1;
  G2Affine beta2, delta2;
  // scratch for computeH
  Fr *ha = nullptr, *hb = nullptr, *hc = nullptr;
  float t_h_ms = 0, t_msm_ms[5] = {0, 0, 0, 0, 0};
  cudaEvent_t h0 = nullptr, h1 = nullptr;  // computeH timing, created on first use
};

static int dev_alloc(void** p, size_t bytes) {
  cudaError_t e = cudaMalloc(p, bytes ? bytes : 16);
  if (e != cudaSuccess) {
    set_error("cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
    return GPW_ENOMEM;
  }
  return GPW_OK;
}

extern "C" void gpw_groth16_pk_free(gpw_pk* pk) {
  if (!pk) return;
  cudaSetDevice(pk->ctx->device);
  cudaFree(pk->A);
  cudaFree(pk->B1);
  cudaFree(pk->K);
  cudaFree(pk->Z);
  cudaFree(pk->B2);
  cudaFree(pk->ha);
  cudaFree(pk->hb);
  cudaFree(pk->hc);
  if (pk->h0) cudaEventDestroy(pk->h0);
  if (pk->h1) cudaEventDestroy(pk->h1);
  delete pk;
}

template <class F>
static Affine<F> host_gen_mul(uint64_t k) {
  uint32_t kw[8] = {(uint32_t)k, (uint32_t)(k >> 32), 0, 0, 0, 0, 0, 0};
  return to_affine(host_scalar_mul(generator<F>(), kw));
}

// Synthetic key with KNOWN discrete logs (every base is [k]G for a documented k), the analogue of gnark's
// groth16.DummySetup (benchmark.go:214): same shapes and cost as a real key, and - unlike DummySetup - it
// lets the tests recompute the expected proof exactly "in the exponent".
//   A_i = [1 + i]G1, B1_i = [1 + m + i]G1, K_i = [1 + 2m + i]G1 (i >= n_pub), Z_j = [1 + 3m + j]G1 (j < N-1),
//   B2_i = [1 + i]G2, alpha1 = [seed+1]G1, beta1 = [seed+2]G1, delta1 = [seed+3]G1, beta2 = [seed+2]G2,
//   delta2 = [seed+3]G2.
extern "C" int gpw_groth16_pk_synthetic(gpw_ctx* ctx, size_t m, size_t n_pub, int logN, uint64_t seed, gpw_pk** out) {
  if (!ctx || !out || m == 0 || n_pub > m || logN < 1 || logN > 27) {
    set_error("pk_synthetic: bad argument");
    return GPW_EINVAL;
  }
  GPW_CUDA(cudaSetDevice(ctx->device));
  gpw_pk* pk = new gpw_pk();
  pk->ctx = ctx;
  pk->m = m;
  pk->n_pub = n_pub;
  pk->logN = logN;
  const size_t N = (size_t)1 << logN;
  int rc;
  if ((rc = dev_alloc((void**)&pk->A, m * sizeof(G1Affine))) || (rc = dev_alloc((void**)&pk->B1, m * sizeof(G1Affine))) ||
      (rc = dev_alloc((void**)&pk->K, m * sizeof(G1Affine))) || (rc = dev_alloc((void**)&pk->Z, N * sizeof(G1Affine))) ||
      (rc = dev_alloc((void**)&pk->B2, m * sizeof(G2Affine))) || (rc = dev_alloc((void**)&pk->ha, N * sizeof(Fr))) ||
      (rc = dev_alloc((void**)&pk->hb, N * sizeof(Fr))) || (rc = dev_alloc((void**)&pk->hc, N * sizeof(Fr)))) {
    gpw_groth16_pk_free(pk);
    return rc;
  }
  if ((rc = gpw_ec_generator_multiples_dev(ctx, 1, 1, m, (uint64_t)pk->A)) ||
      (rc = gpw_ec_generator_multiples_dev(ctx, 1, 1 + m, m, (uint64_t)pk->B1)) ||
      (rc = gpw_ec_generator_multiples_dev(ctx, 1, 1 + 2 * m, m, (uint64_t)pk->K)) ||
      (rc = gpw_ec_generator_multiples_dev(ctx, 1, 1 + 3 * m, N - 1, (uint64_t)pk->Z)) ||
      (rc = gpw_ec_generator_multiples_dev(ctx, 2, 1, m, (uint64_t)pk->B2))) {
    gpw_groth16_pk_free(pk);  // every failure after the allocations releases the key
    return rc;
  }
  pk->alpha1 = host_gen_mul<Fp>(seed + 1);
  pk->beta1 = host_gen_mul<Fp>(seed + 2);
  pk->delta1 = host_gen_mul<Fp>(seed + 3);
  pk->beta2 = host_gen_mul<Fp2>(seed + 2);
  pk->delta2 = host_gen_mul<Fp2>(seed + 3);
  if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) {
    set_error("pk_synthetic: %s", cudaGetErrorString(cudaGetLastError()));
    gpw_groth16_pk_free(pk);
    return GPW_ECUDA;
  }
  *out = pk;
  return GPW_OK;
}

extern "C" int gpw_groth16_pk_info(const gpw_pk* pk, uint64_t* m, uint64_t* n_pub, int* logN) {
  if (!pk) return GPW_EINVAL;
  if (m) *m = pk->m;
  if (n_pub) *n_pub = pk->n_pub;
  if (logN) *logN = pk->logN;
  return GPW_OK;
}

// h = (A.B - C) / Z_H in coefficient form (natural order), left in a_dev. a, b, c: N evaluations on H (device,
// Montgomery), all three are clobbered.  gnark: computeH (3 iFFT, 3 coset FFT, pointwise, coset iFFT).
extern "C" int gpw_groth16_compute_h_dev(gpw_ctx* ctx, uint64_t a_dev, uint64_t b_dev, uint64_t c_dev, int logN) {
  if (!ctx || !a_dev || !b_dev || !c_dev) {
    set_error("compute_h: null argument");
    return GPW_EINVAL;
  }
  const size_t N = (size_t)1 << logN;
  if (logN < 0 || logN > 27) {
    set_error("compute_h: logN=%d out of range", logN);
    return GPW_EINVAL;
  }
  GPW_CUDA(cudaSetDevice(ctx->device));
  // gnark's computeH transforms a, b AND c to the coset g.H (3 inverse + 3 coset-forward transforms), forms
  // (A.B - C) / Z_H there and goes back (1 coset-inverse): 7 transforms. The same h with 6: let W = A B = lo + X^N hi.
  // The coset-inverse transform of the pointwise products A(g w^i) B(g w^i) is V = lo + g^N hi (degree < N, since
  // (g w^i)^N = g^N); on H the products are the c_i of a satisfied system, so C = lo + hi; hence
  //   h = (A B - C) / (X^N - 1) = hi = (V - C) / (g^N - 1),
  // and C (coefficients) is one inverse transform of c - its trip to the coset and back is the identity and is skipped.
  // Field element for field element gnark's h (the coset-inverse transform is linear); an unsatisfied system never
  // reaches this point as a proof (GPW_EUNSAT).
  // a, b: evaluations on H -> coefficients (bit-reversed) -> evaluations on the coset g.H (natural)
  uint64_t v[2] = {a_dev, b_dev};
  for (int i = 0; i < 2; i++) {
    GPW_TRY(gpw_ntt_fr_dev(ctx, v[i], logN, /*inverse*/ 1, /*coset*/ 0, /*in_bitrev*/ 0, /*out_bitrev*/ 1));
    GPW_TRY(gpw_ntt_fr_dev(ctx, v[i], logN, 0, 1, 1, 0));
  }
  k_fr_mul_inplace<<<div_up(N, 256), 256, 0, ctx->stream>>>((Fr*)a_dev, (const Fr*)b_dev, N);
  GPW_CHECK_LAUNCH();
  GPW_TRY(gpw_ntt_fr_dev(ctx, a_dev, logN, 1, 1, 0, 1));  // V, bit-reversed
  GPW_TRY(gpw_ntt_fr_dev(ctx, c_dev, logN, 1, 0, 0, 1));  // C, bit-reversed
  // Z_H(g w^k) = g^N - 1 on the whole coset
  Fr g = fr_from_u64_host(5);
  Fr gN = g;
  for (int i = 0; i < logN; i++) gN = sqr(gN);
  Fr zinv = inv(sub(gN, Fr::one()));
  k_h_finish_bitrev<<<div_up(N, 256), 256, 0, ctx->stream>>>((Fr*)a_dev, (const Fr*)c_dev, logN, zinv);
  GPW_CHECK_LAUNCH();
  ctx->launches += 2;
  return GPW_OK;
}

// Proof = (Ar, Bs, Krs). w: m wire values (device, Fr Montgomery). a, b, c: N = 2^logN evaluation vectors A.w, B.w,
// C.w padded with zeros (device; clobbered). r, s: the prover's blinding scalars (canonical, 4 x u64) - gnark
// samples them from crypto/rand; here they are an argument so that proofs are reproducible.
// out: Ar (G1 affine, 8 u64) | Bs (G2 affine, 16 u64) | Krs (G1 affine, 8 u64), Montgomery coordinates.
extern "C" int gpw_groth16_prove_dev(gpw_pk* pk, uint64_t w_dev, uint64_t a_dev, uint64_t b_dev, uint64_t c_dev,
                                     const uint64_t* r_canon, const uint64_t* s_canon, uint64_t* out) {
  if (!pk || !w_dev || !a_dev || !b_dev || !c_dev || !r_canon || !s_canon || !out) {
    set_error("groth16_prove: null argument");
    return GPW_EINVAL;
  }
  gpw_ctx* ctx = pk->ctx;
  GPW_CUDA(cudaSetDevice(ctx->device));
  const size_t N = (size_t)1 << pk->logN;
  if (!pk->h0 && (cudaEventCreate(&pk->h0) != cudaSuccess || cudaEventCreate(&pk->h1) != cudaSuccess)) {
    set_error("cudaEventCreate failed");
    return GPW_ECUDA;
  }
  cudaEvent_t h0 = pk->h0, h1 = pk->h1;
  GPW_CUDA(cudaEventRecord(h0, ctx->stream));
  GPW_TRY(gpw_groth16_compute_h_dev(ctx, a_dev, b_dev, c_dev, pk->logN));
  GPW_CUDA(cudaEventRecord(h1, ctx->stream));
  G1Affine mA, mB1, mK, mZ;
  G2Affine mB2;
  GPW_TRY(gpw_msm_g1_dev(ctx, w_dev, (uint64_t)pk->A, pk->m, 1, 0, 0, 0, (uint64_t*)&mA));
  pk->t_msm_ms[0] = ctx->msm_total_ms;
  GPW_TRY(gpw_msm_g1_dev(ctx, w_dev, (uint64_t)pk->B1, pk->m, 1, 0, 0, 0, (uint64_t*)&mB1));
  pk->t_msm_ms[1] = ctx->msm_total_ms;
  GPW_TRY(gpw_msm_g2_dev(ctx, w_dev, (uint64_t)pk->B2, pk->m, 1, 0, 0, 0, (uint64_t*)&mB2));
  pk->t_msm_ms[2] = ctx->msm_total_ms;
  GPW_TRY(gpw_msm_g1_dev(ctx, w_dev + pk->n_pub * sizeof(Fr), (uint64_t)(pk->K + pk->n_pub), pk->m - pk->n_pub, 1, 0, 0, 0,
                         (uint64_t*)&mK));
  pk->t_msm_ms[3] = ctx->msm_total_ms;
  GPW_TRY(gpw_msm_g1_dev(ctx, a_dev, (uint64_t)pk->Z, N - 1, 1, 0, 0, 0, (uint64_t*)&mZ));
  pk->t_msm_ms[4] = ctx->msm_total_ms;
  GPW_CUDA(cudaEventElapsedTime(&pk->t_h_ms, h0, h1));
  // host assembly (a few hundred group operations)
  uint32_t rw[8], sw[8];
  memcpy(rw, r_canon, 32);
  memcpy(sw, s_canon, 32);
  // Ar = alpha + sum w_i A_i + r delta
  G1XYZZ Ar = G1XYZZ::from_affine(pk->alpha1);
  add_mixed(Ar, mA, false);
  add_full(Ar, host_scalar_mul(pk->delta1, rw));
  // Bs1 = beta + sum w_i B1_i + s delta   (G1) ; Bs = same in G2
  G1XYZZ Bs1 = G1XYZZ::from_affine(pk->beta1);
  add_mixed(Bs1, mB1, false);
  add_full(Bs1, host_scalar_mul(pk->delta1, sw));
  G2XYZZ Bs = G2XYZZ::from_affine(pk->beta2);
  add_mixed(Bs, mB2, false);
  add_full(Bs, host_scalar_mul(pk->delta2, sw));
  // Krs = sum_{private} w_i K_i + sum h_j Z_j + s Ar + r Bs1 - r s delta
  G1Affine ArA = to_affine(Ar), Bs1A = to_affine(Bs1);
  G1XYZZ Krs = G1XYZZ::from_affine(mK);
  add_mixed(Krs, mZ, false);
  add_full(Krs, host_scalar_mul(ArA, sw));
  add_full(Krs, host_scalar_mul(Bs1A, rw));
  Fr rf, sf;
  memcpy(&rf, r_canon, 32);
  memcpy(&sf, s_canon, 32);
  Fr rs = from_mont(mul(to_mont(rf), to_mont(sf)));
  uint32_t rsw[8];
  memcpy(rsw, &rs, 32);
  G1XYZZ rsd = host_scalar_mul(pk->delta1, rsw);
  add_full(Krs, neg(rsd));
  G1Affine KrsA = to_affine(Krs);
  G2Affine BsA = to_affine(Bs);
  memcpy(out, &ArA, sizeof(ArA));
  memcpy(out + 8, &BsA, sizeof(BsA));
  memcpy(out + 24, &KrsA, sizeof(KrsA));
  return GPW_OK;
}

extern "C" int gpw_groth16_last_stats(const gpw_pk* pk, float* h_ms, float* msm_ms5) {
  if (!pk) return GPW_EINVAL;
  if (h_ms) *h_ms = pk->t_h_ms;
  if (msm_ms5)
    for (int i = 0; i < 5; i++) msm_ms5[i] = pk->t_msm_ms[i];
  return GPW_OK;
}
