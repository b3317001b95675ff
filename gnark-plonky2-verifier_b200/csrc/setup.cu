// Groth16 trusted setup for a compiled circuit, key / proof serialisation.
//
// Replaces, for the wrap circuit,
//     pk, vk, err = groth16.Setup(r1cs)          /root/reference/benchmark.go:217
//     pk.WriteTo / vk.WriteTo                     /root/reference/benchmark.go:224-232
//     proof.WriteRawTo                            /root/reference/benchmark.go:272-274
// (gnark v0.9.1 backend/groth16/bn254/setup.go + marshal.go, un-vendored: the algorithm below is the published
// Groth16 setup with gnark's BSB22 commitment extension, restated from the paper and from gnark's documented field
// order; the byte layouts are "recalled" and flagged as such in INTEGRATION.md.)
//
//   toxic waste  tau, alpha, beta, gamma, delta (+ sigma, rho for the Pedersen key of the range-check commitment)
//   L_j(tau)     Lagrange basis of the size-N domain at tau                        (host, O(N), batched inversion)
//   A_i, B_i, C_i = sum_j M[j][i] L_j(tau)  over the R1CS rows                     (host threads, wire-partitioned)
//   pk.A = [A_i]1 (wires occurring in some L row), pk.B = [B_i]1 / [B_i]2 (wires occurring in some R row),
//   pk.K_i = [(beta A_i + alpha B_i + C_i)/delta]1 for private, non-committed wires,
//   ck_i   = [(beta A_i + alpha B_i + C_i)/gamma]1 for the committed wires (+ [sigma ck_i]1 for the proof of knowledge),
//   vk.K_i = the same over gamma for ONE, the public inputs and the commitment-challenge wire,
//   pk.Z_j = [tau^j (tau^N - 1)/delta]1.
// The ~35 M scalar multiplications of the generators run on the GPU: k_fixed_base_mul, 16 additions per scalar from
// a table of the 16 x 65536 multiples d 2^(16 w) G.
#include <sys/random.h>

#include <atomic>
#include <cstring>
#include <functional>
#include <thread>

#include "host/frontend.h"
#include "host_ec.cuh"
#include "wrap_internal.cuh"

const gpw::fe::API* gpw_circuit_api_internal(const gpw_circuit* c);
extern "C" void gpw_wrap_key_free(gpw_wrap_key* k);

namespace gpw {
namespace h64 {  // host-only 4 x 64-bit Montgomery arithmetic in Fr (same memory layout as Fe<FrParams>)

typedef unsigned __int128 u128;
struct F {
  uint64_t v[4];
};
static const uint64_t Q[4] = {0x43e1f593f0000001ull, 0x2833e84879b97091ull, 0xb85045b68181585dull, 0x30644e72e131a029ull};
static uint64_t qinv_neg() {
  uint64_t x = 1;
  for (int i = 0; i < 7; i++) x *= 2 - Q[0] * x;  // Newton: x = Q[0]^-1 mod 2^64
  return (uint64_t)0 - x;
}
static const uint64_t NINV = qinv_neg();

static inline bool geq_q(const uint64_t* t) {
  for (int i = 3; i >= 0; i--) {
    if (t[i] > Q[i]) return true;
    if (t[i] < Q[i]) return false;
  }
  return true;
}
static inline void sub_q(uint64_t* t) {
  u128 b = 0;
  for (int i = 0; i < 4; i++) {
    u128 d = (u128)t[i] - Q[i] - (uint64_t)b;
    t[i] = (uint64_t)d;
    b = (d >> 64) & 1;
  }
}
static inline F mul(const F& a, const F& b) {
  uint64_t t[6] = {0, 0, 0, 0, 0, 0};
  for (int i = 0; i < 4; i++) {
    uint64_t c = 0;
    for (int j = 0; j < 4; j++) {
      u128 s = (u128)a.v[j] * b.v[i] + t[j] + c;
      t[j] = (uint64_t)s;
      c = (uint64_t)(s >> 64);
    }
    u128 s = (u128)t[4] + c;
    t[4] = (uint64_t)s;
    t[5] = (uint64_t)(s >> 64);
    uint64_t m = t[0] * NINV;
    s = (u128)m * Q[0] + t[0];
    c = (uint64_t)(s >> 64);
    for (int j = 1; j < 4; j++) {
      s = (u128)m * Q[j] + t[j] + c;
      t[j - 1] = (uint64_t)s;
      c = (uint64_t)(s >> 64);
    }
    s = (u128)t[4] + c;
    t[3] = (uint64_t)s;
    t[4] = t[5] + (uint64_t)(s >> 64);
  }
  if (t[4] || geq_q(t)) sub_q(t);
  return {{t[0], t[1], t[2], t[3]}};
}
static inline F add(const F& a, const F& b) {
  uint64_t t[4];
  u128 c = 0;
  for (int i = 0; i < 4; i++) {
    c += (u128)a.v[i] + b.v[i];
    t[i] = (uint64_t)c;
    c >>= 64;
  }
  if (c || geq_q(t)) sub_q(t);
  return {{t[0], t[1], t[2], t[3]}};
}
static inline F neg(const F& a) {
  if (!(a.v[0] | a.v[1] | a.v[2] | a.v[3])) return a;
  uint64_t t[4];
  u128 b = 0;
  for (int i = 0; i < 4; i++) {
    u128 d = (u128)Q[i] - a.v[i] - (uint64_t)b;
    t[i] = (uint64_t)d;
    b = (d >> 64) & 1;
  }
  return {{t[0], t[1], t[2], t[3]}};
}
static inline F sub(const F& a, const F& b) { return add(a, neg(b)); }
static inline bool is_zero(const F& a) { return !(a.v[0] | a.v[1] | a.v[2] | a.v[3]); }
static inline F from_fr(const Fr& a) {
  F r;
  memcpy(&r, &a, 32);
  return r;
}
static inline Fr to_fr(const F& a) {
  Fr r;
  memcpy(&r, &a, 32);
  return r;
}
static const F ONE = from_fr(Fr::one());
static const F ZERO = {{0, 0, 0, 0}};
static inline F from_u64(uint64_t v) { return from_fr(fr_from_u64_host(v)); }
static F pow_u64(F a, uint64_t e) {
  F r = ONE;
  while (e) {
    if (e & 1) r = mul(r, a);
    a = mul(a, a);
    e >>= 1;
  }
  return r;
}
static F inv(const F& a) {  // a^(q-2)
  uint64_t e[4] = {Q[0] - 2, Q[1], Q[2], Q[3]};
  F r = ONE, b = a;
  for (int w = 0; w < 4; w++)
    for (int i = 0; i < 64; i++) {
      if ((e[w] >> i) & 1) r = mul(r, b);
      b = mul(b, b);
    }
  return r;
}
// in-place batched inversion (Montgomery's trick); zeros stay zero
static void batch_inv(F* a, size_t n, std::vector<F>& tmp) {
  tmp.resize(n);
  F run = ONE;
  for (size_t i = 0; i < n; i++) {
    tmp[i] = run;
    if (!is_zero(a[i])) run = mul(run, a[i]);
  }
  F r = inv(run);
  for (size_t i = n; i-- > 0;) {
    if (is_zero(a[i])) continue;
    F t = mul(r, tmp[i]);
    r = mul(r, a[i]);
    a[i] = t;
  }
}

}  // namespace h64

// ---- GPU: bulk multiplication of a fixed generator --------------------------------------------------------------------
constexpr int FB_C = 16, FB_W = 16, FB_ROW = 1 << FB_C;

template <class F, int PER>
__global__ void __launch_bounds__(128) k_base_multiples(Affine<F> g, size_t n, Affine<F>* __restrict__ out) {
  // out[i] = [i] g, i < n; PER consecutive multiples per thread
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t i0 = t * PER;
  if (i0 >= n) return;
  XYZZ<F> cur = XYZZ<F>::inf();
  for (int b = 31; b >= 0; b--) {
    cur = dbl(cur);
    if ((i0 >> b) & 1ull) add_mixed(cur, g, false);
  }
  int cnt = (int)min((size_t)PER, n - i0);
#pragma unroll 1
  for (int j = 0; j < cnt; j++) {
    out[i0 + j] = to_affine(cur);
    add_mixed(cur, g, false);
  }
}

// out[i] = [s_i] G from the table T[w][d] = [d 2^(16 w)] G. scalars: Fr, Montgomery.
template <class F>
__global__ void __launch_bounds__(128) k_fixed_base_mul(const Affine<F>* __restrict__ table, const Fr* __restrict__ scalars, size_t n,
                                                        Affine<F>* __restrict__ out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Fr s = from_mont(scalars[i]);
  XYZZ<F> acc = XYZZ<F>::inf();
#pragma unroll 1
  for (int w = 0; w < FB_W; w++) {
    uint32_t d = (s.l[w >> 1] >> (16 * (w & 1))) & 0xffffu;
    if (d) add_mixed(acc, table[(size_t)w * FB_ROW + d], false);
  }
  out[i] = to_affine(acc);
}

// affine Montgomery points <-> gnark-crypto raw big-endian coordinates. G1: X | Y (64 B). G2: X.A1 | X.A0 | Y.A1 | Y.A0 (128 B).
// A point at infinity is written as 0x40 followed by zeros (gnark-crypto's mUncompressedInfinity flag).
__device__ __host__ inline void fp_to_be(const Fp& mont, uint8_t* out) {
  Fp c = from_mont(mont);
  for (int i = 0; i < 8; i++)
    for (int b = 0; b < 4; b++) out[31 - (4 * i + b)] = (uint8_t)(c.l[i] >> (8 * b));
}
__device__ __host__ inline Fp fp_from_be(const uint8_t* in) {
  Fp c = Fp::zero();
  for (int i = 0; i < 8; i++)
    for (int b = 0; b < 4; b++) c.l[i] |= (uint32_t)in[31 - (4 * i + b)] << (8 * b);
  return to_mont(c);
}
__device__ __host__ inline void g1_to_raw(const G1Affine& p, uint8_t* out) {
  fp_to_be(p.x, out);
  fp_to_be(p.y, out + 32);
  if (p.is_inf()) out[0] = 0x40;
}
__device__ __host__ inline G1Affine g1_from_raw(const uint8_t* in) {
  if ((in[0] & 0xC0) == 0x40) return {Fp::zero(), Fp::zero()};
  return {fp_from_be(in), fp_from_be(in + 32)};
}
__device__ __host__ inline void g2_to_raw(const G2Affine& p, uint8_t* out) {
  fp_to_be(p.x.c1, out);
  fp_to_be(p.x.c0, out + 32);
  fp_to_be(p.y.c1, out + 64);
  fp_to_be(p.y.c0, out + 96);
  if (p.is_inf()) out[0] = 0x40;
}
__device__ __host__ inline G2Affine g2_from_raw(const uint8_t* in) {
  if ((in[0] & 0xC0) == 0x40) return {Fp2::zero(), Fp2::zero()};
  return {{fp_from_be(in + 32), fp_from_be(in)}, {fp_from_be(in + 96), fp_from_be(in + 64)}};
}
__global__ void k_g1_to_raw(const G1Affine* __restrict__ p, size_t n, uint8_t* __restrict__ out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) g1_to_raw(p[i], out + 64 * i);
}
__global__ void k_g1_from_raw(const uint8_t* __restrict__ in, size_t n, G1Affine* __restrict__ p) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = g1_from_raw(in + 64 * i);
}
__global__ void k_g2_to_raw(const G2Affine* __restrict__ p, size_t n, uint8_t* __restrict__ out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) g2_to_raw(p[i], out + 128 * i);
}
__global__ void k_g2_from_raw(const uint8_t* __restrict__ in, size_t n, G2Affine* __restrict__ p) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = g2_from_raw(in + 128 * i);
}

}  // namespace gpw

using namespace gpw;
using h64::F;

namespace {

struct DevBuf {
  void* p = nullptr;
  ~DevBuf() { cudaFree(p); }
  int alloc(size_t bytes) {
    cudaError_t e = cudaMalloc(&p, bytes ? bytes : 16);
    if (e != cudaSuccess) {
      set_error("setup: cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
      p = nullptr;
      return GPW_ENOMEM;
    }
    return GPW_OK;
  }
};

// [s_i] G for n scalars (host, Montgomery) -> device points
template <class Fq>
int bulk_generator_mul(gpw_ctx* ctx, const Affine<Fq>* table_dev, const F* scalars, size_t n, Affine<Fq>* out_dev, Fr* stage_dev) {
  if (!n) return GPW_OK;
  GPW_CUDA(cudaMemcpyAsync(stage_dev, scalars, n * 32, cudaMemcpyHostToDevice, ctx->stream));
  k_fixed_base_mul<Fq><<<div_up(n, 128), 128, 0, ctx->stream>>>(table_dev, stage_dev, n, out_dev);
  GPW_CHECK_LAUNCH();
  ctx->launches++;
  GPW_CUDA(cudaStreamSynchronize(ctx->stream));  // `scalars` is pageable host memory reused by the caller
  return GPW_OK;
}

template <class Fq>
int build_generator_table(gpw_ctx* ctx, Affine<Fq>* table_dev) {
  XYZZ<Fq> g = XYZZ<Fq>::from_affine(generator<Fq>());
  for (int w = 0; w < FB_W; w++) {
    Affine<Fq> gw = to_affine(g);
    k_base_multiples<Fq, 32><<<div_up(div_up(FB_ROW, 32), 128), 128, 0, ctx->stream>>>(gw, FB_ROW, table_dev + (size_t)w * FB_ROW);
    GPW_CHECK_LAUNCH();
    ctx->launches++;
    for (int i = 0; i < FB_C; i++) g = dbl(g);
  }
  return GPW_OK;
}

template <class Fq>
Affine<Fq> host_gen_mul_fr(const F& s_mont) {
  Fr c = from_mont(h64::to_fr(s_mont));
  return to_affine(host_scalar_mul(generator<Fq>(), c.l));
}

F derive_scalar(const uint8_t seed[32], const char* label) {
  uint8_t msg[64];
  memcpy(msg, seed, 32);
  memset(msg + 32, 0, 32);
  strncpy((char*)msg + 32, label, 31);
  uint64_t out[4];
  F r;
  for (uint8_t ctr = 0;; ctr++) {  // never zero (a zero toxic value would make the key degenerate)
    msg[63] = ctr;
    hash_to_fr(msg, 64, "gpw-groth16-setup", out);
    r = h64::from_fr(fe::fr_from_limbs(out));
    if (!h64::is_zero(r)) return r;
  }
}

void parallel_for(size_t n, const std::function<void(size_t, size_t, int)>& fn) {
  int T = (int)std::thread::hardware_concurrency();
  if (const char* e = getenv("GPW_SETUP_THREADS")) T = atoi(e);
  T = std::max(1, std::min(T, 64));
  if (n < 4096) T = 1;
  std::vector<std::thread> th;
  for (int t = 0; t < T; t++) {
    size_t lo = n * t / T, hi = n * (t + 1) / T;
    th.emplace_back(fn, lo, hi, t);
  }
  for (auto& x : th) x.join();
}

}  // namespace

// Real Groth16 setup for `circ` (see the header comment). seed32: 32 bytes from which the toxic waste is derived
// (reproducible keys for tests / ceremonies run elsewhere); NULL = fresh OS entropy (getrandom), discarded on return.
extern "C" int gpw_wrap_key_setup(gpw_ctx* ctx, gpw_circuit* circ, const uint8_t* seed32, gpw_wrap_key** out) {
  if (!ctx || !circ || !out) {
    set_error("wrap_key_setup: null argument");
    return GPW_EINVAL;
  }
  uint8_t seed[32];
  if (seed32) memcpy(seed, seed32, 32);
  else if (getrandom(seed, 32, 0) != 32) {
    set_error("wrap_key_setup: getrandom failed");
    return GPW_EINVAL;
  }
  const fe::API* api = gpw_circuit_api_internal(circ);
  gpw_wrap_key* k = nullptr;
  GPW_TRY(wrap_key_alloc(ctx, circ, &k));
  auto fail = [&](int code) {
    gpw_wrap_key_free(k);
    return code;
  };
  const size_t N = (size_t)1 << k->logN, m = k->m, n_cons = k->n_cons;
  const F tau = derive_scalar(seed, "tau"), alpha = derive_scalar(seed, "alpha"), beta = derive_scalar(seed, "beta"),
          gamma = derive_scalar(seed, "gamma"), delta = derive_scalar(seed, "delta"), sigma = derive_scalar(seed, "sigma"),
          rho = derive_scalar(seed, "rho");
  memset(seed, 0, sizeof(seed));
  // ---- Lagrange basis at tau: L_j = (tau^N - 1)/N . w^j / (tau - w^j) ---------------------------------------------------
  F w = h64::from_u64(5);  // gnark-crypto fr: multiplicative generator 5, 2-adicity 28
  {
    // w_N = 5^((r - 1) / N): exponent (r - 1) >> logN as 4 words
    uint64_t e[4] = {h64::Q[0] - 1, h64::Q[1], h64::Q[2], h64::Q[3]};
    for (int s = 0; s < k->logN; s++) {
      for (int i = 0; i < 3; i++) e[i] = (e[i] >> 1) | (e[i + 1] << 63);
      e[3] >>= 1;
    }
    F r = h64::ONE, b = w;
    for (int wd = 0; wd < 4; wd++)
      for (int i = 0; i < 64; i++) {
        if ((e[wd] >> i) & 1) r = h64::mul(r, b);
        b = h64::mul(b, b);
      }
    w = r;
  }
  F tauN = tau;
  for (int i = 0; i < k->logN; i++) tauN = h64::mul(tauN, tauN);
  const F zh = h64::sub(tauN, h64::ONE);  // Z_H(tau)
  if (h64::is_zero(zh)) {
    set_error("wrap_key_setup: tau lies in the evaluation domain");
    return fail(GPW_EINVAL);
  }
  const F lag_c = h64::mul(zh, h64::inv(h64::from_u64(N)));
  std::vector<F> lag(n_cons);
  parallel_for(n_cons, [&](size_t lo, size_t hi, int) {
    std::vector<F> wp(hi - lo), tmp;
    F cur = h64::pow_u64(w, lo);
    for (size_t j = lo; j < hi; j++) {
      wp[j - lo] = cur;
      lag[j] = h64::sub(tau, cur);
      cur = h64::mul(cur, w);
    }
    h64::batch_inv(lag.data() + lo, hi - lo, tmp);
    for (size_t j = lo; j < hi; j++) lag[j] = h64::mul(lag_c, h64::mul(wp[j - lo], lag[j]));
  });
  // ---- per-LE weights, then A_i / B_i / C_i --------------------------------------------------------------------------
  const auto& cons = api->Constraints();
  const auto& off = api->LeOffsets();
  const auto& lw = api->LeWires();
  const auto& lc = api->LeCoeffIds();
  const auto& coeffs = api->Coeffs();
  const size_t n_le = off.size() - 1;
  std::vector<F> W[3];
  std::vector<uint8_t> used(n_le, 0);
  for (int s = 0; s < 3; s++) W[s].assign(n_le, h64::ZERO);
  for (size_t j = 0; j < n_cons; j++)
    for (int s = 0; s < 3; s++) {
      const uint32_t le = cons[3 * j + s];
      W[s][le] = h64::add(W[s][le], lag[j]);
      used[le] |= (uint8_t)(1 << s);
    }
  std::vector<F>().swap(lag);
  std::vector<F> abc[3];
  for (int s = 0; s < 3; s++) abc[s].assign(m, h64::ZERO);
  std::vector<F> cf(coeffs.size());
  for (size_t i = 0; i < coeffs.size(); i++) cf[i] = h64::from_fr(coeffs[i]);
  parallel_for(m, [&](size_t wlo, size_t whi, int) {
    for (size_t le = 0; le < n_le; le++) {
      const uint8_t u = used[le];
      if (!u) continue;
      for (uint32_t t = off[le]; t < off[le + 1]; t++) {
        const uint32_t wire = lw[t];
        if (wire < wlo || wire >= whi) continue;
        const uint32_t ci = lc[t];
        for (int s = 0; s < 3; s++) {
          if (!((u >> s) & 1)) continue;
          F& acc = abc[s][wire];
          if (ci == fe::API::COEFF_ONE) acc = h64::add(acc, W[s][le]);
          else if (ci == fe::API::COEFF_NEG_ONE) acc = h64::sub(acc, W[s][le]);
          else acc = h64::add(acc, h64::mul(cf[ci], W[s][le]));
        }
      }
    }
  });
  for (int s = 0; s < 3; s++) std::vector<F>().swap(W[s]);
  // ---- scalars of every base ---------------------------------------------------------------------------------------
  const F gamma_inv = h64::inv(gamma), delta_inv = h64::inv(delta);
  const uint32_t c_lo = k->n_committed ? k->limb_start : (uint32_t)m, c_hi = c_lo + k->n_committed;
  const bool has_commit = k->n_committed != 0;
  auto is_vk_wire = [&](size_t i) { return i <= k->n_pub || (has_commit && i == k->commit_wire); };
  std::vector<F> kd(m), ck(k->n_committed), cks(k->n_committed);
  std::vector<F> vk_scal;
  parallel_for(m, [&](size_t lo, size_t hi, int) {
    for (size_t i = lo; i < hi; i++) {
      F t = h64::add(h64::add(h64::mul(beta, abc[0][i]), h64::mul(alpha, abc[1][i])), abc[2][i]);
      if (i >= c_lo && i < c_hi) {
        ck[i - c_lo] = h64::mul(t, gamma_inv);
        cks[i - c_lo] = h64::mul(ck[i - c_lo], sigma);
        kd[i] = h64::ZERO;
      } else if (is_vk_wire(i)) {
        kd[i] = h64::mul(t, gamma_inv);  // moved to vk below; the prover never touches pk.K at these wires
      } else {
        kd[i] = h64::mul(t, delta_inv);
      }
    }
  });
  for (size_t i = 0; i <= k->n_pub; i++) vk_scal.push_back(kd[i]);
  if (has_commit) vk_scal.push_back(kd[k->commit_wire]);
  for (size_t i = 0; i <= k->n_pub; i++) kd[i] = h64::ZERO;
  if (has_commit) kd[k->commit_wire] = h64::ZERO;
  // ---- bulk generator multiplications on the GPU --------------------------------------------------------------------
  DevBuf t1, t2, stage;
  int rc = 0;
  const size_t stage_n = std::max<size_t>(std::max<size_t>(m, N), 1);
  if ((rc = t1.alloc((size_t)FB_W * FB_ROW * sizeof(G1Affine))) || (rc = t2.alloc((size_t)FB_W * FB_ROW * sizeof(G2Affine))) ||
      (rc = stage.alloc(stage_n * 32)))
    return fail(rc);
  if ((rc = build_generator_table<Fp>(ctx, (G1Affine*)t1.p)) || (rc = build_generator_table<Fp2>(ctx, (G2Affine*)t2.p))) return fail(rc);
  std::vector<F> sc(std::max<size_t>(k->nA, std::max<size_t>(k->nB, N)));
  for (size_t j = 0; j < k->nA; j++) sc[j] = abc[0][k->suppA_host[j]];
  if ((rc = bulk_generator_mul<Fp>(ctx, (G1Affine*)t1.p, sc.data(), k->nA, k->A, (Fr*)stage.p))) return fail(rc);
  for (size_t j = 0; j < k->nB; j++) sc[j] = abc[1][k->suppB_host[j]];
  if ((rc = bulk_generator_mul<Fp>(ctx, (G1Affine*)t1.p, sc.data(), k->nB, k->B1, (Fr*)stage.p)) ||
      (rc = bulk_generator_mul<Fp2>(ctx, (G2Affine*)t2.p, sc.data(), k->nB, k->B2, (Fr*)stage.p)))
    return fail(rc);
  if ((rc = bulk_generator_mul<Fp>(ctx, (G1Affine*)t1.p, kd.data(), m, k->K, (Fr*)stage.p)) ||
      (rc = bulk_generator_mul<Fp>(ctx, (G1Affine*)t1.p, ck.data(), k->n_committed, k->CK, (Fr*)stage.p)) ||
      (rc = bulk_generator_mul<Fp>(ctx, (G1Affine*)t1.p, cks.data(), k->n_committed, k->CKs, (Fr*)stage.p)))
    return fail(rc);
  {
    const F zd = h64::mul(zh, delta_inv);
    parallel_for(N - 1, [&](size_t lo, size_t hi, int) {
      F cur = h64::mul(zd, h64::pow_u64(tau, lo));
      for (size_t j = lo; j < hi; j++) {
        sc[j] = cur;
        cur = h64::mul(cur, tau);
      }
    });
    if ((rc = bulk_generator_mul<Fp>(ctx, (G1Affine*)t1.p, sc.data(), N - 1, k->Z, (Fr*)stage.p))) return fail(rc);
  }
  // ---- the handful of named elements --------------------------------------------------------------------------------
  k->alpha1 = host_gen_mul_fr<Fp>(alpha);
  k->beta1 = host_gen_mul_fr<Fp>(beta);
  k->delta1 = host_gen_mul_fr<Fp>(delta);
  k->beta2 = host_gen_mul_fr<Fp2>(beta);
  k->delta2 = host_gen_mul_fr<Fp2>(delta);
  k->gamma2 = host_gen_mul_fr<Fp2>(gamma);
  for (const F& s : vk_scal) k->vkK.push_back(host_gen_mul_fr<Fp>(s));
  k->ped_g = host_gen_mul_fr<Fp2>(rho);
  k->ped_g_root_sigma_neg = host_gen_mul_fr<Fp2>(h64::neg(h64::mul(rho, h64::inv(sigma))));
  k->real = true;
  GPW_TRY(wrap_key_finish(k));  // frees the key on failure
  *out = k;
  return GPW_OK;
}

// ---- serialisation ----------------------------------------------------------------------------------------------------
static void put_u32(std::vector<uint8_t>& b, uint32_t v) {
  for (int i = 3; i >= 0; i--) b.push_back((uint8_t)(v >> (8 * i)));
}
static void put_g1(std::vector<uint8_t>& b, const G1Affine& p) {
  uint8_t t[64];
  g1_to_raw(p, t);
  b.insert(b.end(), t, t + 64);
}
static void put_g2(std::vector<uint8_t>& b, const G2Affine& p) {
  uint8_t t[128];
  g2_to_raw(p, t);
  b.insert(b.end(), t, t + 128);
}
static int emit(const std::vector<uint8_t>& b, uint8_t* out, size_t cap, size_t* len) {
  if (len) *len = b.size();
  if (!out) return GPW_OK;  // size query
  if (cap < b.size()) {
    set_error("output buffer too small: %zu < %zu", cap, b.size());
    return GPW_EINVAL;
  }
  memcpy(out, b.data(), b.size());
  return GPW_OK;
}

// gnark groth16 (bn254) Proof.WriteRawTo: Ar | Bs | Krs | uint32 n | n commitments | CommitmentPok - uncompressed
// big-endian coordinates, G2 as X.A1 | X.A0 | Y.A1 | Y.A0 (/root/reference/benchmark.go:272-291 reads the first 256 bytes
// in exactly this order). proof64 = the 64-word output of gpw_wrap_prove*.
extern "C" int gpw_wrap_proof_write_raw(const gpw_wrap_key* k, const uint64_t* proof64, uint8_t* out, size_t cap, size_t* len) {
  if (!k || !proof64) {
    set_error("proof_write_raw: null argument");
    return GPW_EINVAL;
  }
  G1Affine ar, krs, d, pok;
  G2Affine bs;
  memcpy(&ar, proof64, 64);
  memcpy(&bs, proof64 + 8, 128);
  memcpy(&krs, proof64 + 24, 64);
  memcpy(&d, proof64 + 32, 64);
  memcpy(&pok, proof64 + 40, 64);
  std::vector<uint8_t> b;
  put_g1(b, ar);
  put_g2(b, bs);
  put_g1(b, krs);
  put_u32(b, k->n_committed ? 1 : 0);
  if (k->n_committed) put_g1(b, d);
  put_g1(b, pok);
  return emit(b, out, cap, len);
}

// gnark groth16 (bn254) VerifyingKey.WriteRawTo: alpha1 | beta1 | beta2 | gamma2 | delta1 | delta2 | uint32 len(K) | K |
// uint32 n_commitments | per commitment: uint32 n, n x uint64 (public wires hashed with the commitment: none here) |
// Pedersen vk (G | GRootSigmaNeg) when there is a commitment.
extern "C" int gpw_wrap_key_vk_write_raw(const gpw_wrap_key* k, uint8_t* out, size_t cap, size_t* len) {
  if (!k) return GPW_EINVAL;
  if (!k->real) {
    set_error("vk_write_raw: the key comes from gpw_wrap_key_synthetic (DummySetup analogue) and has no verifying key");
    return GPW_EINVAL;
  }
  std::vector<uint8_t> b;
  put_g1(b, k->alpha1);
  put_g1(b, k->beta1);
  put_g2(b, k->beta2);
  put_g2(b, k->gamma2);
  put_g1(b, k->delta1);
  put_g2(b, k->delta2);
  put_u32(b, (uint32_t)k->vkK.size());
  for (const G1Affine& p : k->vkK) put_g1(b, p);
  put_u32(b, k->n_committed ? 1 : 0);
  if (k->n_committed) {
    put_u32(b, 0);
    put_g2(b, k->ped_g);
    put_g2(b, k->ped_g_root_sigma_neg);
  }
  return emit(b, out, cap, len);
}

// ---- proving key file: gnark ProvingKey.WriteRawTo field order ------------------------------------------------------------
//   alpha1 | beta1 | delta1 | A | B1 | Z | K | beta2 | delta2 | B2 | uint32 nbWires | uint32 nbInfinityA | uint32 nbInfinityB |
//   InfinityA | InfinityB | uint32 n_commitment_keys | per key: basis | basisExpSigma
// (slices = uint32 BE length + raw points; bool slices = uint32 length + one byte each). A / B hold only the wires whose
// polynomial is non-zero (our supports; the others are flagged in InfinityA / InfinityB), K only the private,
// non-committed wires - as in gnark. Wire numbering is that of OUR frontend, so a key is portable between processes of this
// library for the same compiled circuit, not to gnark's own R1CS of the reference circuit.
namespace {
struct FileW {
  FILE* f;
  bool ok = true;
  void bytes(const void* p, size_t n) { ok = ok && fwrite(p, 1, n, f) == n; }
  void u32(uint32_t v) {
    uint8_t b[4] = {(uint8_t)(v >> 24), (uint8_t)(v >> 16), (uint8_t)(v >> 8), (uint8_t)v};
    bytes(b, 4);
  }
};
struct FileR {
  FILE* f;
  bool ok = true;
  void bytes(void* p, size_t n) { ok = ok && fread(p, 1, n, f) == n; }
  uint32_t u32() {
    uint8_t b[4] = {0, 0, 0, 0};
    bytes(b, 4);
    return (uint32_t)b[0] << 24 | (uint32_t)b[1] << 16 | (uint32_t)b[2] << 8 | b[3];
  }
};
constexpr size_t IO_CHUNK = 1 << 18;  // points per staging round

template <class Fq>
int write_points(gpw_ctx* ctx, FileW& w, const Affine<Fq>* dev, size_t n, uint8_t* raw_dev, std::vector<uint8_t>& host) {
  const size_t sz = sizeof(Affine<Fq>);
  w.u32((uint32_t)n);
  for (size_t o = 0; o < n; o += IO_CHUNK) {
    size_t c = std::min(IO_CHUNK, n - o);
    if (sz == 64) k_g1_to_raw<<<div_up(c, 128), 128, 0, ctx->stream>>>((const G1Affine*)dev + o, c, raw_dev);
    else k_g2_to_raw<<<div_up(c, 128), 128, 0, ctx->stream>>>((const G2Affine*)dev + o, c, raw_dev);
    GPW_CHECK_LAUNCH();
    GPW_CUDA(cudaMemcpyAsync(host.data(), raw_dev, c * sz, cudaMemcpyDeviceToHost, ctx->stream));
    GPW_CUDA(cudaStreamSynchronize(ctx->stream));
    w.bytes(host.data(), c * sz);
  }
  return GPW_OK;
}
template <class Fq>
int read_points(gpw_ctx* ctx, FileR& r, Affine<Fq>* dev, size_t n_expect, uint8_t* raw_dev, std::vector<uint8_t>& host, const char* what) {
  const size_t sz = sizeof(Affine<Fq>);
  const uint32_t n = r.u32();
  if (!r.ok || n != n_expect) {
    set_error("wrap_key_load: %s has %u points, the circuit needs %zu", what, n, n_expect);
    return GPW_EINVAL;
  }
  for (size_t o = 0; o < n; o += IO_CHUNK) {
    size_t c = std::min(IO_CHUNK, (size_t)n - o);
    r.bytes(host.data(), c * sz);
    if (!r.ok) {
      set_error("wrap_key_load: truncated file in %s", what);
      return GPW_EINVAL;
    }
    GPW_CUDA(cudaMemcpyAsync(raw_dev, host.data(), c * sz, cudaMemcpyHostToDevice, ctx->stream));
    if (sz == 64) k_g1_from_raw<<<div_up(c, 128), 128, 0, ctx->stream>>>(raw_dev, c, (G1Affine*)dev + o);
    else k_g2_from_raw<<<div_up(c, 128), 128, 0, ctx->stream>>>(raw_dev, c, (G2Affine*)dev + o);
    GPW_CHECK_LAUNCH();
    GPW_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  return GPW_OK;
}
void file_put_g1(FileW& w, const G1Affine& p) {
  uint8_t t[64];
  g1_to_raw(p, t);
  w.bytes(t, 64);
}
void file_put_g2(FileW& w, const G2Affine& p) {
  uint8_t t[128];
  g2_to_raw(p, t);
  w.bytes(t, 128);
}
G1Affine file_get_g1(FileR& r) {
  uint8_t t[64] = {0};
  r.bytes(t, 64);
  return g1_from_raw(t);
}
G2Affine file_get_g2(FileR& r) {
  uint8_t t[128] = {0};
  r.bytes(t, 128);
  return g2_from_raw(t);
}
}  // namespace

// pk.WriteRawTo(proving.key) + vk.WriteRawTo(verifying.key) (/root/reference/benchmark.go:224-232). vk_path may be NULL.
extern "C" int gpw_wrap_key_save(const gpw_wrap_key* k, const char* pk_path, const char* vk_path) {
  if (!k || !pk_path) {
    set_error("wrap_key_save: null argument");
    return GPW_EINVAL;
  }
  gpw_ctx* ctx = k->ctx;
  GPW_CUDA(cudaSetDevice(ctx->device));
  FILE* f = fopen(pk_path, "wb");
  if (!f) {
    set_error("wrap_key_save: cannot open %s", pk_path);
    return GPW_EINVAL;
  }
  FileW w{f};
  DevBuf raw;
  int rc = raw.alloc(IO_CHUNK * 128);
  std::vector<uint8_t> host(IO_CHUNK * 128);
  const size_t N = (size_t)1 << k->logN;
  const uint32_t k_lo = 1 + k->n_pub, c_lo = k->n_committed ? k->limb_start : k->m;
  auto pts1 = [&](const G1Affine* p, size_t n) { return write_points<Fp>(ctx, w, p, n, (uint8_t*)raw.p, host); };
  if (!rc) {
    file_put_g1(w, k->alpha1);
    file_put_g1(w, k->beta1);
    file_put_g1(w, k->delta1);
  }
  if (!rc) rc = pts1(k->A, k->nA);
  if (!rc) rc = pts1(k->B1, k->nB);
  if (!rc) rc = pts1(k->Z, N - 1);
  if (!rc) {  // K: the two private, non-committed wire ranges as ONE slice
    const size_t n1 = c_lo - k_lo, n2 = k->m - k->k2_lo;
    w.u32((uint32_t)(n1 + n2));
    // (write_points writes its own length prefix: emit the ranges through a length-less variant)
    for (int part = 0; part < 2 && !rc; part++) {
      const G1Affine* base = part == 0 ? k->K + k_lo : k->K + k->k2_lo;
      const size_t n = part == 0 ? n1 : n2;
      for (size_t o = 0; o < n && !rc; o += IO_CHUNK) {
        size_t c = std::min(IO_CHUNK, n - o);
        k_g1_to_raw<<<div_up(c, 128), 128, 0, ctx->stream>>>(base + o, c, (uint8_t*)raw.p);
        if (cudaMemcpyAsync(host.data(), raw.p, c * 64, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess ||
            cudaStreamSynchronize(ctx->stream) != cudaSuccess) {
          set_error("wrap_key_save: %s", cudaGetErrorString(cudaGetLastError()));
          rc = GPW_ECUDA;
        }
        w.bytes(host.data(), c * 64);
      }
    }
  }
  if (!rc) {
    file_put_g2(w, k->beta2);
    file_put_g2(w, k->delta2);
    rc = write_points<Fp2>(ctx, w, k->B2, k->nB, (uint8_t*)raw.p, host);
  }
  if (!rc) {
    w.u32(k->m);
    w.u32(k->m - k->nA);
    w.u32(k->m - k->nB);
    std::vector<uint8_t> inf(k->m, 1);
    for (uint32_t i : k->suppA_host) inf[i] = 0;
    w.u32(k->m);
    w.bytes(inf.data(), inf.size());
    std::fill(inf.begin(), inf.end(), 1);
    for (uint32_t i : k->suppB_host) inf[i] = 0;
    w.u32(k->m);
    w.bytes(inf.data(), inf.size());
    w.u32(k->n_committed ? 1 : 0);
  }
  if (!rc && k->n_committed) {
    rc = pts1(k->CK, k->n_committed);
    if (!rc) rc = pts1(k->CKs, k->n_committed);
  }
  if (!rc && !w.ok) {
    set_error("wrap_key_save: write to %s failed", pk_path);
    rc = GPW_EINVAL;
  }
  fclose(f);
  if (rc) return rc;
  if (vk_path) {
    size_t len = 0;
    GPW_TRY(gpw_wrap_key_vk_write_raw(k, nullptr, 0, &len));
    std::vector<uint8_t> b(len);
    GPW_TRY(gpw_wrap_key_vk_write_raw(k, b.data(), b.size(), &len));
    FILE* g = fopen(vk_path, "wb");
    if (!g || fwrite(b.data(), 1, len, g) != len) {
      if (g) fclose(g);
      set_error("wrap_key_save: cannot write %s", vk_path);
      return GPW_EINVAL;
    }
    fclose(g);
  }
  return GPW_OK;
}

// Reads a key written by gpw_wrap_key_save for the SAME compiled circuit (shapes are checked). vk_path may be NULL (the key
// can then prove but not export a verifying key).
extern "C" int gpw_wrap_key_load(gpw_ctx* ctx, gpw_circuit* circ, const char* pk_path, const char* vk_path, gpw_wrap_key** out) {
  if (!ctx || !circ || !pk_path || !out) {
    set_error("wrap_key_load: null argument");
    return GPW_EINVAL;
  }
  FILE* f = fopen(pk_path, "rb");
  if (!f) {
    set_error("wrap_key_load: cannot open %s", pk_path);
    return GPW_EINVAL;
  }
  gpw_wrap_key* k = nullptr;
  int rc = wrap_key_alloc(ctx, circ, &k);
  if (rc) {
    fclose(f);
    return rc;
  }
  FileR r{f};
  DevBuf raw;
  rc = raw.alloc(IO_CHUNK * 128);
  std::vector<uint8_t> host(IO_CHUNK * 128);
  const size_t N = (size_t)1 << k->logN;
  const uint32_t k_lo = 1 + k->n_pub, c_lo = k->n_committed ? k->limb_start : k->m;
  auto pts1 = [&](G1Affine* p, size_t n, const char* what) { return read_points<Fp>(ctx, r, p, n, (uint8_t*)raw.p, host, what); };
  if (!rc) {
    k->alpha1 = file_get_g1(r);
    k->beta1 = file_get_g1(r);
    k->delta1 = file_get_g1(r);
  }
  if (!rc) rc = pts1(k->A, k->nA, "A");
  if (!rc) rc = pts1(k->B1, k->nB, "B1");
  if (!rc) rc = pts1(k->Z, N - 1, "Z");
  if (!rc) {
    const size_t n1 = c_lo - k_lo, n2 = k->m - k->k2_lo;
    const uint32_t n = r.u32();
    if (!r.ok || n != n1 + n2) {
      set_error("wrap_key_load: K has %u points, the circuit needs %zu", n, n1 + n2);
      rc = GPW_EINVAL;
    }
    if (!rc) rc = cudaMemsetAsync(k->K, 0, (size_t)k->m * sizeof(G1Affine), ctx->stream) == cudaSuccess ? GPW_OK : GPW_ECUDA;
    for (int part = 0; part < 2 && !rc; part++) {
      G1Affine* base = part == 0 ? k->K + k_lo : k->K + k->k2_lo;
      const size_t np = part == 0 ? n1 : n2;
      for (size_t o = 0; o < np && !rc; o += IO_CHUNK) {
        size_t c = std::min(IO_CHUNK, np - o);
        r.bytes(host.data(), c * 64);
        if (!r.ok) {
          set_error("wrap_key_load: truncated file in K");
          rc = GPW_EINVAL;
          break;
        }
        if (cudaMemcpyAsync(raw.p, host.data(), c * 64, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess) rc = GPW_ECUDA;
        k_g1_from_raw<<<div_up(c, 128), 128, 0, ctx->stream>>>((const uint8_t*)raw.p, c, base + o);
        if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) rc = GPW_ECUDA;
        if (rc) set_error("wrap_key_load: %s", cudaGetErrorString(cudaGetLastError()));
      }
    }
  }
  if (!rc) {
    k->beta2 = file_get_g2(r);
    k->delta2 = file_get_g2(r);
    rc = read_points<Fp2>(ctx, r, k->B2, k->nB, (uint8_t*)raw.p, host, "B2");
  }
  if (!rc) {
    const uint32_t nw = r.u32(), nia = r.u32(), nib = r.u32();
    if (!r.ok || nw != k->m || nia != k->m - k->nA || nib != k->m - k->nB) {
      set_error("wrap_key_load: key is for a different circuit (wires %u/%u, infinityA %u/%u, infinityB %u/%u)", nw, k->m, nia,
                k->m - k->nA, nib, k->m - k->nB);
      rc = GPW_EINVAL;
    }
  }
  if (!rc) {
    std::vector<uint8_t> inf(k->m);
    for (int side = 0; side < 2 && !rc; side++) {
      const uint32_t n = r.u32();
      if (n != k->m) rc = GPW_EINVAL;
      else r.bytes(inf.data(), n);
      const auto& supp = side == 0 ? k->suppA_host : k->suppB_host;
      size_t zeros = 0;
      for (uint32_t i = 0; i < k->m && !rc; i++) zeros += inf[i] == 0;
      for (uint32_t i : supp)
        if (inf[i]) rc = GPW_EINVAL;
      if (!rc && zeros != supp.size()) rc = GPW_EINVAL;
      if (rc) set_error("wrap_key_load: Infinity%c does not match the circuit's support", side ? 'B' : 'A');
    }
  }
  if (!rc) {
    const uint32_t nk = r.u32();
    if (!r.ok || nk != (k->n_committed ? 1u : 0u)) {
      set_error("wrap_key_load: %u commitment keys, the circuit has %u", nk, k->n_committed ? 1u : 0u);
      rc = GPW_EINVAL;
    }
  }
  if (!rc && k->n_committed) {
    rc = pts1(k->CK, k->n_committed, "commitment basis");
    if (!rc) rc = pts1(k->CKs, k->n_committed, "commitment basisExpSigma");
  }
  fclose(f);
  if (!rc && vk_path) {
    FILE* g = fopen(vk_path, "rb");
    if (!g) {
      set_error("wrap_key_load: cannot open %s", vk_path);
      rc = GPW_EINVAL;
    } else {
      FileR v{g};
      file_get_g1(v);  // alpha1, beta1: already in the pk
      file_get_g1(v);
      file_get_g2(v);  // beta2
      k->gamma2 = file_get_g2(v);
      file_get_g1(v);
      file_get_g2(v);
      const uint32_t nk = v.u32();
      if (!v.ok || nk != 1 + k->n_pub + (k->n_committed ? 1 : 0)) {
        set_error("wrap_key_load: verifying key has %u public bases, the circuit needs %u", nk, 1 + k->n_pub + (k->n_committed ? 1 : 0));
        rc = GPW_EINVAL;
      }
      for (uint32_t i = 0; i < nk && !rc; i++) k->vkK.push_back(file_get_g1(v));
      if (!rc) {
        const uint32_t nc = v.u32();
        for (uint32_t i = 0; i < nc; i++) {
          const uint32_t n = v.u32();
          for (uint32_t j = 0; j < 2 * n; j++) v.u32();
        }
        if (nc) {
          k->ped_g = file_get_g2(v);
          k->ped_g_root_sigma_neg = file_get_g2(v);
        }
        if (!v.ok) {
          set_error("wrap_key_load: truncated verifying key");
          rc = GPW_EINVAL;
        }
      }
      fclose(g);
      k->real = !rc;
    }
  }
  if (rc) {
    gpw_wrap_key_free(k);
    return rc;
  }
  GPW_TRY(wrap_key_finish(k));
  *out = k;
  return GPW_OK;
}
