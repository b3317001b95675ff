// PLONK / KZG backend over BN254 for circuits compiled by this library (BASELINE.json configs[3]).
//
// Replaces, for `-proof-system plonk` of the reference,
//     srs, _ := test.NewKZGSRS(r1cs)                        /root/reference/benchmark.go:105
//     pk, vk, _ := plonk.Setup(r1cs, srs)                   /root/reference/benchmark.go:130
//     proof, _ := plonk.Prove(r1cs, pk, witness)            /root/reference/benchmark.go:162
// (gnark v0.9.1 backend/plonk/bn254, un-vendored - go.mod:6). The constraint system is the lowering of the compiled circuit
// in host/scs.{h,cc}; the protocol is PLONK as published (Gabizon, Williamson, Ciobotaru 2019) with gnark's BSB22 commitment
// column (Qcp . P2) for the range-check challenge:
//   round 0  solve phase 1 on the GPU; P2 = the committed wires on their rows; [P2]; challenge X = hash_to_field([P2]) (as in
//            the Groth16 path); solve phase 2; chain variables (k_scs_chains)
//   round 1  a, b, c = the three wire columns, [a], [b], [c]
//   round 2  beta, gamma; Z = grand product of (w + beta k_col w^i + gamma) / (w + beta S_col + gamma), [Z]
//   round 3  alpha; quotient t = (gate + alpha perm + alpha^2 (Z - 1) L_0) / Z_H on four cosets of size N (the selectors',
//            sigmas' and L_0's coset evaluations are precomputed at setup: 6 transforms per coset and proof), split into
//            t_0, t_1, t_2 of degree < N, [t_0], [t_1], [t_2]
//   round 4  zeta; all 17 polynomials evaluated at zeta, Z at zeta w
//   round 5  nu; ONE batched opening at zeta and one at zeta w (W evaluated on H, interpolated, committed)
// Differences from gnark's implementation, stated because proof bytes cannot be compared with gnark here anyway (no Go
// toolchain): every polynomial is opened at zeta (no linearisation polynomial - 18 field elements instead of 7), no blinding
// factors (the proof is not zero-knowledge), own transcript labels (SHA-256 / RFC 9380 hash-to-field, the library's
// gpw_hash_to_fr). oracle/plonk_verify.py is the verifier the tests check every proof with.
//
// All polynomial work is device resident: NTTs through gpw_ntt_fr_dev, commitments through gpw_msm_g1_dev with the SRS as
// bases, batched inversions with Montgomery's trick (8 per thread).
#include <sys/random.h>

#include <cstring>
#include <string>
#include <vector>

#include "common.cuh"
#include "ec.cuh"
#include "host/frontend.h"
#include "host/scs.h"
#include "host_ec.cuh"
#include "wrap_internal.cuh"

const gpw::fe::API* gpw_circuit_api_internal(const gpw_circuit* c);
extern "C" {
int gpw_circuit_info(const gpw_circuit* c, uint64_t* info16);
int gpw_witness_solve_phase1_on(gpw_circuit* c, gpw_ctx* lane, uint64_t inputs_dev, int n_proofs, uint64_t wires_dev, size_t wire_stride);
int gpw_witness_solve_phase2_on(gpw_circuit* c, gpw_ctx* lane, const uint64_t* ch, int n_proofs, uint64_t wires_dev, size_t wire_stride);
int gpw_ntt_fr_dev(gpw_ctx* ctx, uint64_t data_dev, int logn, int inverse, int coset, int in_bitrev, int out_bitrev);
int gpw_msm_g1_dev(gpw_ctx* ctx, uint64_t s, uint64_t p, size_t n, int mont, int c, int lo, int hi, uint64_t* out);
int gpw_msm_g1_fixed_table(gpw_ctx* ctx, uint64_t points_dev, size_t n, int window_bits, int n_windows, uint64_t table_dev);
int gpw_msm_g1_fixed_dev(gpw_ctx* ctx, uint64_t s, uint64_t table, size_t n, int mont, int c, int n_windows, uint64_t* out);
}

namespace gpw {
namespace plonk {

__device__ __forceinline__ Fr ldf(const Fr* p) {
  Fr r;
  const uint4* s = reinterpret_cast<const uint4*>(p);
  uint4* d = reinterpret_cast<uint4*>(&r);
  d[0] = s[0];
  d[1] = s[1];
  return r;
}
__device__ __forceinline__ void stf(Fr* p, const Fr& v) {
  const uint4* s = reinterpret_cast<const uint4*>(&v);
  uint4* d = reinterpret_cast<uint4*>(p);
  d[0] = s[0];
  d[1] = s[1];
}
__host__ __device__ inline Fr pow_u64(Fr a, uint64_t e) {
  uint32_t w[2] = {(uint32_t)e, (uint32_t)(e >> 32)};
  return pow_words(a, w, 2);
}

// ---- witness extension: one thread per chain of a level ---------------------------------------------------------------
__global__ void k_scs_chains(const scs::Chain* __restrict__ chains, uint32_t n, const uint32_t* __restrict__ cw,
                             const uint32_t* __restrict__ cc, const Fr* __restrict__ coeffs, Fr* __restrict__ v) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const scs::Chain ch = chains[i];
  auto term = [&](uint32_t t) {
    const uint32_t ci = cc[ch.term_off + t];
    const Fr x = ldf(v + cw[ch.term_off + t]);
    return ci == scs::System::C_ONE ? x : ci == scs::System::C_NEG_ONE ? neg(x) : mul(ldf(coeffs + ci), x);
  };
  Fr acc = term(0);
  if (ch.konst != scs::System::C_ZERO) acc = add(acc, ldf(coeffs + ch.konst));
  if (ch.n_terms == 1) stf(v + ch.out, acc);
  for (uint32_t j = 1; j < ch.n_terms; j++) {
    acc = add(acc, term(j));
    stf(v + ch.out + j - 1, acc);
  }
}

// out[row] = v[ids[row]] for row < n_rows, v[0] beyond (padding rows hold variable 0 in every column)
__global__ void k_gather_col(const Fr* __restrict__ v, const uint32_t* __restrict__ ids, uint32_t n_rows, uint32_t N, Fr* __restrict__ out) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  stf(out + i, ldf(v + (i < n_rows ? ids[i] : 0u)));
}
// selector evaluations: coeffs[q[row]], 0 on the padding rows
__global__ void k_coeff_col(const Fr* __restrict__ coeffs, const uint32_t* __restrict__ q, uint32_t n_rows, uint32_t N, Fr* __restrict__ out) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  stf(out + i, i < n_rows ? ldf(coeffs + q[i]) : Fr::zero());
}
__global__ void k_flag_col(const uint8_t* __restrict__ f, uint32_t n_rows, uint32_t N, Fr* __restrict__ out) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  stf(out + i, (i < n_rows && f[i]) ? Fr::one() : Fr::zero());
}
// P2 on H: the committed wire on its Qcp row, 0 elsewhere
__global__ void k_p2_evals(const Fr* __restrict__ v, const uint32_t* __restrict__ a_ids, uint32_t row_lo, uint32_t n, uint32_t N,
                           Fr* __restrict__ out) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  stf(out + i, (i >= row_lo && i < row_lo + n) ? ldf(v + a_ids[i]) : Fr::zero());
}
// out[i] = c0 base^i
__global__ void k_pow_table(Fr* __restrict__ out, uint32_t n, Fr base, Fr c0) {
  constexpr uint32_t PER = 64;
  const uint64_t i0 = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * PER;
  if (i0 >= n) return;
  Fr cur = mul(pow_u64(base, i0), c0);
  const uint32_t end = (uint32_t)min((uint64_t)n, i0 + PER);
  for (uint32_t i = (uint32_t)i0; i < end; i++) {
    stf(out + i, cur);
    cur = mul(cur, base);
  }
}
// a[i] *= c0 s^i
__global__ void k_scale_pow(Fr* __restrict__ a, uint32_t n, Fr s, Fr c0) {
  constexpr uint32_t PER = 64;
  const uint64_t i0 = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * PER;
  if (i0 >= n) return;
  Fr cur = mul(pow_u64(s, i0), c0);
  const uint32_t end = (uint32_t)min((uint64_t)n, i0 + PER);
  for (uint32_t i = (uint32_t)i0; i < end; i++) {
    stf(a + i, mul(ldf(a + i), cur));
    cur = mul(cur, s);
  }
}
// S_col on H from the permutation: slot sigma = col' N + row' -> k_col' w^row'
__global__ void k_sigma_evals(const uint32_t* __restrict__ sigma_col, uint32_t N, int logN, const Fr* __restrict__ omega_pow, Fr k1, Fr k2,
                              Fr* __restrict__ out) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const uint32_t s = sigma_col[i], col = s >> logN, row = s & (N - 1u);
  Fr w = ldf(omega_pow + row);
  if (col == 1) w = mul(w, k1);
  else if (col == 2) w = mul(w, k2);
  stf(out + i, w);
}
// grand-product factors: num = prod_col (w + beta k_col w^i + gamma), den = prod_col (w + beta S_col + gamma)
__global__ void k_perm_terms(const Fr* __restrict__ a, const Fr* __restrict__ b, const Fr* __restrict__ c, const Fr* __restrict__ s1,
                             const Fr* __restrict__ s2, const Fr* __restrict__ s3, const Fr* __restrict__ omega_pow, uint32_t N, Fr beta,
                             Fr gamma, Fr k1, Fr k2, Fr* __restrict__ num, Fr* __restrict__ den) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const Fr va = ldf(a + i), vb = ldf(b + i), vc = ldf(c + i), bw = mul(beta, ldf(omega_pow + i));
  Fr n = add(add(va, bw), gamma);
  n = mul(n, add(add(vb, mul(bw, k1)), gamma));
  n = mul(n, add(add(vc, mul(bw, k2)), gamma));
  Fr d = add(add(va, mul(beta, ldf(s1 + i))), gamma);
  d = mul(d, add(add(vb, mul(beta, ldf(s2 + i))), gamma));
  d = mul(d, add(add(vc, mul(beta, ldf(s3 + i))), gamma));
  stf(num + i, n);
  stf(den + i, d);
}
// in-place inversion, 8 elements per thread with Montgomery's trick (zeros stay zero)
__global__ void k_batch_inv(Fr* __restrict__ a, uint32_t n) {
  constexpr int PER = 8;
  const uint64_t i0 = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * PER;
  if (i0 >= n) return;
  const int cnt = (int)min((uint64_t)PER, n - i0);
  Fr x[PER], pre[PER];
  Fr run = Fr::one();
#pragma unroll
  for (int j = 0; j < PER; j++) {
    x[j] = j < cnt ? ldf(a + i0 + j) : Fr::one();
    pre[j] = run;
    if (!x[j].is_zero()) run = mul(run, x[j]);
  }
  Fr r = inv(run);
#pragma unroll
  for (int j = PER - 1; j >= 0; j--) {
    if (x[j].is_zero()) continue;
    const Fr t = mul(r, pre[j]);
    r = mul(r, x[j]);
    if (j < cnt) stf(a + i0 + j, t);
  }
}
__global__ void k_mul_arrays(Fr* __restrict__ a, const Fr* __restrict__ b, uint32_t n) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) stf(a + i, mul(ldf(a + i), ldf(b + i)));
}

// ---- exclusive prefix product out[i] = prod_{j < i} a[j] (in place), three kernels ----------------------------------------
constexpr uint32_t SCAN_PER = 16, SCAN_THREADS = 256, SCAN_BLOCK = SCAN_PER * SCAN_THREADS;
__global__ void __launch_bounds__(SCAN_THREADS) k_pp_block_totals(const Fr* __restrict__ a, uint32_t n, Fr* __restrict__ totals) {
  __shared__ Fr sh[SCAN_THREADS];
  const uint64_t base = (uint64_t)blockIdx.x * SCAN_BLOCK + (uint64_t)threadIdx.x * SCAN_PER;
  Fr p = Fr::one();
  for (uint32_t j = 0; j < SCAN_PER; j++)
    if (base + j < n) p = mul(p, ldf(a + base + j));
  sh[threadIdx.x] = p;
  __syncthreads();
  for (uint32_t d = SCAN_THREADS / 2; d >= 1; d >>= 1) {
    if (threadIdx.x < d) sh[threadIdx.x] = mul(sh[threadIdx.x], sh[threadIdx.x + d]);
    __syncthreads();
  }
  if (threadIdx.x == 0) stf(totals + blockIdx.x, sh[0]);
}
// totals -> exclusive prefix products of the block totals (one CTA; nblocks <= 65536)
__global__ void __launch_bounds__(SCAN_THREADS) k_pp_scan_totals(Fr* __restrict__ totals, uint32_t nblocks) {
  __shared__ Fr sh[SCAN_THREADS];
  const uint32_t per = (nblocks + SCAN_THREADS - 1) / SCAN_THREADS;
  const uint32_t lo = threadIdx.x * per, hi = min(nblocks, lo + per);
  Fr p = Fr::one();
  for (uint32_t i = lo; i < hi; i++) p = mul(p, ldf(totals + i));
  sh[threadIdx.x] = p;
  __syncthreads();
  if (threadIdx.x == 0) {  // 256 sequential multiplications
    Fr run = Fr::one();
    for (uint32_t t = 0; t < SCAN_THREADS; t++) {
      const Fr x = sh[t];
      sh[t] = run;
      run = mul(run, x);
    }
  }
  __syncthreads();
  Fr run = sh[threadIdx.x];
  for (uint32_t i = lo; i < hi; i++) {
    const Fr x = ldf(totals + i);
    stf(totals + i, run);
    run = mul(run, x);
  }
}
__global__ void __launch_bounds__(SCAN_THREADS) k_pp_apply(Fr* __restrict__ a, uint32_t n, const Fr* __restrict__ totals) {
  __shared__ Fr sh[SCAN_THREADS];
  const uint64_t base = (uint64_t)blockIdx.x * SCAN_BLOCK + (uint64_t)threadIdx.x * SCAN_PER;
  Fr x[SCAN_PER];
  Fr p = Fr::one();
#pragma unroll
  for (uint32_t j = 0; j < SCAN_PER; j++) {
    x[j] = base + j < n ? ldf(a + base + j) : Fr::one();
    p = mul(p, x[j]);
  }
  sh[threadIdx.x] = p;
  __syncthreads();
  if (threadIdx.x == 0) {
    Fr run = ldf(totals + blockIdx.x);
    for (uint32_t t = 0; t < SCAN_THREADS; t++) {
      const Fr y = sh[t];
      sh[t] = run;
      run = mul(run, y);
    }
  }
  __syncthreads();
  Fr run = sh[threadIdx.x];
#pragma unroll
  for (uint32_t j = 0; j < SCAN_PER; j++) {
    if (base + j < n) stf(a + base + j, run);
    run = mul(run, x[j]);
  }
}

// ---- quotient on one coset ------------------------------------------------------------------------------------------------
struct CosetEvals {
  const Fr *a, *b, *c, *z, *p2, *pi, *ql, *qr, *qm, *qo, *qc, *qcp, *s1, *s2, *s3, *l0;
};
__global__ void k_quotient(CosetEvals e, uint32_t N, const Fr* __restrict__ omega_pow, Fr shift, Fr beta, Fr gamma, Fr alpha, Fr k1, Fr k2,
                           Fr zh_inv, Fr* __restrict__ t) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const Fr a = ldf(e.a + i), b = ldf(e.b + i), c = ldf(e.c + i), z = ldf(e.z + i), zn = ldf(e.z + ((i + 1u) & (N - 1u)));
  const Fr x = mul(shift, ldf(omega_pow + i));
  Fr gate = mul(ldf(e.ql + i), a);
  gate = add(gate, mul(ldf(e.qr + i), b));
  gate = add(gate, mul(ldf(e.qm + i), mul(a, b)));
  gate = add(gate, mul(ldf(e.qo + i), c));
  gate = add(gate, ldf(e.qc + i));
  gate = add(gate, ldf(e.pi + i));
  gate = add(gate, mul(ldf(e.qcp + i), ldf(e.p2 + i)));
  const Fr bx = mul(beta, x);
  Fr p1 = mul(z, add(add(a, bx), gamma));
  p1 = mul(p1, add(add(b, mul(bx, k1)), gamma));
  p1 = mul(p1, add(add(c, mul(bx, k2)), gamma));
  Fr p2 = mul(zn, add(add(a, mul(beta, ldf(e.s1 + i))), gamma));
  p2 = mul(p2, add(add(b, mul(beta, ldf(e.s2 + i))), gamma));
  p2 = mul(p2, add(add(c, mul(beta, ldf(e.s3 + i))), gamma));
  const Fr bound = mul(sub(z, Fr::one()), ldf(e.l0 + i));
  Fr r = add(gate, mul(alpha, add(sub(p1, p2), mul(alpha, bound))));
  stf(t + i, mul(r, zh_inv));
}
// d_j[r] = sum_k (c[k N + r] g^(N k)) i4^(j k)  ->  c[k N + r], in place (d0..d3 become t0..t3)
__global__ void k_combine_t(Fr* __restrict__ d0, Fr* __restrict__ d1, Fr* __restrict__ d2, Fr* __restrict__ d3, uint32_t N, Fr i4inv,
                            Fr quarter, Fr gNinv) {
  const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= N) return;
  const Fr x0 = ldf(d0 + r), x1 = ldf(d1 + r), x2 = ldf(d2 + r), x3 = ldf(d3 + r);
  // e_k = (1/4) sum_j i4^(-j k) x_j
  const Fr w1 = i4inv, w2 = mul(i4inv, i4inv), w3 = mul(w2, i4inv);
  Fr e0 = add(add(x0, x1), add(x2, x3));
  Fr e1 = add(add(x0, mul(w1, x1)), add(mul(w2, x2), mul(w3, x3)));
  Fr e2 = add(add(x0, mul(w2, x1)), add(x2, mul(w2, x3)));  // w2^2 = 1, w2^3 = w2
  Fr e3 = add(add(x0, mul(w3, x1)), add(mul(w2, x2), mul(w1, x3)));  // w3^2 = w2, w3^3 = w1
  const Fr g1 = gNinv, g2 = mul(g1, g1), g3 = mul(g2, g1);
  stf(d0 + r, mul(e0, quarter));
  stf(d1 + r, mul(mul(e1, quarter), g1));
  stf(d2 + r, mul(mul(e2, quarter), g2));
  stf(d3 + r, mul(mul(e3, quarter), g3));
}

// ---- evaluation of a coefficient-form polynomial at one point: per-block partial sums, then a one-block finish -----------
__global__ void __launch_bounds__(256) k_eval_partial(const Fr* __restrict__ coef, uint32_t n, Fr x, Fr* __restrict__ partial) {
  __shared__ Fr sh[256];
  constexpr uint32_t PER = 32;
  const uint64_t i0 = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * PER;
  Fr acc = Fr::zero();
  if (i0 < n) {
    const uint32_t end = (uint32_t)min((uint64_t)n, i0 + PER);
    for (uint32_t i = end; i-- > (uint32_t)i0;) acc = add(mul(acc, x), ldf(coef + i));  // Horner inside the chunk
    acc = mul(acc, pow_u64(x, i0));
  }
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (uint32_t d = 128; d >= 1; d >>= 1) {
    if (threadIdx.x < d) sh[threadIdx.x] = add(sh[threadIdx.x], sh[threadIdx.x + d]);
    __syncthreads();
  }
  if (threadIdx.x == 0) stf(partial + blockIdx.x, sh[0]);
}
__global__ void __launch_bounds__(256) k_sum_partials(const Fr* __restrict__ partial, uint32_t n, Fr* __restrict__ out) {
  __shared__ Fr sh[256];
  Fr acc = Fr::zero();
  for (uint32_t i = threadIdx.x; i < n; i += 256) acc = add(acc, ldf(partial + i));
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (uint32_t d = 128; d >= 1; d >>= 1) {
    if (threadIdx.x < d) sh[threadIdx.x] = add(sh[threadIdx.x], sh[threadIdx.x + d]);
    __syncthreads();
  }
  if (threadIdx.x == 0) stf(out, sh[0]);
}
// W[i] += nu_k (f[i] - y)
__global__ void k_open_accumulate(Fr* __restrict__ W, const Fr* __restrict__ f, uint32_t N, Fr y, Fr nu_k) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N) stf(W + i, add(ldf(W + i), mul(nu_k, sub(ldf(f + i), y))));
}
// den[i] = w^i - point
__global__ void k_open_denominators(const Fr* __restrict__ omega_pow, uint32_t N, Fr point, Fr* __restrict__ den) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N) stf(den + i, sub(ldf(omega_pow + i), point));
}
__global__ void k_fill(Fr* __restrict__ a, uint32_t n, Fr v) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) stf(a + i, v);
}
__global__ void k_set_public(Fr* __restrict__ pi, const Fr* __restrict__ v, const uint32_t* __restrict__ public_var, uint32_t n) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) stf(pi + i, neg(ldf(v + public_var[i])));
}
// out[i] = [s_i] G for the SRS (s_i = tau^i, Montgomery), from the table of d 2^(16 w) G
__global__ void __launch_bounds__(128) k_srs_points(const G1Affine* __restrict__ table, const Fr* __restrict__ scalars, size_t n,
                                                    G1Affine* __restrict__ out) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const Fr s = from_mont(scalars[i]);
  G1XYZZ acc = G1XYZZ::inf();
#pragma unroll 1
  for (int w = 0; w < 16; w++) {
    const uint32_t d = (s.l[w >> 1] >> (16 * (w & 1))) & 0xffffu;
    if (d) add_mixed(acc, table[(size_t)w * 65536 + d], false);
  }
  out[i] = to_affine(acc);
}
__global__ void __launch_bounds__(128) k_gen_table_row(G1Affine g, G1Affine* __restrict__ out) {
  // out[i] = [i] g, i < 65536; 32 consecutive multiples per thread
  const size_t i0 = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 32;
  if (i0 >= 65536) return;
  G1XYZZ cur = G1XYZZ::inf();
  for (int b = 16; b >= 0; b--) {
    cur = dbl(cur);
    if ((i0 >> b) & 1ull) add_mixed(cur, g, false);
  }
  for (int j = 0; j < 32; j++) {
    out[i0 + j] = to_affine(cur);
    add_mixed(cur, g, false);
  }
}

}  // namespace plonk
}  // namespace gpw

using namespace gpw;
using namespace gpw::plonk;

// ---- key ---------------------------------------------------------------------------------------------------------------------
struct gpw_plonk_key {
  gpw_ctx* ctx = nullptr;
  gpw_circuit* circ = nullptr;
  int logN = 0;
  uint32_t N = 0, n_gates = 0, n_vars = 0, n_orig = 0, n_public_rows = 0, n_qcp_rows = 0, n_inputs = 0, n_pub = 0;
  bool has_commit = false;
  // device
  std::vector<void*> allocs;
  uint32_t *a_ids = nullptr, *b_ids = nullptr, *c_ids = nullptr, *sigma = nullptr, *public_var = nullptr, *chain_wire = nullptr,
           *chain_coeff = nullptr;
  scs::Chain* chains = nullptr;
  std::vector<uint32_t> level_off;
  Fr *coeffs = nullptr, *omega_pow = nullptr;
  Fr* sel_c[6] = {};   // qL qR qM qO qC Qcp, coefficient form
  Fr* sel_e[6] = {};   // the same on H (kept: the opening proof is built in evaluation form)
  Fr* sig_c[3] = {};   // S1 S2 S3, coefficient form
  Fr* sig_e[3] = {};
  Fr* l0_c = nullptr;  // L_0 = (1/N) sum X^k
  // the ten proof-independent polynomials (selectors, sigmas, L_0) evaluated on the four quotient cosets at setup:
  // fixed_ce[j * 10 + p], 40 N-sized arrays (43 GB at 2^25) that save 40 of the 68 transforms of every proof
  Fr* fixed_ce[40] = {};
  G1Affine* srs = nullptr;  // [tau^i] G1, i < N
  // the same SRS in the Lagrange basis, [L_i(tau)] G1: polynomials known by their values on H (the wire columns, P2) are
  // committed from those values directly - witness values are mostly small (bits, 16-bit limbs, 64-bit field elements), so the
  // MSM skips most digits, whereas their coefficient forms are full-width. (Generated from tau like the monomial SRS; for a
  // ceremony SRS it would come from an inverse FFT "in the exponent".)
  G1Affine* srs_lagrange = nullptr;
  // fixed-base table of the monomial SRS, 2^(22 w) [tau^i] G1 for the 12 windows of a scalar (25.8 GB at 2^25): the six
  // commitments to full-width coefficient vectors of every proof (Z, t0..t2, the two opening polynomials) then need 12 instead
  // of 16 bucket additions per scalar into one bucket set (the Z MSM of the Groth16 path does the same). Built when the
  // device has the room (GPW_PLONK_FIXED=0 switches it off); nullptr: plain windowed MSM.
  G1Affine* srs_t = nullptr;
  G2Affine tau2;            // [tau] G2
  G1Affine vk_com[9];       // [qL] [qR] [qM] [qO] [qC] [Qcp] [S1] [S2] [S3]
  Fr k1, k2, omega;
  uint8_t vk_digest[32];
  // per-proof buffers
  Fr* v = nullptr;      // n_vars variables
  uint64_t* inputs_dev = nullptr;
  Fr* buf[24] = {};     // N-sized work arrays
  float t_ms[8] = {};
};

namespace {

int dalloc(gpw_plonk_key* k, void** p, size_t bytes) {
  cudaError_t e = cudaMalloc(p, bytes ? bytes : 16);
  if (e != cudaSuccess) {
    set_error("plonk: cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
    return GPW_ENOMEM;
  }
  k->allocs.push_back(*p);
  return GPW_OK;
}
template <class T>
int upload(gpw_plonk_key* k, const std::vector<T>& h, T** d) {
  GPW_TRY(dalloc(k, (void**)d, h.size() * sizeof(T)));
  if (!h.empty()) GPW_CUDA(cudaMemcpy(*d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
  return GPW_OK;
}

Fr fr_u64(uint64_t v) { return fr_from_u64_host(v); }

Fr root_of_unity(int logn) {  // w_28 = 5^((r-1)/2^28), squared down
  Fr m = modulus<FrParams>();
  uint32_t e[8];
  for (int i = 0; i < 8; i++) e[i] = m.l[i];
  e[0] -= 1;
  uint32_t s[8];
  for (int i = 0; i < 8; i++) {
    uint64_t v = e[i];
    if (i + 1 < 8) v |= (uint64_t)e[i + 1] << 32;
    s[i] = (uint32_t)(v >> 28);
  }
  Fr w = pow_words(fr_u64(5), s, 8);
  for (int i = 0; i < 28 - logn; i++) w = sqr(w);
  return w;
}

void fr_be(const Fr& mont, uint8_t out[32]) {
  Fr c = from_mont(mont);
  for (int i = 0; i < 8; i++)
    for (int b = 0; b < 4; b++) out[31 - (4 * i + b)] = (uint8_t)(c.l[i] >> (8 * b));
}
void g1_be(const G1Affine& p, uint8_t out[64]) {
  Fp x = from_mont(p.x), y = from_mont(p.y);
  for (int i = 0; i < 8; i++)
    for (int b = 0; b < 4; b++) {
      out[31 - (4 * i + b)] = (uint8_t)(x.l[i] >> (8 * b));
      out[63 - (4 * i + b)] = (uint8_t)(y.l[i] >> (8 * b));
    }
  if (p.is_inf()) out[0] = 0x40;
}
struct Transcript {
  std::vector<uint8_t> m;
  void fr(const Fr& x) {
    uint8_t b[32];
    fr_be(x, b);
    m.insert(m.end(), b, b + 32);
  }
  void g1(const G1Affine& p) {
    uint8_t b[64];
    g1_be(p, b);
    m.insert(m.end(), b, b + 64);
  }
  void bytes(const uint8_t* p, size_t n) { m.insert(m.end(), p, p + n); }
  Fr challenge(const char* label) {
    uint64_t out[4];
    hash_to_fr(m.data(), m.size(), label, out);
    m.clear();
    return fe::fr_from_limbs(out);
  }
};

constexpr int PLONK_FIXED_C = 22, PLONK_FIXED_W = (254 + PLONK_FIXED_C) / PLONK_FIXED_C;
int commit(gpw_plonk_key* k, const Fr* coef, G1Affine* out) {
  if (k->srs_t)
    return gpw_msm_g1_fixed_dev(k->ctx, (uint64_t)coef, (uint64_t)k->srs_t, k->N, 1, PLONK_FIXED_C, PLONK_FIXED_W, (uint64_t*)out);
  return gpw_msm_g1_dev(k->ctx, (uint64_t)coef, (uint64_t)k->srs, k->N, 1, 0, 0, 0, (uint64_t*)out);
}
// the same commitment from the polynomial's values on H (Lagrange-basis SRS)
int commit_evals(gpw_plonk_key* k, const Fr* evals, G1Affine* out) {
  return gpw_msm_g1_dev(k->ctx, (uint64_t)evals, (uint64_t)k->srs_lagrange, k->N, 1, 0, 0, 0, (uint64_t*)out);
}
int intt(gpw_plonk_key* k, Fr* a) { return gpw_ntt_fr_dev(k->ctx, (uint64_t)a, k->logN, 1, 0, 0, 0); }
int ntt(gpw_plonk_key* k, Fr* a) { return gpw_ntt_fr_dev(k->ctx, (uint64_t)a, k->logN, 0, 0, 0, 0); }

int eval_at(gpw_plonk_key* k, const Fr* coef, const Fr& x, Fr* partial, Fr* out_host) {
  const uint32_t nb = div_up(div_up(k->N, 32), 256);
  k_eval_partial<<<nb, 256, 0, k->ctx->stream>>>(coef, k->N, x, partial);
  GPW_CHECK_LAUNCH();
  k_sum_partials<<<1, 256, 0, k->ctx->stream>>>(partial, nb, partial + nb);
  GPW_CHECK_LAUNCH();
  k->ctx->launches += 2;
  GPW_CUDA(cudaMemcpyAsync(out_host, partial + nb, sizeof(Fr), cudaMemcpyDeviceToHost, k->ctx->stream));
  GPW_CUDA(cudaStreamSynchronize(k->ctx->stream));
  return GPW_OK;
}

int prefix_product(gpw_plonk_key* k, Fr* a, Fr* totals) {
  const uint32_t nb = div_up(k->N, SCAN_BLOCK);
  cudaStream_t st = k->ctx->stream;
  k_pp_block_totals<<<nb, SCAN_THREADS, 0, st>>>(a, k->N, totals);
  GPW_CHECK_LAUNCH();
  k_pp_scan_totals<<<1, SCAN_THREADS, 0, st>>>(totals, nb);
  GPW_CHECK_LAUNCH();
  k_pp_apply<<<nb, SCAN_THREADS, 0, st>>>(a, k->N, totals);
  GPW_CHECK_LAUNCH();
  k->ctx->launches += 3;
  return GPW_OK;
}

}  // namespace

extern "C" void gpw_plonk_key_free(gpw_plonk_key* k) {
  if (!k) return;
  cudaSetDevice(k->ctx->device);
  for (void* p : k->allocs) cudaFree(p);
  delete k;
}

// plonk.Setup(ccs, srs) with srs = test.NewKZGSRS(ccs) (benchmark.go:105, 130): the SRS is generated here from a seed (tau from
// the seed or the OS), as the reference's benchmark does with gnark's test SRS - a production deployment would load a ceremony
// SRS into the same array.
extern "C" int gpw_plonk_setup(gpw_ctx* ctx, gpw_circuit* circ, const uint8_t* seed32, gpw_plonk_key** out) {
  if (!ctx || !circ || !out) {
    set_error("plonk_setup: null argument");
    return GPW_EINVAL;
  }
  GPW_CUDA(cudaSetDevice(ctx->device));
  const fe::API* api = gpw_circuit_api_internal(circ);
  scs::System sys;
  std::vector<uint32_t> sigma;
  try {
    sys = scs::Build(*api);
    scs::BuildPermutation(sys, &sigma);
  } catch (const std::exception& e) {
    set_error("plonk_setup: %s", e.what());
    return GPW_EINVAL;
  }
  gpw_plonk_key* k = new gpw_plonk_key();
  k->ctx = ctx;
  k->circ = circ;
  k->logN = sys.logN;
  k->N = 1u << sys.logN;
  k->n_gates = sys.n_gates;
  k->n_vars = sys.n_vars;
  k->n_orig = sys.n_orig;
  k->n_public_rows = sys.n_public_rows;
  k->n_qcp_rows = sys.n_qcp_rows;
  k->has_commit = sys.has_commit;
  k->n_pub = api->NumPublic();
  k->n_inputs = api->NumPublic() + api->NumSecret();
  k->level_off = sys.level_off;
  const uint32_t N = k->N;
  cudaStream_t st = ctx->stream;
  int rc = 0;
  auto fail = [&](int code) {
    gpw_plonk_key_free(k);
    return code;
  };
  std::vector<uint8_t> qcp_flags(sys.qcp);
  uint8_t* qcp_dev = nullptr;
  uint32_t* qid[5] = {};
  if ((rc = upload(k, sys.a, &k->a_ids)) || (rc = upload(k, sys.b, &k->b_ids)) || (rc = upload(k, sys.c, &k->c_ids)) ||
      (rc = upload(k, sigma, &k->sigma)) || (rc = upload(k, sys.public_var, &k->public_var)) || (rc = upload(k, sys.chain_wire, &k->chain_wire)) ||
      (rc = upload(k, sys.chain_coeff, &k->chain_coeff)) || (rc = upload(k, sys.chains, &k->chains)) || (rc = upload(k, sys.coeffs, &k->coeffs)) ||
      (rc = upload(k, qcp_flags, &qcp_dev)) || (rc = upload(k, sys.ql, &qid[0])) || (rc = upload(k, sys.qr, &qid[1])) ||
      (rc = upload(k, sys.qm, &qid[2])) || (rc = upload(k, sys.qo, &qid[3])) || (rc = upload(k, sys.qc, &qid[4])))
    return fail(rc);
  std::vector<uint32_t>().swap(sigma);
  if ((rc = dalloc(k, (void**)&k->omega_pow, (size_t)N * sizeof(Fr))) || (rc = dalloc(k, (void**)&k->l0_c, (size_t)N * sizeof(Fr))) ||
      (rc = dalloc(k, (void**)&k->srs, (size_t)N * sizeof(G1Affine))) || (rc = dalloc(k, (void**)&k->srs_lagrange, (size_t)N * sizeof(G1Affine))) || (rc = dalloc(k, (void**)&k->v, (size_t)k->n_vars * sizeof(Fr))) ||
      (rc = dalloc(k, (void**)&k->inputs_dev, (size_t)(k->n_inputs ? k->n_inputs : 1) * 32)))
    return fail(rc);
  for (int i = 0; i < 6; i++)
    if ((rc = dalloc(k, (void**)&k->sel_c[i], (size_t)N * sizeof(Fr))) || (rc = dalloc(k, (void**)&k->sel_e[i], (size_t)N * sizeof(Fr)))) return fail(rc);
  for (int i = 0; i < 3; i++)
    if ((rc = dalloc(k, (void**)&k->sig_c[i], (size_t)N * sizeof(Fr))) || (rc = dalloc(k, (void**)&k->sig_e[i], (size_t)N * sizeof(Fr)))) return fail(rc);
  for (auto& b : k->buf)
    if ((rc = dalloc(k, (void**)&b, (size_t)N * sizeof(Fr)))) return fail(rc);
  k->omega = root_of_unity(k->logN);
  k->k1 = fr_u64(5);
  k->k2 = fr_u64(25);
  const int G = div_up(N, 256), GP = div_up(div_up(N, 64), 128);
  k_pow_table<<<GP, 128, 0, st>>>(k->omega_pow, N, k->omega, Fr::one());
  k_fill<<<G, 256, 0, st>>>(k->l0_c, N, inv(fr_u64(N)));
  // selectors and permutation polynomials: evaluations on H -> coefficients
  for (int i = 0; i < 5; i++) k_coeff_col<<<G, 256, 0, st>>>(k->coeffs, qid[i], k->n_gates, N, k->sel_e[i]);
  k_flag_col<<<G, 256, 0, st>>>(qcp_dev, k->n_gates, N, k->sel_e[5]);
  for (int c = 0; c < 3; c++) k_sigma_evals<<<G, 256, 0, st>>>(k->sigma + (size_t)c * N, N, k->logN, k->omega_pow, k->k1, k->k2, k->sig_e[c]);
  if (cudaGetLastError() != cudaSuccess) {
    set_error("plonk_setup: kernel launch failed");
    return fail(GPW_ECUDA);
  }
  for (int i = 0; i < 6; i++) {
    if (cudaMemcpyAsync(k->sel_c[i], k->sel_e[i], (size_t)N * sizeof(Fr), cudaMemcpyDeviceToDevice, st) != cudaSuccess) return fail(GPW_ECUDA);
    if ((rc = intt(k, k->sel_c[i]))) return fail(rc);
  }
  for (int i = 0; i < 3; i++) {
    if (cudaMemcpyAsync(k->sig_c[i], k->sig_e[i], (size_t)N * sizeof(Fr), cudaMemcpyDeviceToDevice, st) != cudaSuccess) return fail(GPW_ECUDA);
    if ((rc = intt(k, k->sig_c[i]))) return fail(rc);
  }
  {  // coset evaluations of the fixed polynomials
    const Fr g = fr_u64(5), rho = root_of_unity(k->logN + 2);
    const Fr* fixed_c[10] = {k->sel_c[0], k->sel_c[1], k->sel_c[2], k->sel_c[3], k->sel_c[4], k->sel_c[5], k->sig_c[0], k->sig_c[1], k->sig_c[2], k->l0_c};
    for (int j = 0; j < 4; j++) {
      const Fr shift = mul(g, pow_u64(rho, (uint64_t)j));
      for (int p = 0; p < 10; p++) {
        Fr*& dst = k->fixed_ce[j * 10 + p];
        if ((rc = dalloc(k, (void**)&dst, (size_t)N * sizeof(Fr)))) return fail(rc);
        if (cudaMemcpyAsync(dst, fixed_c[p], (size_t)N * sizeof(Fr), cudaMemcpyDeviceToDevice, st) != cudaSuccess) return fail(GPW_ECUDA);
        k_scale_pow<<<GP, 128, 0, st>>>(dst, N, shift, Fr::one());
        if ((rc = ntt(k, dst))) return fail(rc);
      }
    }
    ctx->launches += 40;
  }
  // SRS: tau^i G1 (i < N), tau G2
  uint8_t seed[64];
  memset(seed, 0, sizeof(seed));
  if (seed32) memcpy(seed, seed32, 32);
  else if (getrandom(seed, 32, 0) != 32) {
    set_error("plonk_setup: getrandom failed");
    return fail(GPW_EINVAL);
  }
  memcpy(seed + 32, "kzg-tau", 7);
  uint64_t tl[4];
  hash_to_fr(seed, 64, "gpw-plonk-setup", tl);
  const Fr tau = fe::fr_from_limbs(tl);
  memset(seed, 0, sizeof(seed));
  {
    G1Affine* table = nullptr;
    if ((rc = dalloc(k, (void**)&table, (size_t)16 * 65536 * sizeof(G1Affine)))) return fail(rc);
    G1XYZZ g = G1XYZZ::from_affine(generator<Fp>());
    for (int w = 0; w < 16; w++) {
      k_gen_table_row<<<div_up(65536 / 32, 128), 128, 0, st>>>(to_affine(g), table + (size_t)w * 65536);
      for (int i = 0; i < 16; i++) g = dbl(g);
    }
    Fr* pw = k->buf[0];
    k_pow_table<<<GP, 128, 0, st>>>(pw, N, tau, Fr::one());
    k_srs_points<<<div_up(N, 128), 128, 0, st>>>(table, pw, N, k->srs);
    {  // L_i(tau) = w^i (tau^N - 1) / (N (tau - w^i))
      Fr tN = tau;
      for (int i = 0; i < k->logN; i++) tN = sqr(tN);
      const Fr c = mul(sub(tN, Fr::one()), inv(fr_u64(N)));
      Fr* den = k->buf[1];
      k_open_denominators<<<G, 256, 0, st>>>(k->omega_pow, N, tau, den);  // w^i - tau
      k_batch_inv<<<div_up(div_up(N, 8), 128), 128, 0, st>>>(den, N);
      k_mul_arrays<<<G, 256, 0, st>>>(den, k->omega_pow, N);
      k_scale_pow<<<GP, 128, 0, st>>>(den, N, Fr::one(), neg(c));          // * -(tau^N - 1) / N
      k_srs_points<<<div_up(N, 128), 128, 0, st>>>(table, den, N, k->srs_lagrange);
    }
    if (cudaStreamSynchronize(st) != cudaSuccess) {
      set_error("plonk_setup: SRS generation failed: %s", cudaGetErrorString(cudaGetLastError()));
      return fail(GPW_ECUDA);
    }
    cudaFree(table);
    k->allocs.pop_back();
    Fr tc = from_mont(tau);
    k->tau2 = to_affine(host_scalar_mul(generator<Fp2>(), tc.l));
  }
  {  // fixed-base table of the monomial SRS, if the device has the room for it next to the proofs' scratch
    static const bool want = !getenv("GPW_PLONK_FIXED") || atoi(getenv("GPW_PLONK_FIXED")) != 0;
    const size_t need = (size_t)N * PLONK_FIXED_W * sizeof(G1Affine);
    size_t free_b = 0, total_b = 0;
    // (small domains: the reduction of 2^21 buckets would cost more than the four additions per scalar it saves)
    if (want && N >= (1u << 22) && (uint64_t)N * PLONK_FIXED_W < (1ull << 31) && cudaMemGetInfo(&free_b, &total_b) == cudaSuccess &&
        free_b > need + ((size_t)24 << 30)) {
      if ((rc = dalloc(k, (void**)&k->srs_t, need))) return fail(rc);
      if ((rc = gpw_msm_g1_fixed_table(ctx, (uint64_t)k->srs, N, PLONK_FIXED_C, PLONK_FIXED_W, (uint64_t)k->srs_t))) return fail(rc);
      if (cudaStreamSynchronize(st) != cudaSuccess) {
        set_error("plonk_setup: fixed-base table failed: %s", cudaGetErrorString(cudaGetLastError()));
        return fail(GPW_ECUDA);
      }
    }
  }
  // verifying key: commitments to the selectors and the permutation polynomials
  for (int i = 0; i < 6; i++)
    if ((rc = commit(k, k->sel_c[i], &k->vk_com[i]))) return fail(rc);
  for (int i = 0; i < 3; i++)
    if ((rc = commit(k, k->sig_c[i], &k->vk_com[6 + i]))) return fail(rc);
  {
    std::vector<uint8_t> m;
    for (int i = 0; i < 9; i++) {
      uint8_t b[64];
      g1_be(k->vk_com[i], b);
      m.insert(m.end(), b, b + 64);
    }
    for (int i = 0; i < 4; i++) m.push_back((uint8_t)(N >> (8 * i)));
    for (int i = 0; i < 4; i++) m.push_back((uint8_t)(k->n_public_rows >> (8 * i)));
    sha256_bytes(m.data(), m.size(), k->vk_digest);
  }
  ctx->launches += 16 + 9 + 16 + 2;
  *out = k;
  return GPW_OK;
}

// info8: {logN, gates, variables, public rows (ONE + public inputs + challenge), Qcp rows, inputs, has_commit, chain levels}
extern "C" int gpw_plonk_key_info(const gpw_plonk_key* k, uint64_t* info8) {
  if (!k || !info8) return GPW_EINVAL;
  uint64_t v[8] = {(uint64_t)k->logN, k->n_gates, k->n_vars, k->n_public_rows, k->n_qcp_rows, k->n_inputs, k->has_commit ? 1u : 0u,
                   (uint64_t)k->level_off.size() - 1};
  memcpy(info8, v, sizeof(v));
  return GPW_OK;
}

// vk bytes (own layout, documented in oracle/plonk_verify.py): u32 logN | u32 n_public_rows | u32 has_commit | k1 | k2 | omega (32 B BE
// each) | 9 commitments (64 B raw each: qL qR qM qO qC Qcp S1 S2 S3) | [tau] G2 (128 B raw: X.A1 X.A0 Y.A1 Y.A0)
extern "C" int gpw_plonk_vk_write(const gpw_plonk_key* k, uint8_t* out, size_t cap, size_t* len) {
  if (!k) return GPW_EINVAL;
  std::vector<uint8_t> b;
  auto u32 = [&](uint32_t v) {
    for (int i = 3; i >= 0; i--) b.push_back((uint8_t)(v >> (8 * i)));
  };
  u32((uint32_t)k->logN);
  u32(k->n_public_rows);
  u32(k->has_commit ? 1 : 0);
  uint8_t t[128];
  for (const Fr* f : {&k->k1, &k->k2, &k->omega}) {
    fr_be(*f, t);
    b.insert(b.end(), t, t + 32);
  }
  for (int i = 0; i < 9; i++) {
    g1_be(k->vk_com[i], t);
    b.insert(b.end(), t, t + 64);
  }
  {
    const Fp* cs[4] = {&k->tau2.x.c1, &k->tau2.x.c0, &k->tau2.y.c1, &k->tau2.y.c0};
    for (int j = 0; j < 4; j++) {
      Fp c = from_mont(*cs[j]);
      for (int i = 0; i < 8; i++)
        for (int bb = 0; bb < 4; bb++) t[32 * j + 31 - (4 * i + bb)] = (uint8_t)(c.l[i] >> (8 * bb));
    }
    b.insert(b.end(), t, t + 128);
  }
  if (len) *len = b.size();
  if (!out) return GPW_OK;
  if (cap < b.size()) {
    set_error("plonk_vk_write: buffer too small");
    return GPW_EINVAL;
  }
  memcpy(out, b.data(), b.size());
  return GPW_OK;
}

constexpr size_t PLONK_PROOF_BYTES = 10 * 64 + 18 * 32;

// plonk.Prove (benchmark.go:162). inputs: n_inputs x 4 u64 canonical (host), gpw_circuit_parse_inputs order. out_proof
// (PLONK_PROOF_BYTES = 1216): 10 commitments, 64 B raw big-endian each: [a] [b] [c] [P2] [Z] [t0] [t1] [t2] [W_zeta] [W_zeta_w];
// then 18 evaluations, 32 B big-endian each: a b c z p2 qL qR qM qO qC Qcp S1 S2 S3 t0 t1 t2 at zeta, z at zeta w.
extern "C" int gpw_plonk_prove(gpw_plonk_key* k, const uint64_t* inputs, uint8_t* out_proof, size_t cap) {
  if (!k || !inputs || !out_proof || cap < PLONK_PROOF_BYTES) {
    set_error("plonk_prove: bad argument (proof needs %zu bytes)", PLONK_PROOF_BYTES);
    return GPW_EINVAL;
  }
  gpw_ctx* ctx = k->ctx;
  GPW_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  const uint32_t N = k->N;
  const int G = div_up(N, 256), GP = div_up(div_up(N, 64), 128);
  cudaEvent_t ev[7];
  for (auto& e : ev) GPW_CUDA(cudaEventCreate(&e));
  auto mark = [&](int i) { return cudaEventRecord(ev[i], st); };
  // buffers: evaluations on H (kept for the opening) and coefficient forms
  Fr *a_e = k->buf[0], *b_e = k->buf[1], *c_e = k->buf[2], *z_e = k->buf[3], *p2_e = k->buf[4];
  Fr *a_c = k->buf[5], *b_c = k->buf[6], *c_c = k->buf[7], *z_c = k->buf[8], *p2_c = k->buf[9], *pi_c = k->buf[10];
  Fr* d[4] = {k->buf[11], k->buf[12], k->buf[13], k->buf[14]};  // coset quotient pieces -> t0..t3
  Fr *tmpA = k->buf[15], *tmpB = k->buf[16];
  GPW_CUDA(mark(0));
  // ---- round 0: witness -------------------------------------------------------------------------------------------------
  GPW_CUDA(cudaMemcpyAsync(k->inputs_dev, inputs, (size_t)k->n_inputs * 32, cudaMemcpyHostToDevice, st));
  GPW_CUDA(cudaMemsetAsync(k->v, 0, (size_t)k->n_vars * sizeof(Fr), st));
  GPW_TRY(gpw_witness_solve_phase1_on(k->circ, ctx, (uint64_t)k->inputs_dev, 1, (uint64_t)k->v, k->n_orig));
  G1Affine com[10];
  for (auto& c : com) c = G1Affine{Fp::zero(), Fp::zero()};
  uint64_t X[4] = {0, 0, 0, 0};
  k_p2_evals<<<G, 256, 0, st>>>(k->v, k->a_ids, k->n_public_rows, k->n_qcp_rows, N, p2_e);
  GPW_CHECK_LAUNCH();
  GPW_CUDA(cudaMemcpyAsync(p2_c, p2_e, (size_t)N * sizeof(Fr), cudaMemcpyDeviceToDevice, st));
  GPW_TRY(intt(k, p2_c));
  if (k->has_commit) {
    GPW_TRY(commit_evals(k, p2_e, &com[3]));
    uint8_t ser[64];
    g1_be(com[3], ser);
    hash_to_fr(ser, 64, "bsb22-commitment", X);
  }
  GPW_TRY(gpw_witness_solve_phase2_on(k->circ, ctx, X, 1, (uint64_t)k->v, k->n_orig));
  for (size_t l = 0; l + 1 < k->level_off.size(); l++) {
    const uint32_t lo = k->level_off[l], n = k->level_off[l + 1] - lo;
    if (!n) continue;
    k_scs_chains<<<div_up(n, 128), 128, 0, st>>>(k->chains + lo, n, k->chain_wire, k->chain_coeff, k->coeffs, k->v);
    GPW_CHECK_LAUNCH();
    ctx->launches++;
  }
  GPW_CUDA(mark(1));
  // ---- round 1: wire columns ---------------------------------------------------------------------------------------------
  k_gather_col<<<G, 256, 0, st>>>(k->v, k->a_ids, k->n_gates, N, a_e);
  k_gather_col<<<G, 256, 0, st>>>(k->v, k->b_ids, k->n_gates, N, b_e);
  k_gather_col<<<G, 256, 0, st>>>(k->v, k->c_ids, k->n_gates, N, c_e);
  GPW_CHECK_LAUNCH();
  Fr* ev_c[3][2] = {{a_e, a_c}, {b_e, b_c}, {c_e, c_c}};
  for (int i = 0; i < 3; i++) {
    GPW_CUDA(cudaMemcpyAsync(ev_c[i][1], ev_c[i][0], (size_t)N * sizeof(Fr), cudaMemcpyDeviceToDevice, st));
    GPW_TRY(intt(k, ev_c[i][1]));
    GPW_TRY(commit_evals(k, ev_c[i][0], &com[i]));
  }
  // public-input polynomial: -x_i on the public rows
  GPW_CUDA(cudaMemsetAsync(pi_c, 0, (size_t)N * sizeof(Fr), st));
  k_set_public<<<div_up(k->n_public_rows, 128), 128, 0, st>>>(pi_c, k->v, k->public_var, k->n_public_rows);
  GPW_CHECK_LAUNCH();
  std::vector<Fr> xs(k->n_public_rows);
  {
    std::vector<uint32_t> pv(k->n_public_rows);
    GPW_CUDA(cudaMemcpy(pv.data(), k->public_var, pv.size() * 4, cudaMemcpyDeviceToHost));
    for (uint32_t i = 0; i < k->n_public_rows; i++) GPW_CUDA(cudaMemcpy(&xs[i], k->v + pv[i], sizeof(Fr), cudaMemcpyDeviceToHost));
  }
  GPW_TRY(intt(k, pi_c));
  GPW_CUDA(mark(2));
  // ---- round 2: permutation grand product ---------------------------------------------------------------------------------
  Transcript tr;
  tr.bytes(k->vk_digest, 32);
  for (const Fr& x : xs) tr.fr(x);
  for (int i = 0; i < 4; i++) tr.g1(com[i]);
  const Fr beta = tr.challenge("gpw-plonk-beta");
  tr.fr(beta);
  const Fr gamma = tr.challenge("gpw-plonk-gamma");
  k_perm_terms<<<G, 256, 0, st>>>(a_e, b_e, c_e, k->sig_e[0], k->sig_e[1], k->sig_e[2], k->omega_pow, N, beta, gamma, k->k1, k->k2, z_e, tmpA);
  GPW_CHECK_LAUNCH();
  k_batch_inv<<<div_up(div_up(N, 8), 128), 128, 0, st>>>(tmpA, N);
  GPW_CHECK_LAUNCH();
  k_mul_arrays<<<G, 256, 0, st>>>(z_e, tmpA, N);
  GPW_CHECK_LAUNCH();
  GPW_TRY(prefix_product(k, z_e, tmpB));
  GPW_CUDA(cudaMemcpyAsync(z_c, z_e, (size_t)N * sizeof(Fr), cudaMemcpyDeviceToDevice, st));
  GPW_TRY(intt(k, z_c));
  GPW_TRY(commit(k, z_c, &com[4]));
  ctx->launches += 12;
  GPW_CUDA(mark(3));
  // ---- round 3: quotient on four cosets -----------------------------------------------------------------------------------
  tr.fr(gamma);
  tr.g1(com[4]);
  const Fr alpha = tr.challenge("gpw-plonk-alpha");
  const Fr g = fr_u64(5), rho = root_of_unity(k->logN + 2);
  const Fr* srcs[6] = {a_c, b_c, c_c, z_c, p2_c, pi_c};
  Fr* ce[9];  // six coset-evaluation buffers of the proof's own polynomials (+ three more work arrays used by round 5)
  for (int i = 0; i < 9; i++) ce[i] = k->buf[15 + i];
  for (int j = 0; j < 4; j++) {
    const Fr shift = mul(g, pow_u64(rho, (uint64_t)j));
    for (int p = 0; p < 6; p++) {
      GPW_CUDA(cudaMemcpyAsync(ce[p], srcs[p], (size_t)N * sizeof(Fr), cudaMemcpyDeviceToDevice, st));
      k_scale_pow<<<GP, 128, 0, st>>>(ce[p], N, shift, Fr::one());
      GPW_CHECK_LAUNCH();
      GPW_TRY(ntt(k, ce[p]));
    }
    Fr sN = shift;
    for (int i = 0; i < k->logN; i++) sN = sqr(sN);
    const Fr zh_inv = inv(sub(sN, Fr::one()));
    Fr* const* f = k->fixed_ce + j * 10;
    CosetEvals e{ce[0], ce[1], ce[2], ce[3], ce[4], ce[5], f[0], f[1], f[2], f[3], f[4], f[5], f[6], f[7], f[8], f[9]};
    k_quotient<<<G, 256, 0, st>>>(e, N, k->omega_pow, shift, beta, gamma, alpha, k->k1, k->k2, zh_inv, d[j]);
    GPW_CHECK_LAUNCH();
    GPW_TRY(intt(k, d[j]));
    k_scale_pow<<<GP, 128, 0, st>>>(d[j], N, inv(shift), Fr::one());
    GPW_CHECK_LAUNCH();
    ctx->launches += 14;
  }
  {
    Fr gN = g;
    for (int i = 0; i < k->logN; i++) gN = sqr(gN);
    Fr i4 = rho;
    for (int i = 0; i < k->logN; i++) i4 = sqr(i4);  // rho^N: a primitive fourth root of unity
    k_combine_t<<<G, 256, 0, st>>>(d[0], d[1], d[2], d[3], N, inv(i4), inv(fr_u64(4)), inv(gN));
    GPW_CHECK_LAUNCH();
  }
  for (int i = 0; i < 3; i++) GPW_TRY(commit(k, d[i], &com[5 + i]));
  GPW_CUDA(mark(4));
  // ---- round 4: evaluations ------------------------------------------------------------------------------------------------
  tr.fr(alpha);
  for (int i = 0; i < 3; i++) tr.g1(com[5 + i]);
  const Fr zeta = tr.challenge("gpw-plonk-zeta");
  const Fr* open_c[17] = {a_c, b_c, c_c, z_c, p2_c, k->sel_c[0], k->sel_c[1], k->sel_c[2], k->sel_c[3], k->sel_c[4], k->sel_c[5],
                          k->sig_c[0], k->sig_c[1], k->sig_c[2], d[0], d[1], d[2]};
  Fr evals[18];
  for (int i = 0; i < 17; i++) GPW_TRY(eval_at(k, open_c[i], zeta, tmpA, &evals[i]));
  const Fr zeta_w = mul(zeta, k->omega);
  GPW_TRY(eval_at(k, z_c, zeta_w, tmpA, &evals[17]));
  {  // the quotient must have no fourth part: t3 == 0 identically iff the witness satisfies every gate and copy constraint
    Fr t3;
    GPW_TRY(eval_at(k, d[3], zeta, tmpA, &t3));
    if (!t3.is_zero()) {
      set_error("plonk_prove: the quotient is not a polynomial of degree < 3N (constraint system not satisfied)");
      for (auto& e : ev) cudaEventDestroy(e);
      return GPW_EUNSAT;
    }
  }
  GPW_CUDA(mark(5));
  // ---- round 5: batched openings (built on H, interpolated, committed) --------------------------------------------------------
  tr.fr(zeta);
  for (const Fr& e : evals) tr.fr(e);
  const Fr nu = tr.challenge("gpw-plonk-nu");
  // evaluations on H of the 17 polynomials: wires / Z / P2 / selectors / sigmas are at hand, t0..t2 need a forward NTT
  Fr* t_e[3] = {ce[0], ce[1], ce[2]};
  for (int i = 0; i < 3; i++) {
    GPW_CUDA(cudaMemcpyAsync(t_e[i], d[i], (size_t)N * sizeof(Fr), cudaMemcpyDeviceToDevice, st));
    GPW_TRY(ntt(k, t_e[i]));
  }
  const Fr* open_e[17] = {a_e, b_e, c_e, z_e, p2_e, k->sel_e[0], k->sel_e[1], k->sel_e[2], k->sel_e[3], k->sel_e[4], k->sel_e[5],
                          k->sig_e[0], k->sig_e[1], k->sig_e[2], t_e[0], t_e[1], t_e[2]};
  Fr* W = ce[3];
  Fr* den = ce[4];
  GPW_CUDA(cudaMemsetAsync(W, 0, (size_t)N * sizeof(Fr), st));
  Fr nu_k = Fr::one();
  for (int i = 0; i < 17; i++) {
    k_open_accumulate<<<G, 256, 0, st>>>(W, open_e[i], N, evals[i], nu_k);
    GPW_CHECK_LAUNCH();
    nu_k = mul(nu_k, nu);
  }
  k_open_denominators<<<G, 256, 0, st>>>(k->omega_pow, N, zeta, den);
  k_batch_inv<<<div_up(div_up(N, 8), 128), 128, 0, st>>>(den, N);
  k_mul_arrays<<<G, 256, 0, st>>>(W, den, N);
  GPW_CHECK_LAUNCH();
  GPW_TRY(intt(k, W));
  GPW_TRY(commit(k, W, &com[8]));
  GPW_CUDA(cudaMemsetAsync(W, 0, (size_t)N * sizeof(Fr), st));
  k_open_accumulate<<<G, 256, 0, st>>>(W, z_e, N, evals[17], Fr::one());
  k_open_denominators<<<G, 256, 0, st>>>(k->omega_pow, N, zeta_w, den);
  k_batch_inv<<<div_up(div_up(N, 8), 128), 128, 0, st>>>(den, N);
  k_mul_arrays<<<G, 256, 0, st>>>(W, den, N);
  GPW_CHECK_LAUNCH();
  GPW_TRY(intt(k, W));
  GPW_TRY(commit(k, W, &com[9]));
  ctx->launches += 17 + 8;
  GPW_CUDA(mark(6));
  GPW_CUDA(cudaStreamSynchronize(st));
  for (int i = 0; i < 6; i++) cudaEventElapsedTime(&k->t_ms[i], ev[i], ev[i + 1]);
  for (auto& e : ev) cudaEventDestroy(e);
  uint8_t* o = out_proof;
  for (int i = 0; i < 10; i++, o += 64) g1_be(com[i], o);
  for (int i = 0; i < 18; i++, o += 32) fr_be(evals[i], o);
  return GPW_OK;
}

// ms of the last prove: witness + P2, wire columns, grand product, quotient, evaluations, openings
extern "C" int gpw_plonk_last_stats(const gpw_plonk_key* k, float* ms6) {
  if (!k || !ms6) return GPW_EINVAL;
  for (int i = 0; i < 6; i++) ms6[i] = k->t_ms[i];
  return GPW_OK;
}
