// End-to-end wrap prover for one Plonky2 proof: JSON -> witness (GPU tape) -> BSB22-style commitment of the
// range-check limbs -> log-derivative argument -> R1CS evaluation -> Groth16 proof.
// This is the drop-in for the reference's
//     witness, _ := frontend.NewWitness(&assignment, ecc.BN254.ScalarField())
//     proof, _   := groth16.Prove(r1cs, pk, witness)            (benchmark.go:240-249)
// with the circuit from gpw_circuit_compile_verifier standing in for frontend.Compile (benchmark.go:55) and
// gpw_wrap_key_synthetic for groth16.DummySetup (benchmark.go:214).
#include <sys/random.h>

#include <atomic>
#include <cstring>
#include <mutex>
#include <thread>

#include "common.cuh"
#include "ec.cuh"
#include "host_ec.cuh"
#include "host/frontend.h"

struct gpw_circuit;
extern "C" {
int gpw_circuit_info(const gpw_circuit* c, uint64_t* info16);
int gpw_circuit_parse_inputs(const gpw_circuit* c, const char* p, const char* v, uint64_t* out, size_t cap);
int gpw_witness_solve_phase1_on(gpw_circuit* c, gpw_ctx* lane, uint64_t inputs_dev, int n_proofs, uint64_t wires_dev, size_t wire_stride);
int gpw_witness_solve_phase2_on(gpw_circuit* c, gpw_ctx* lane, const uint64_t* ch, int n_proofs, uint64_t wires_dev, size_t wire_stride);
int gpw_r1cs_eval_on(gpw_circuit* c, gpw_ctx* lane, uint64_t wires_dev, uint64_t a_dev, uint64_t b_dev, uint64_t c_dev, uint64_t* n_unsat);
int gpw_circuit_supports(const gpw_circuit* c, int side, uint32_t* out, size_t cap, size_t* n);
int gpw_ntt_share_tables(gpw_ctx* from, gpw_ctx* to, int logn);
}

namespace gpw {

// ---- SHA-256 + expand_message_xmd (RFC 9380) -> Fr : the commitment challenge ---------------------------------
// gnark derives the BSB22 challenge with gnark-crypto's fr.Hash(msg, dst = "bsb22-commitment", 1): expand_message_xmd
// over SHA-256 to 48 bytes, interpreted big-endian and reduced mod r (recalled from gnark v0.9.1 - the sources are not
// in the reference tree).
struct Sha256 {
  uint32_t h[8];
  uint8_t buf[64];
  uint64_t len = 0;
  size_t fill = 0;
  Sha256() {
    static const uint32_t iv[8] = {0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a, 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19};
    memcpy(h, iv, sizeof(iv));
  }
  static uint32_t rotr(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }
  void block(const uint8_t* p) {
    static const uint32_t K[64] = {
        0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5, 0xd807aa98, 0x12835b01,
        0x243185be, 0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174, 0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc,
        0x2de92c6f, 0x4a7484aa, 0x5cb0a9dc, 0x76f988da, 0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147,
        0x06ca6351, 0x14292967, 0x27b70a85, 0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85,
        0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3, 0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070, 0x19a4c116, 0x1e376c08,
        0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f, 0x682e6ff3, 0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208,
        0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2};
    uint32_t w[64];
    for (int i = 0; i < 16; i++) w[i] = (uint32_t)p[4 * i] << 24 | (uint32_t)p[4 * i + 1] << 16 | (uint32_t)p[4 * i + 2] << 8 | p[4 * i + 3];
    for (int i = 16; i < 64; i++) {
      uint32_t s0 = rotr(w[i - 15], 7) ^ rotr(w[i - 15], 18) ^ (w[i - 15] >> 3);
      uint32_t s1 = rotr(w[i - 2], 17) ^ rotr(w[i - 2], 19) ^ (w[i - 2] >> 10);
      w[i] = w[i - 16] + s0 + w[i - 7] + s1;
    }
    uint32_t a = h[0], b = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
    for (int i = 0; i < 64; i++) {
      uint32_t S1 = rotr(e, 6) ^ rotr(e, 11) ^ rotr(e, 25);
      uint32_t ch = (e & f) ^ (~e & g);
      uint32_t t1 = hh + S1 + ch + K[i] + w[i];
      uint32_t S0 = rotr(a, 2) ^ rotr(a, 13) ^ rotr(a, 22);
      uint32_t mj = (a & b) ^ (a & c) ^ (b & c);
      uint32_t t2 = S0 + mj;
      hh = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
    }
    h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += g; h[7] += hh;
  }
  void update(const uint8_t* p, size_t n) {
    len += n;
    while (n) {
      size_t k = std::min(n, 64 - fill);
      memcpy(buf + fill, p, k);
      fill += k; p += k; n -= k;
      if (fill == 64) { block(buf); fill = 0; }
    }
  }
  void final(uint8_t out[32]) {
    uint64_t bits = len * 8;
    uint8_t pad = 0x80;
    update(&pad, 1);
    uint8_t z = 0;
    while (fill != 56) update(&z, 1);
    uint8_t lb[8];
    for (int i = 0; i < 8; i++) lb[i] = (uint8_t)(bits >> (56 - 8 * i));
    update(lb, 8);
    for (int i = 0; i < 8; i++) { out[4 * i] = h[i] >> 24; out[4 * i + 1] = h[i] >> 16; out[4 * i + 2] = h[i] >> 8; out[4 * i + 3] = h[i]; }
  }
};

void sha256_bytes(const uint8_t* msg, size_t len, uint8_t out[32]) {
  Sha256 s;
  s.update(msg, len);
  s.final(out);
}

static void sha256(const std::vector<uint8_t>& m, uint8_t out[32]) {
  Sha256 s;
  s.update(m.data(), m.size());
  s.final(out);
}

// expand_message_xmd(msg, dst, 48) then big-endian mod r -> canonical limbs
void hash_to_fr(const uint8_t* msg, size_t msg_len, const char* dst, uint64_t out_canonical[4]) {
  const size_t L = 48, dst_len = strlen(dst);
  std::vector<uint8_t> dst_prime(dst, dst + dst_len);
  dst_prime.push_back((uint8_t)dst_len);
  std::vector<uint8_t> m(64, 0);  // Z_pad
  m.insert(m.end(), msg, msg + msg_len);
  m.push_back(0);
  m.push_back((uint8_t)L);
  m.push_back(0);
  m.insert(m.end(), dst_prime.begin(), dst_prime.end());
  uint8_t b0[32], b1[32], b2[32];
  sha256(m, b0);
  std::vector<uint8_t> t(b0, b0 + 32);
  t.push_back(1);
  t.insert(t.end(), dst_prime.begin(), dst_prime.end());
  sha256(t, b1);
  std::vector<uint8_t> t2(32);
  for (int i = 0; i < 32; i++) t2[i] = b0[i] ^ b1[i];
  t2.push_back(2);
  t2.insert(t2.end(), dst_prime.begin(), dst_prime.end());
  sha256(t2, b2);
  uint8_t u[48];
  memcpy(u, b1, 32);
  memcpy(u + 32, b2, 16);
  // big-endian 384-bit integer mod r, via Horner in Fr (Montgomery arithmetic)
  Fr acc = Fr::zero();
  const Fr c256 = fe::fr_from_u64(256);
  for (size_t i = 0; i < L; i++) acc = add(mul(acc, c256), fe::fr_from_u64(u[i]));
  fe::fr_to_limbs(acc, out_canonical);
}

__global__ void k_gather_fr(const Fr* __restrict__ w, const uint32_t* __restrict__ idx, size_t n, Fr* __restrict__ out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint4* s = reinterpret_cast<const uint4*>(w + idx[i]);
  uint4* d = reinterpret_cast<uint4*>(out + i);
  d[0] = s[0];
  d[1] = s[1];
}

}  // namespace gpw

using namespace gpw;

extern "C" {
int gpw_ec_generator_multiples_dev(gpw_ctx* ctx, int group, uint64_t k0, size_t n, uint64_t out_dev);
int gpw_msm_g1_dev(gpw_ctx* ctx, uint64_t s, uint64_t p, size_t n, int mont, int c, int lo, int hi, uint64_t* out);
int gpw_msm_g2_dev(gpw_ctx* ctx, uint64_t s, uint64_t p, size_t n, int mont, int c, int lo, int hi, uint64_t* out);
int gpw_groth16_compute_h_dev(gpw_ctx* ctx, uint64_t a_dev, uint64_t b_dev, uint64_t c_dev, int logN);
int gpw_msm_g1_fixed_table(gpw_ctx* ctx, uint64_t points_dev, size_t n, int window_bits, int n_windows, uint64_t table_dev);
int gpw_msm_g1_fixed_dev(gpw_ctx* ctx, uint64_t s, uint64_t table, size_t n, int mont, int c, int n_windows, uint64_t* out);
int gpw_msm_g1_shared_dev(gpw_ctx* ctx, uint64_t s, uint64_t p, size_t n, int mont, int c, int fixed_windows, const char* sort_tag, int reuse,
                          uint64_t* out);
int gpw_msm_g2_shared_dev(gpw_ctx* ctx, uint64_t s, uint64_t p, size_t n, int mont, int c, const char* sort_tag, int reuse, uint64_t* out);
int gpwi_msm_defer_begin(gpw_ctx* ctx);
int gpwi_msm_finish(gpw_ctx* ctx, int idx);
int gpwi_msm_defer_end(gpw_ctx* ctx);
}

#include "wrap_internal.cuh"

static int wk_alloc(void** p, size_t bytes) {
  cudaError_t e = cudaMalloc(p, bytes ? bytes : 16);
  if (e != cudaSuccess) {
    set_error("cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
    return GPW_ENOMEM;
  }
  return GPW_OK;
}

static void lane_free(WrapLane* l) {
  if (!l) return;
  void* ps[] = {l->wires, l->va, l->vb, l->vc, l->gathA, l->gathB, l->inputs_dev};
  for (void* p : ps) cudaFree(p);
  if (l->done) cudaEventDestroy(l->done);
  for (cudaEvent_t e : l->tev)
    if (e) cudaEventDestroy(e);
  if (l->own_ctx) gpw_ctx_destroy(l->ctx);
  delete l;
}

// ctx == nullptr: the lane gets a context of its own
static int lane_create(gpw_wrap_key* k, gpw_ctx* ctx, WrapLane** out) {
  WrapLane* l = new WrapLane();
  if (ctx) {
    l->ctx = ctx;
  } else {
    int rc = gpw_ctx_create(k->ctx->device, &l->ctx);
    if (rc != GPW_OK) {
      delete l;
      return rc;
    }
    l->own_ctx = true;
    rc = gpw_ntt_share_tables(k->ctx, l->ctx, k->logN);
    if (rc != GPW_OK) {
      lane_free(l);
      return rc;
    }
  }
  const size_t N = (size_t)1 << k->logN;
  int rc = 0;
  if ((rc = wk_alloc((void**)&l->wires, (size_t)k->m * sizeof(Fr))) || (rc = wk_alloc((void**)&l->va, N * sizeof(Fr))) ||
      (rc = wk_alloc((void**)&l->vb, N * sizeof(Fr))) || (rc = wk_alloc((void**)&l->vc, N * sizeof(Fr))) ||
      (rc = wk_alloc((void**)&l->gathA, (size_t)k->nA * sizeof(Fr))) || (rc = wk_alloc((void**)&l->gathB, (size_t)k->nB * sizeof(Fr))) ||
      (rc = wk_alloc((void**)&l->inputs_dev, (size_t)k->n_inputs * 32))) {
    lane_free(l);
    return rc;
  }
  bool ev_ok = cudaEventCreateWithFlags(&l->done, cudaEventDisableTiming) == cudaSuccess;
  for (cudaEvent_t& e : l->tev) ev_ok = ev_ok && cudaEventCreate(&e) == cudaSuccess;
  if (!ev_ok) {
    set_error("cudaEventCreate failed");
    lane_free(l);
    return GPW_ECUDA;
  }
  *out = l;
  return GPW_OK;
}

extern "C" void gpw_wrap_key_free(gpw_wrap_key* k) {
  if (!k) return;
  cudaSetDevice(k->ctx->device);
  for (WrapLane* l : k->lanes) lane_free(l);
  void* ps[] = {k->suppA, k->suppB, k->A, k->B1, k->K, k->Z, k->CK, k->CKs, k->B2, k->Zt, k->K2t, k->At};
  for (void* p : ps) cudaFree(p);
  delete k;
}

template <class F>
static Affine<F> gen_mul_host(uint64_t k) {
  uint32_t kw[8] = {(uint32_t)k, (uint32_t)(k >> 32), 0, 0, 0, 0, 0, 0};
  return to_affine(host_scalar_mul(generator<F>(), kw));
}

int wrap_key_alloc(gpw_ctx* ctx, gpw_circuit* circ, gpw_wrap_key** out) {
  if (!ctx || !circ || !out) {
    set_error("wrap_key: null argument");
    return GPW_EINVAL;
  }
  GPW_CUDA(cudaSetDevice(ctx->device));
  uint64_t info[16];
  GPW_TRY(gpw_circuit_info(circ, info));
  gpw_wrap_key* k = new gpw_wrap_key();
  k->ctx = ctx;
  k->circ = circ;
  k->m = (uint32_t)info[0];
  k->n_pub = (uint32_t)info[1];
  k->n_inputs = (uint32_t)(info[1] + info[2]);
  k->n_cons = (uint32_t)info[3];
  k->limb_start = (uint32_t)info[7];
  k->n_committed = info[6] ? (uint32_t)(info[6] + 65536) : 0;
  k->commit_wire = (uint32_t)info[9];
  k->logN = 1;
  while ((1ull << k->logN) < k->n_cons) k->logN++;
  const size_t N = (size_t)1 << k->logN;
  std::vector<uint32_t>&sa = k->suppA_host, &sb = k->suppB_host;
  sa.resize(k->m);
  sb.resize(k->m);
  size_t na = 0, nb = 0;
  int rc = 0;
  auto fail = [&](int code) {
    gpw_wrap_key_free(k);
    return code;
  };
  if ((rc = gpw_circuit_supports(circ, 0, sa.data(), sa.size(), &na)) || (rc = gpw_circuit_supports(circ, 1, sb.data(), sb.size(), &nb)))
    return fail(rc);
  sa.resize(na);
  sb.resize(nb);
  k->nA = (uint32_t)na;
  k->nB = (uint32_t)nb;
  if ((rc = wk_alloc((void**)&k->suppA, na * 4)) || (rc = wk_alloc((void**)&k->suppB, nb * 4)) ||
      (rc = wk_alloc((void**)&k->A, na * sizeof(G1Affine))) || (rc = wk_alloc((void**)&k->B1, nb * sizeof(G1Affine))) ||
      (rc = wk_alloc((void**)&k->B2, nb * sizeof(G2Affine))) || (rc = wk_alloc((void**)&k->K, (size_t)k->m * sizeof(G1Affine))) ||
      (rc = wk_alloc((void**)&k->Z, N * sizeof(G1Affine))) || (rc = wk_alloc((void**)&k->CK, (size_t)k->n_committed * sizeof(G1Affine))) ||
      (rc = wk_alloc((void**)&k->CKs, (size_t)k->n_committed * sizeof(G1Affine))))
    return fail(rc);
  {
    WrapLane* l0 = nullptr;
    if ((rc = lane_create(k, ctx, &l0))) return fail(rc);
    k->lanes.push_back(l0);
  }
  if (const char* e = getenv("GPW_WRAP_LANES")) k->want_lanes = std::max(1, std::min(WRAP_MAX_LANES, atoi(e)));
  if (cudaMemcpy(k->suppA, sa.data(), na * 4, cudaMemcpyHostToDevice) != cudaSuccess ||
      cudaMemcpy(k->suppB, sb.data(), nb * 4, cudaMemcpyHostToDevice) != cudaSuccess) {
    set_error("wrap_key: support upload failed: %s", cudaGetErrorString(cudaGetLastError()));
    return fail(GPW_ECUDA);
  }
  const uint32_t c_hi = k->n_committed ? k->limb_start + k->n_committed : k->m;
  // the challenge wire (if any) is the first wire behind the committed range
  k->k2_lo = (k->n_committed && k->commit_wire == c_hi && c_hi < k->m) ? c_hi + 1 : c_hi;
  *out = k;
  return GPW_OK;
}

// Fixed-base tables over the filled-in bases (Z, the K range behind the commitment, A's quotient suffix).
int wrap_key_finish(gpw_wrap_key* k) {
  gpw_ctx* ctx = k->ctx;
  const size_t N = (size_t)1 << k->logN;
  const std::vector<uint32_t>& sa = k->suppA_host;
  const size_t na = sa.size();
  int rc = 0;
  auto fail = [&](int code) {
    gpw_wrap_key_free(k);
    return code;
  };
  const char* e = getenv("GPW_FIXED_BASE");
  const uint32_t c_hi = k->n_committed ? k->limb_start + k->n_committed : k->m;
  if (!(e && atoi(e) == 0)) {
    if ((rc = wk_alloc((void**)&k->Zt, (size_t)FIXED_W * (N - 1) * sizeof(G1Affine))) ||
        (k->k2_lo < k->m && (rc = wk_alloc((void**)&k->K2t, (size_t)FIXED_WQ * (k->m - k->k2_lo) * sizeof(G1Affine)))))
      return fail(rc);
    if ((rc = gpw_msm_g1_fixed_table(ctx, (uint64_t)k->Z, N - 1, FIXED_C, FIXED_W, (uint64_t)k->Zt))) return fail(rc);
    if (k->K2t && (rc = gpw_msm_g1_fixed_table(ctx, (uint64_t)(k->K + k->k2_lo), k->m - k->k2_lo, FIXED_CQ, FIXED_WQ, (uint64_t)k->K2t)))
      return fail(rc);
    while (k->nA_tail < na && sa[na - 1 - k->nA_tail] >= c_hi) k->nA_tail++;  // supports are sorted by wire id
    // A's suffix is exactly the wires [k2_lo, m) iff it has that many entries and starts there (sorted, distinct)
    k->share_q_sort = k->K2t && k->nA_tail == k->m - k->k2_lo && k->nA_tail > 0 && sa[na - k->nA_tail] == k->k2_lo;
    if (k->nA_tail) {
      if ((rc = wk_alloc((void**)&k->At, (size_t)FIXED_WQ * k->nA_tail * sizeof(G1Affine)))) return fail(rc);
      if ((rc = gpw_msm_g1_fixed_table(ctx, (uint64_t)(k->A + (na - k->nA_tail)), k->nA_tail, FIXED_CQ, FIXED_WQ, (uint64_t)k->At)))
        return fail(rc);
    }
  }
  if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) {
    set_error("wrap_key: %s", cudaGetErrorString(cudaGetLastError()));
    return fail(GPW_ECUDA);
  }
  return GPW_OK;
}

// Discrete logs (all bases are [k]G): A_j = [1 + j], B1_j = [2^32 + j], B2_j = [1 + j] (G2), K_i = [2^33 + i],
// Z_j = [2^34 + j], CK_i = [2^35 + i], CKs_i = [2^36 + i]; alpha = [seed+1], beta = [seed+2], delta = [seed+3].
// (j indexes the compacted A / B support lists, i is the wire id.)
extern "C" int gpw_wrap_key_synthetic(gpw_ctx* ctx, gpw_circuit* circ, uint64_t seed, gpw_wrap_key** out) {
  gpw_wrap_key* k = nullptr;
  GPW_TRY(wrap_key_alloc(ctx, circ, &k));
  k->seed = seed;
  const size_t N = (size_t)1 << k->logN;
  int rc = 0;
  if ((rc = gpw_ec_generator_multiples_dev(ctx, 1, 1, k->nA, (uint64_t)k->A)) ||
      (rc = gpw_ec_generator_multiples_dev(ctx, 1, 1ull << 32, k->nB, (uint64_t)k->B1)) ||
      (rc = gpw_ec_generator_multiples_dev(ctx, 2, 1, k->nB, (uint64_t)k->B2)) ||
      (rc = gpw_ec_generator_multiples_dev(ctx, 1, 1ull << 33, k->m, (uint64_t)k->K)) ||
      (rc = gpw_ec_generator_multiples_dev(ctx, 1, 1ull << 34, N - 1, (uint64_t)k->Z)) ||
      (rc = gpw_ec_generator_multiples_dev(ctx, 1, (1ull << 35) + k->limb_start, k->n_committed, (uint64_t)k->CK)) ||
      (rc = gpw_ec_generator_multiples_dev(ctx, 1, (1ull << 36) + k->limb_start, k->n_committed, (uint64_t)k->CKs))) {
    gpw_wrap_key_free(k);
    return rc;
  }
  k->alpha1 = gen_mul_host<Fp>(seed + 1);
  k->beta1 = gen_mul_host<Fp>(seed + 2);
  k->delta1 = gen_mul_host<Fp>(seed + 3);
  k->beta2 = gen_mul_host<Fp2>(seed + 2);
  k->delta2 = gen_mul_host<Fp2>(seed + 3);
  GPW_TRY(wrap_key_finish(k));  // frees the key on failure
  *out = k;
  return GPW_OK;
}

// info: [m, n_pub, n_cons, logN, nA, nB, n_committed, limb_start]
extern "C" int gpw_wrap_key_info(const gpw_wrap_key* k, uint64_t* info8) {
  if (!k || !info8) return GPW_EINVAL;
  uint64_t v[8] = {k->m, k->n_pub, k->n_cons, (uint64_t)k->logN, k->nA, k->nB, k->n_committed, k->limb_start};
  memcpy(info8, v, sizeof(v));
  return GPW_OK;
}

extern "C" uint64_t gpw_wrap_key_wires_dev(const gpw_wrap_key* k) { return k ? (uint64_t)k->lanes[0]->wires : 0; }
// after a gpw_wrap_prove: the quotient polynomial's coefficients h_0 .. h_{N-2} (Fr, Montgomery) of that proof - test aid
extern "C" uint64_t gpw_wrap_key_h_dev(const gpw_wrap_key* k) { return k ? (uint64_t)k->lanes[0]->va : 0; }

// Number of proofs gpw_wrap_prove_many keeps in flight (default 6; each lane holds ~1.5 GB of vectors plus its MSM scratch).
extern "C" int gpw_wrap_set_lanes(gpw_wrap_key* k, int n) {
  if (!k || n < 1 || n > WRAP_MAX_LANES) {
    set_error("wrap_set_lanes: n must be in [1, %d]", WRAP_MAX_LANES);
    return GPW_EINVAL;
  }
  k->want_lanes = n;
  return GPW_OK;
}

// uniform in [0, r): 254 random bits, rejection (gnark: fr.Element.SetRandom over crypto/rand)
static int sample_fr(uint64_t out[4]) {
  static const uint64_t RQ[4] = {0x43e1f593f0000001ull, 0x2833e84879b97091ull, 0xb85045b68181585dull, 0x30644e72e131a029ull};
  for (;;) {
    if (getrandom(out, 32, 0) != 32) {
      set_error("getrandom failed");
      return GPW_EINVAL;
    }
    out[3] &= 0x3fffffffffffffffull;
    for (int i = 3; i >= 0; i--) {
      if (out[i] < RQ[i]) return GPW_OK;
      if (out[i] > RQ[i]) break;
    }
  }
}

static void ser_g1_be(const G1Affine& p, uint8_t out[64]) {
  Fp x = from_mont(p.x), y = from_mont(p.y);
  for (int i = 0; i < 8; i++)
    for (int b = 0; b < 4; b++) {
      out[31 - (4 * i + b)] = (uint8_t)(x.l[i] >> (8 * b));
      out[63 - (4 * i + b)] = (uint8_t)(y.l[i] >> (8 * b));
    }
}

// One wrap proof. inputs: n_inputs x 4 u64 canonical on the HOST (gpw_circuit_parse_inputs order). r, s canonical.
// out_proof (u64 x 64): Ar (8) | Bs (16) | Krs (8) | commitment D (8) | commitment PoK (8) | challenge (4, canonical) |
//                       n_unsatisfied (1) | reserved.  If check != 0 the R1CS is verified on the device (a*b == c on every
// row) and GPW_EUNSAT is returned on failure.
static int wrap_stage2(gpw_wrap_key* k, WrapLane* L, const uint64_t* r_canon, const uint64_t* s_canon, int check, uint64_t* out_proof);

// solve phase 1 + everything after it for one proof on lane L (inputs already on the device)
static int wrap_one(gpw_wrap_key* k, WrapLane* L, uint64_t inputs_dev, const uint64_t* r_canon, const uint64_t* s_canon, int check,
                    uint64_t* out_proof) {
  cudaStream_t st = L->ctx->stream;
  cudaEvent_t e0 = L->tev[0], e1 = L->tev[1];
  GPW_CUDA(cudaEventRecord(e0, st));
  // The solve (one-CTA spine segments + the Merkle levels' small grids) runs on the context's HIGH-PRIORITY stream: the spine
  // CTA needs a whole SM to itself and, at equal priority, waits behind every bulk CTA the other lanes have queued before it -
  // with a whole proof's MSMs queued at once (deferred MSMs) that starved it (10.4 proofs/s); in front of them it is 13.6
  // against 13.2 with synchronous MSMs. GPW_SPINE_HI=0: the lane's ordinary stream.
  static const bool spine_hi = !getenv("GPW_SPINE_HI") || atoi(getenv("GPW_SPINE_HI")) != 0;
  const cudaStream_t hs = (spine_hi && L->ctx->stream_hi) ? L->ctx->stream_hi : st;
  if (hs != st) {
    GPW_CUDA(cudaEventRecord(L->ctx->ev_hop, st));
    GPW_CUDA(cudaStreamWaitEvent(hs, L->ctx->ev_hop, 0));
    L->ctx->stream = hs;
  }
  int rc1 = gpw_witness_solve_phase1_on(k->circ, L->ctx, inputs_dev, 1, (uint64_t)L->wires, k->m);
  L->ctx->stream = st;
  GPW_TRY(rc1);
  float t1 = 0;
  GPW_CUDA(cudaEventRecord(e1, hs));
  GPW_CUDA(cudaEventSynchronize(e1));
  GPW_CUDA(cudaEventElapsedTime(&t1, e0, e1));
  int rc = wrap_stage2(k, L, r_canon, s_canon, check, out_proof);
  L->t_ms[0] = t1;
  return rc;
}

extern "C" int gpw_wrap_prove_dev(gpw_wrap_key* k, uint64_t inputs_dev, const uint64_t* r_canon, const uint64_t* s_canon, int check,
                                  uint64_t* out_proof);

extern "C" int gpw_wrap_prove(gpw_wrap_key* k, const uint64_t* inputs, const uint64_t* r_canon, const uint64_t* s_canon, int check,
                              uint64_t* out_proof) {
  if (!k || !inputs) {
    set_error("wrap_prove: null argument");
    return GPW_EINVAL;
  }
  WrapLane* L = k->lanes[0];
  GPW_CUDA(cudaSetDevice(k->ctx->device));
  GPW_CUDA(cudaMemcpyAsync(L->inputs_dev, inputs, (size_t)k->n_inputs * 32, cudaMemcpyHostToDevice, L->ctx->stream));
  return gpw_wrap_prove_dev(k, (uint64_t)L->inputs_dev, r_canon, s_canon, check, out_proof);
}

// Same with the parsed inputs already resident on the device (n_inputs x 4 u64 canonical).
extern "C" int gpw_wrap_prove_dev(gpw_wrap_key* k, uint64_t inputs_dev, const uint64_t* r_canon, const uint64_t* s_canon, int check,
                                  uint64_t* out_proof) {
  if (!k || !inputs_dev || !out_proof) {
    set_error("wrap_prove: null argument");
    return GPW_EINVAL;
  }
  GPW_CUDA(cudaSetDevice(k->ctx->device));
  {
    static const char* overlap_env = getenv("GPW_MSM_OVERLAP");
    const int overlap_mode = overlap_env ? atoi(overlap_env) : k->ctx->msm_overlap_mode;
    k->lanes[0]->ctx->msm_overlap = overlap_mode < 0 ? true : overlap_mode != 0;  // a lone proof: see gpw_wrap_prove_many
  }
  return wrap_one(k, k->lanes[0], inputs_dev, r_canon, s_canon, check, out_proof);
}

// Everything after the first solve phase, on the context's stream: commitment, second solve phase, R1CS evaluation,
// computeH, MSMs, assembly. `wires` = the proof's wire vector (phase 1 complete).
static int wrap_stage2(gpw_wrap_key* k, WrapLane* L, const uint64_t* r_canon, const uint64_t* s_canon, int check, uint64_t* out_proof) {
  gpw_ctx* ctx = L->ctx;
  cudaStream_t st = ctx->stream;
  Fr* wires = L->wires;
  const size_t N = (size_t)1 << k->logN;
  cudaEvent_t* ev = L->tev + 2;  // 7 events
  memset(out_proof, 0, 64 * 8);
  const bool dbg = getenv("GPW_DEBUG_WRAP") != nullptr;
  GPW_CUDA(cudaEventRecord(ev[0], st));
  GPW_CUDA(cudaEventRecord(ev[1], st));
  // The MSMs of the proof are DEFERRED (common.cuh): each call enqueues its sort + accumulation on the lane's stream and its
  // latency-bound tail on the lane's high-priority stream, so the tail of one MSM runs beside the accumulation of the next
  // and the host never waits between them; results are collected by gpwi_msm_finish / gpwi_msm_defer_end.
  // (results first: the guard below may still write them when an error path unwinds)
  G1Affine D{Fp::zero(), Fp::zero()}, PoK{Fp::zero(), Fp::zero()};
  G1Affine mA, mA2{Fp::zero(), Fp::zero()}, mB1, mK1, mK2, mZ;
  G2Affine mB2;
  struct DeferGuard {
    gpw_ctx* c;
    ~DeferGuard() { gpwi_msm_defer_end(c); }
  } defer_guard{ctx};
  GPW_TRY(gpwi_msm_defer_begin(ctx));
  const char* msm_names[gpw_ctx::MAX_PENDING] = {};
  int n_msm = 0;
  // commitment to the committed wires (range-check limbs + multiplicities) and its proof of knowledge
  uint64_t X[4] = {0, 0, 0, 0};
  if (k->n_committed) {
    uint64_t sc = (uint64_t)(wires + k->limb_start);
    // commitment and proof of knowledge: same scalars, one bucket sort
    GPW_TRY(gpw_msm_g1_shared_dev(ctx, sc, (uint64_t)k->CK, k->n_committed, 1, 0, 0, "sortC", 0, (uint64_t*)&D));
    msm_names[n_msm++] = "CK";
    GPW_TRY(gpw_msm_g1_shared_dev(ctx, sc, (uint64_t)k->CKs, k->n_committed, 1, 0, 0, "sortC", 1, (uint64_t*)&PoK));
    msm_names[n_msm++] = "CKs";
    GPW_TRY(gpwi_msm_finish(ctx, 0));  // the challenge needs D now; the proof of knowledge is collected at the end
    uint8_t ser[64];
    ser_g1_be(D, ser);
    hash_to_fr(ser, 64, "bsb22-commitment", X);
  }
  GPW_CUDA(cudaEventRecord(ev[2], st));
  GPW_TRY(gpw_witness_solve_phase2_on(k->circ, ctx, X, 1, (uint64_t)wires, k->m));
  GPW_CUDA(cudaEventRecord(ev[3], st));
  // R1CS evaluation vectors, zero padded to the FFT domain
  GPW_CUDA(cudaMemsetAsync(L->va, 0, N * sizeof(Fr), st));
  GPW_CUDA(cudaMemsetAsync(L->vb, 0, N * sizeof(Fr), st));
  GPW_CUDA(cudaMemsetAsync(L->vc, 0, N * sizeof(Fr), st));
  uint64_t n_bad = 0;
  int rc = gpw_r1cs_eval_on(k->circ, ctx, (uint64_t)wires, (uint64_t)L->va, (uint64_t)L->vb, (uint64_t)L->vc, &n_bad);
  (void)check;  // an unsatisfied system never yields a proof: with it, (A.B - C)/Z_H is not a polynomial and the output would
                // be a well-formed but invalid proof. The count is reported in slot 52, the status is GPW_EUNSAT.
  out_proof[52] = n_bad;
  if (rc != GPW_OK) return rc;
  GPW_CUDA(cudaEventRecord(ev[4], st));
  GPW_TRY(gpw_groth16_compute_h_dev(ctx, (uint64_t)L->va, (uint64_t)L->vb, (uint64_t)L->vc, k->logN));
  GPW_CUDA(cudaEventRecord(ev[5], st));
  // gather the scalars of the A / B supports
  k_gather_fr<<<div_up(k->nA, 256), 256, 0, st>>>(wires, k->suppA, k->nA, L->gathA);
  GPW_CHECK_LAUNCH();
  k_gather_fr<<<div_up(k->nB, 256), 256, 0, st>>>(wires, k->suppB, k->nB, L->gathB);
  GPW_CHECK_LAUNCH();
  ctx->launches += 2;
  {
    const uint32_t head = k->nA - k->nA_tail;
    GPW_TRY(gpw_msm_g1_dev(ctx, (uint64_t)L->gathA, (uint64_t)k->A, head, 1, 0, 0, 0, (uint64_t*)&mA));
    msm_names[n_msm++] = "A";
    if (k->nA_tail) {
      // the same wires as the K2 MSM below (when both tables exist and cover the same range): K2 reuses this sort
      GPW_TRY(gpw_msm_g1_shared_dev(ctx, (uint64_t)(L->gathA + head), (uint64_t)k->At, k->nA_tail, 1, FIXED_CQ, FIXED_WQ, "sortQ", 0,
                                    (uint64_t*)&mA2));
      msm_names[n_msm++] = "A2";
    }
  }
  GPW_TRY(gpw_msm_g1_shared_dev(ctx, (uint64_t)L->gathB, (uint64_t)k->B1, k->nB, 1, 0, 0, "sortB", 0, (uint64_t*)&mB1));
  msm_names[n_msm++] = "B1";
  GPW_TRY(gpw_msm_g2_shared_dev(ctx, (uint64_t)L->gathB, (uint64_t)k->B2, k->nB, 1, 0, "sortB", 1, (uint64_t*)&mB2));
  msm_names[n_msm++] = "B2";
  // K: private wires that are not committed = [1 + n_pub, limb_start) U [limb_start + n_committed, m), minus the challenge wire
  const uint32_t k_lo = 1 + k->n_pub;
  const uint32_t c_lo = k->n_committed ? k->limb_start : k->m, c_hi = c_lo + k->n_committed;
  GPW_TRY(gpw_msm_g1_dev(ctx, (uint64_t)(wires + k_lo), (uint64_t)(k->K + k_lo), c_lo - k_lo, 1, 0, 0, 0, (uint64_t*)&mK1));
  msm_names[n_msm++] = "K1";
  mK2 = G1Affine{Fp::zero(), Fp::zero()};
  if (k->k2_lo < k->m) {
    const uint32_t n2 = k->m - k->k2_lo;
    if (k->K2t)  // share_q_sort: A's gathered suffix IS wires[k2_lo, m) - same scalars, same buffer, same sort
      GPW_TRY(gpw_msm_g1_shared_dev(ctx, k->share_q_sort ? (uint64_t)(L->gathA + (k->nA - k->nA_tail)) : (uint64_t)(wires + k->k2_lo),
                                    (uint64_t)k->K2t, n2, 1, FIXED_CQ, FIXED_WQ, "sortQ", k->share_q_sort ? 1 : 0, (uint64_t*)&mK2));
    else GPW_TRY(gpw_msm_g1_dev(ctx, (uint64_t)(wires + k->k2_lo), (uint64_t)(k->K + k->k2_lo), n2, 1, 0, 0, 0, (uint64_t*)&mK2));
    msm_names[n_msm++] = "K2";
  }
  if (k->Zt) GPW_TRY(gpw_msm_g1_fixed_dev(ctx, (uint64_t)L->va, (uint64_t)k->Zt, N - 1, 1, FIXED_C, FIXED_W, (uint64_t*)&mZ));
  else GPW_TRY(gpw_msm_g1_dev(ctx, (uint64_t)L->va, (uint64_t)k->Z, N - 1, 1, 0, 0, 0, (uint64_t*)&mZ));
  msm_names[n_msm++] = "Z";
  if (ctx->msm_defer) {
    const int n_done = ctx->n_pend;
    GPW_TRY(gpwi_msm_defer_end(ctx));  // collects every result; the lane's stream is ordered behind all tails
    if (dbg)
      for (int i = 0; i < n_done && i < n_msm; i++)
        fprintf(stderr, "[gpw wrap] MSM %-4s n=%8zu total %7.2f ms accumulate %7.2f ms digits %llu\n", msm_names[i], ctx->pend[i].n,
                ctx->pend[i].total_ms, ctx->pend[i].acc_ms, (unsigned long long)ctx->pend[i].digits);
  }
  if (k->nA_tail) {
    G1XYZZ t = G1XYZZ::from_affine(mA);
    add_mixed(t, mA2, false);
    mA = to_affine(t);
  }
  GPW_CUDA(cudaEventRecord(ev[6], st));
  GPW_CUDA(cudaStreamSynchronize(st));
  for (int i = 0; i < 6; i++) GPW_CUDA(cudaEventElapsedTime(&L->t_ms[i], ev[i], ev[i + 1]));
  // assembly (gnark groth16.Prove). r, s: the caller's (reproducible proofs, tests) or, as gnark does,
  // fresh from the OS CSPRNG when NULL.
  uint64_t r_own[4], s_own[4];
  if (!r_canon) {
    GPW_TRY(sample_fr(r_own));
    r_canon = r_own;
  }
  if (!s_canon) {
    GPW_TRY(sample_fr(s_own));
    s_canon = s_own;
  }
  uint32_t rw[8], sw[8];
  memcpy(rw, r_canon, 32);
  memcpy(sw, s_canon, 32);
  G1XYZZ Ar = G1XYZZ::from_affine(k->alpha1);
  add_mixed(Ar, mA, false);
  add_full(Ar, host_scalar_mul(k->delta1, rw));
  G1XYZZ Bs1 = G1XYZZ::from_affine(k->beta1);
  add_mixed(Bs1, mB1, false);
  add_full(Bs1, host_scalar_mul(k->delta1, sw));
  G2XYZZ Bs = G2XYZZ::from_affine(k->beta2);
  add_mixed(Bs, mB2, false);
  add_full(Bs, host_scalar_mul(k->delta2, sw));
  G1Affine ArA = to_affine(Ar), Bs1A = to_affine(Bs1);
  G1XYZZ Krs = G1XYZZ::from_affine(mK1);
  add_mixed(Krs, mK2, false);
  add_mixed(Krs, mZ, false);
  add_full(Krs, host_scalar_mul(ArA, sw));
  add_full(Krs, host_scalar_mul(Bs1A, rw));
  Fr rf, sf;
  memcpy(&rf, r_canon, 32);
  memcpy(&sf, s_canon, 32);
  Fr rs = from_mont(mul(to_mont(rf), to_mont(sf)));
  uint32_t rsw[8];
  memcpy(rsw, &rs, 32);
  add_full(Krs, neg(host_scalar_mul(k->delta1, rsw)));
  G1Affine KrsA = to_affine(Krs);
  G2Affine BsA = to_affine(Bs);
  memcpy(out_proof, &ArA, 64);
  memcpy(out_proof + 8, &BsA, 128);
  memcpy(out_proof + 24, &KrsA, 64);
  memcpy(out_proof + 32, &D, 64);
  memcpy(out_proof + 40, &PoK, 64);
  // note: slots 40..47 hold the PoK; the challenge and the unsatisfied count follow
  memcpy(out_proof + 48, X, 32);
  out_proof[52] = n_bad;
  return GPW_OK;
}

// ms: [inputs + solve phase 1, commitment MSMs + hash, solve phase 2, R1CS evaluation, computeH, gathers + MSMs]
extern "C" int gpw_wrap_last_stats(const gpw_wrap_key* k, float* ms6) {
  if (!k || !ms6) return GPW_EINVAL;
  for (int i = 0; i < 6; i++) ms6[i] = k->lanes[0]->t_ms[i];
  return GPW_OK;
}

// A stream of n proofs, `lanes` of them in flight at a time (gpw_wrap_set_lanes / GPW_WRAP_LANES, default 4): one host
// thread per lane takes the next unproved input, copies it to the device and runs the whole wrap on the lane's own
// stream. Proofs are independent, so nothing is exchanged between lanes; the device interleaves their kernels.
// inputs: n x n_inputs x 4 u64 canonical (host or device memory); r, s: n x 4 u64 each (host); out: n x 64 u64 (host).
extern "C" int gpw_wrap_prove_many(gpw_wrap_key* k, const uint64_t* inputs, int n, const uint64_t* r_canon, const uint64_t* s_canon,
                                   int check, uint64_t* out_proofs) {
  if (!k || !inputs || n < 1 || !out_proofs) {
    set_error("wrap_prove_many: bad argument");
    return GPW_EINVAL;
  }
  GPW_CUDA(cudaSetDevice(k->ctx->device));
  const int n_lanes = std::min(k->want_lanes, n);
  while ((int)k->lanes.size() < n_lanes) {
    WrapLane* l = nullptr;
    GPW_TRY(lane_create(k, nullptr, &l));
    k->lanes.push_back(l);
  }
  // Deferred MSMs (tail of one MSM beside the accumulation of the next, no host wait in between) shorten a lone proof by
  // ~10 ms and, with the solve spines on the high-priority stream (wrap_one), raise the throughput of a stream of proofs from
  // 13.2 to 13.6 proofs/s - saturated from 3 lanes on instead of 6. (Without the prioritised spine, queueing a whole proof's
  // kernels at once starved the one-SM spines behind hundreds of bulk CTAs: 10.4 proofs/s.) Option "msm_overlap" of the key's
  // context / GPW_MSM_OVERLAP: -1 automatic (on), 0 off, 1 on.
  static const char* overlap_env = getenv("GPW_MSM_OVERLAP");
  const int overlap_mode = overlap_env ? atoi(overlap_env) : k->ctx->msm_overlap_mode;
  for (int j = 0; j < n_lanes; j++) k->lanes[j]->ctx->msm_overlap = overlap_mode != 0;
  const size_t in_words = (size_t)k->n_inputs * 4;
  std::atomic<int> next{0};
  std::atomic<int> first_rc{GPW_OK};
  std::mutex err_mu;
  std::string err_text;
  auto worker = [&](WrapLane* L) {
    cudaSetDevice(k->ctx->device);
    for (;;) {
      const int i = next.fetch_add(1);
      if (i >= n || first_rc.load() != GPW_OK) break;
      int rc = GPW_OK;
      if (cudaMemcpyAsync(L->inputs_dev, inputs + (size_t)i * in_words, in_words * 8, cudaMemcpyDefault, L->ctx->stream) != cudaSuccess) {
        set_error("wrap_prove_many: input copy failed: %s", cudaGetErrorString(cudaGetLastError()));
        rc = GPW_ECUDA;
      }
      if (rc == GPW_OK) rc = wrap_one(k, L, (uint64_t)L->inputs_dev, r_canon ? r_canon + 4 * i : nullptr, s_canon ? s_canon + 4 * i : nullptr, check,
                                     out_proofs + 64 * (size_t)i);
      if (rc == GPW_OK && getenv("GPW_DEBUG_LANES"))
        fprintf(stderr, "[gpw lanes] proof %2d: solve1 %6.1f | commit %5.1f solve2 %5.1f r1cs %5.1f H %5.1f msm %6.1f ms\n", i, L->t_ms[0],
                L->t_ms[1], L->t_ms[2], L->t_ms[3], L->t_ms[4], L->t_ms[5]);
      if (rc != GPW_OK) {
        int expected = GPW_OK;
        if (first_rc.compare_exchange_strong(expected, rc)) {
          std::lock_guard<std::mutex> g(err_mu);
          err_text = std::string("proof ") + std::to_string(i) + ": " + gpw_last_error();  // set_error is thread-local
        }
        break;
      }
    }
    cudaEventRecord(L->done, L->ctx->stream);
  };
  std::vector<std::thread> threads;
  for (int j = 1; j < n_lanes; j++) threads.emplace_back(worker, k->lanes[j]);
  worker(k->lanes[0]);
  for (auto& t : threads) t.join();
  // order the caller's stream after every lane (lane 0 already is the caller's stream)
  for (int j = 1; j < n_lanes; j++) {
    cudaStreamWaitEvent(k->ctx->stream, k->lanes[j]->done, 0);
    k->ctx->launches += k->lanes[j]->ctx->launches;
    k->lanes[j]->ctx->launches = 0;
  }
  if (first_rc.load() != GPW_OK) set_error("%s", err_text.c_str());
  return first_rc.load();
}

extern "C" int gpw_hash_to_fr(const uint8_t* msg, size_t len, const char* dst, uint64_t* out_canonical) {
  if (!msg || !dst || !out_canonical) return GPW_EINVAL;
  hash_to_fr(msg, len, dst, out_canonical);
  return GPW_OK;
}
