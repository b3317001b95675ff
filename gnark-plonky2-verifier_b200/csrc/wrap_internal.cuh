// Internal layout of the wrap proving key, shared by wrap.cu (prover) and setup.cu (trusted setup, key / proof
// serialisation). Not part of the C ABI.
#pragma once
#include <string>
#include <vector>

#include "common.cuh"
#include "ec.cuh"

struct gpw_circuit;
using namespace gpw;

// Proving key for a compiled circuit. Bases are synthetic (known discrete logs, documented below) - the analogue of
// groth16.DummySetup - but have exactly the shapes a real key has for THIS circuit: A / B bases only for wires that
// occur in some L / R row (gnark's pk.InfinityA / InfinityB filtering), K bases for private non-committed wires, a
// Pedersen commitment basis (+ its sigma-twin for the proof of knowledge) for the committed wires, Z for h.
// A proving LANE = one proof in flight: its own context (stream + scratch memory), wire vector, evaluation vectors
// and gathered scalars. Lane 0 lives on the key's context and serves the single-proof entry points; gpw_wrap_prove_many
// runs one host thread per lane, so that the sequential solve spine of one proof (one SM), the host-side glue of another
// (Horner over window sums, proof assembly) and the MSMs / NTTs of the others overlap on the device.
struct WrapLane {
  gpw_ctx* ctx = nullptr;
  bool own_ctx = false;
  Fr *wires = nullptr, *va = nullptr, *vb = nullptr, *vc = nullptr, *gathA = nullptr, *gathB = nullptr;
  uint64_t* inputs_dev = nullptr;
  cudaEvent_t done = nullptr;
  cudaEvent_t tev[9] = {};  // phase timing events of a wrap ([0,1]: solve phase 1, [2..8]: stage 2), created once per lane
  float t_ms[8] = {0, 0, 0, 0, 0, 0, 0, 0};
};
constexpr int WRAP_DEFAULT_LANES = 6, WRAP_MAX_LANES = 16;
// window widths of the fixed-base tables: 22 bits (12 additions per scalar, 2^21 buckets) for the 8.4 M-point Z MSM; 20 bits
// (13 additions, 2^19 buckets) for the 2.5 M-point quotient ranges of A and K, where the reduction of 2^21 buckets would
// cost more than the thirteenth addition
constexpr int FIXED_C = 22, FIXED_W = (254 + FIXED_C) / FIXED_C;
constexpr int FIXED_CQ = 20, FIXED_WQ = (254 + FIXED_CQ) / FIXED_CQ;

struct gpw_wrap_key {
  gpw_ctx* ctx = nullptr;
  gpw_circuit* circ = nullptr;
  uint32_t m = 0, n_pub = 0, n_cons = 0;
  int logN = 0;
  uint32_t limb_start = 0, n_committed = 0, commit_wire = 0;
  uint32_t nA = 0, nB = 0;
  uint32_t *suppA = nullptr, *suppB = nullptr;  // device wire-id lists
  G1Affine *A = nullptr, *B1 = nullptr, *K = nullptr, *Z = nullptr, *CK = nullptr, *CKs = nullptr;
  G2Affine* B2 = nullptr;
  // fixed-base tables (2^(22 w) P, 12 windows) for the two MSMs whose scalars are full-width field elements: Z (the
  // quotient coefficients h) and the K range behind the committed wires (the log-derivative quotients). 12 instead of
  // 16 bucket additions per scalar; 8.4 GB of HBM. GPW_FIXED_BASE=0 keeps the plain windowed MSMs.
  G1Affine *Zt = nullptr, *K2t = nullptr, *At = nullptr;
  // Behind the committed wires sit the commitment-challenge wire and then the log-derivative quotients. The challenge
  // is a public input of the verifier (gnark appends commitment wires to the public witness, their bases live in vk.K),
  // so it is not part of the prover's K MSM: the second K range starts at k2_lo, right behind it. That range is exactly
  // the scalars of A's fixed-base suffix, whose bucket sort the K MSM then reuses.
  uint32_t k2_lo = 0;
  bool share_q_sort = false;
  uint32_t nA_tail = 0;  // the last nA_tail wires of A's support are the log-derivative quotients too (they are the L side
                         // of their own division constraints): that suffix of the A MSM also runs fixed-base
  G1Affine alpha1, beta1, delta1;
  G2Affine beta2, delta2;
  uint32_t n_inputs = 0;
  std::vector<WrapLane*> lanes;
  int want_lanes = WRAP_DEFAULT_LANES;
  uint64_t seed = 0;
  // ---- verifying key (real setup only; csrc/setup.cu). vkK = the gamma-side bases of the verifier's public vector
  // (ONE, the public inputs, then the commitment challenge), gnark's vk.G1.K; ped_* = the Pedersen verifying key of the
  // BSB22 commitment (gnark-crypto pedersen.VerifyingKey{G, GRootSigmaNeg}).
  bool real = false;
  G2Affine gamma2;
  std::vector<G1Affine> vkK;
  G2Affine ped_g, ped_g_root_sigma_neg;
  std::vector<uint32_t> suppA_host, suppB_host;
};

// wrap.cu: allocation of a key for a compiled circuit (supports uploaded, lane 0 created, bases uninitialised) and the
// fixed-base tables over its bases once they are filled in. Every failure frees the key.
int wrap_key_alloc(gpw_ctx* ctx, gpw_circuit* circ, gpw_wrap_key** out);
int wrap_key_finish(gpw_wrap_key* k);
namespace gpw {
void hash_to_fr(const uint8_t* msg, size_t msg_len, const char* dst, uint64_t out_canonical[4]);
void sha256_bytes(const uint8_t* msg, size_t len, uint8_t out[32]);
}

