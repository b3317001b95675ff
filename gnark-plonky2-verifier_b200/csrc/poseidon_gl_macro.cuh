// Native evaluation of one Poseidon-Goldilocks permutation that EMITS, in the order the circuit builder creates them
// (csrc/host/gadgets_core.cc PoseidonGlChip::Poseidon, mirroring poseidon/goldilocks.go:30-37,92-275), the value of
// every wire the solver's OP_POSEIDON_GL macro instruction owns: the (quotient, remainder) pair of each of the 130
// MulAddHint and 630 ReduceHint calls (goldilocks/base.go:223,284) and the 472 unreduced S-box products - 1992 integers
// of up to 192 bits. All arithmetic is exact 64-bit integer arithmetic on the Goldilocks values (gl.cuh); nothing here
// touches BN254 Fr, the caller converts the emitted integers to Montgomery form when it stores them.
//
// Slot layout (the creation order of the gadget):
//   full round f (8 of them, 144 slots each):  [0,24)    gl.Add(state[k], C(rc))        -> (q, r) of element k at 2k
//                                              [24,120)  sBoxMonomial(state[k])         -> x^2, x^3, (q, r), x3r^2, x * x3r^2, (q, r) at 24 + 8k
//                                              [120,144) mdsRowShf(r, state)            -> (q, r) of row r at 120 + 2r
//   partial rounds (840 slots):                [0,24)    gl.Add(state[k], C(first[k]))
//                                              [24,48)   mdsPartialLayerInit            -> (q, r) of result[d] at 24 + 2d
//                                              22 x 36:  S-box of state[0] (8), gl.Add(state[0], C(rc)) (2),
//                                                        mdsPartialLayerFast: Reduce(d) (2), Reduce(result[j]) at 12 + 2j (24)
// Host interpreter (tests/hostlib) and CUDA executor share this file; the sequential form is the specification, the
// warp form (lane k owns state element k) must emit the same integers into the same slots.
#pragma once
#include "gl.cuh"

namespace gpw {
namespace glm {

constexpr uint32_t N_OUT = 1992;
constexpr uint32_t FULL_SLOTS = 144, PARTIAL_BASE = 4 * FULL_SLOTS, PARTIAL_SLOTS = 48 + 22 * 36;
constexpr uint64_t MDS0TO0 = 25;

// constants: one flat table  rc[360] | circ[12] | diag[12] | first[12] | partial_rc[22] | vs[242] | w_hats[242] | init[121]
constexpr uint32_t T_RC = 0, T_CIRC = 360, T_DIAG = 372, T_FIRST = 384, T_PRC = 396, T_VS = 418, T_WHATS = 660, T_INIT = 902,
                   T_TOTAL = 1023;

struct U192 {
  uint64_t l[3];
};

GPW_HD U192 u192(uint64_t a) { return {{a, 0, 0}}; }

// acc += a * b
GPW_HD void mac(U192& acc, uint64_t a, uint64_t b) {
  uint64_t lo, hi;
  gl::mul64(a, b, lo, hi);
  uint64_t s0 = acc.l[0] + lo;
  uint64_t c0 = s0 < lo ? 1u : 0u;
  uint64_t s1 = acc.l[1] + hi;
  uint64_t c1 = s1 < hi ? 1u : 0u;
  uint64_t s1b = s1 + c0;
  c1 += s1b < s1 ? 1u : 0u;
  acc.l[0] = s0;
  acc.l[1] = s1b;
  acc.l[2] += c1;
}

// (a1 * 2^64 + a0) * b, a 128-bit by 64-bit product (< 2^192)
GPW_HD U192 mul128x64(uint64_t a0, uint64_t a1, uint64_t b) {
  U192 r = {{0, 0, 0}};
  uint64_t lo, hi;
  gl::mul64(a0, b, lo, hi);
  r.l[0] = lo;
  r.l[1] = hi;
  gl::mul64(a1, b, lo, hi);
  uint64_t s = r.l[1] + lo;
  r.l[2] = hi + (s < lo ? 1u : 0u);
  r.l[1] = s;
  return r;
}

// ReduceHint (goldilocks/base.go:284-294) specialised to x < p 2^128 (top limb < p; everything the permutation reduces
// is below p^3), so q = floor(x / p) < 2^128. r by three 2^64 = 2^32 - 1 foldings; q = (x - r) * p^-1 mod 2^128 (exact
// division by the odd p) with p^-1 mod 2^128 = (2^64 - 2^32) 2^64 + (2^32 + 1): three 64-bit multiplies instead of the ten
// of the general form. gl::reduce_hint is the general (256-bit) form; the circuit tests hold the two against each other
// (every hint output of the macro is compared with the oracle's trace).
GPW_HD void reduce192(const U192& v, uint64_t& q0, uint64_t& q1, uint64_t& r) {
  uint64_t acc = v.l[2] >= gl::P ? v.l[2] - gl::P : v.l[2];
  acc = gl::reduce128(v.l[1], acc);
  acc = gl::reduce128(v.l[0], acc);
  r = acc;
  const uint64_t d0 = v.l[0] - r;
  const uint64_t d1 = v.l[1] - (v.l[0] < r ? 1u : 0u);
  // q = (d1 2^64 + d0) * ((2^64 - 2^32) 2^64 + (2^32 + 1)) mod 2^128. Written with plain 64-bit multiplies: the
  // hand-expanded shift-and-add form (q0 = d0 + (d0 << 32), q1 = carry + (d0 >> 32) - (d0 << 32) + d1 + (d1 << 32)) is
  // correct on the host but came out of nvcc 12.9 / sm_100a with the "- (d0 << 32)" term added; tools/scratch/
  // t_reduce192.cu checks this function against gl::reduce_hint ON THE DEVICE. Three multiplies do not bound the macro.
  uint64_t hi0;
  gl::mul64(d0, 0x100000001ull, q0, hi0);
  q1 = hi0 + d0 * 0xffffffff00000000ull + d1 * 0x100000001ull;
}

// emits (q, r) of Reduce(v) at slot, slot + 1 and returns r
template <class Emit>
GPW_HD uint64_t reduce_emit(const U192& v, uint32_t slot, Emit& emit) {
  uint64_t q0, q1, r;
  reduce192(v, q0, q1, r);
  emit(slot, U192{{q0, q1, 0}});
  emit(slot + 1, u192(r));
  return r;
}

// gl.Add(x, C(k)) = MulAdd(x, 1, k): emits (q, r), returns r
template <class Emit>
GPW_HD uint64_t addconst_emit(uint64_t x, uint64_t k, uint32_t slot, Emit& emit) {
  uint64_t q, r;
  gl::mul_add_hint(x, 1, k, q, r);
  emit(slot, u192(q));
  emit(slot + 1, u192(r));
  return r;
}

// sBoxMonomial (poseidon/goldilocks.go:138-145): x2 = x x, x3 = x x2, Reduce192, x6 = x3r x3r, x7 = x x6, Reduce192
template <class Emit>
GPW_HD uint64_t sbox_emit(uint64_t x, uint32_t slot, Emit& emit) {
  uint64_t lo, hi;
  gl::mul64(x, x, lo, hi);
  emit(slot, U192{{lo, hi, 0}});
  const U192 x3 = mul128x64(lo, hi, x);
  emit(slot + 1, x3);
  const uint64_t x3r = reduce_emit(x3, slot + 2, emit);
  gl::mul64(x3r, x3r, lo, hi);
  emit(slot + 4, U192{{lo, hi, 0}});
  const U192 x7 = mul128x64(lo, hi, x);
  emit(slot + 5, x7);
  return reduce_emit(x7, slot + 6, emit);
}

GPW_HD U192 mds_row(int r, const uint64_t* st, const uint64_t* T) {
  U192 acc = {{0, 0, 0}};
#pragma unroll 1
  for (int i = 0; i < 12; i++) mac(acc, st[(i + r) % 12], T[T_CIRC + i]);
  mac(acc, st[r], T[T_DIAG + r]);
  return acc;
}

GPW_HD U192 partial_init_row(int d, const uint64_t* st, const uint64_t* T) {
  if (d == 0) return u192(st[0]);
  U192 acc = {{0, 0, 0}};
#pragma unroll 1
  for (int r = 1; r < 12; r++) mac(acc, st[r], T[T_INIT + (r - 1) * 11 + (d - 1)]);
  return acc;
}

GPW_HD uint32_t full_base(int f) { return f < 4 ? (uint32_t)f * FULL_SLOTS : PARTIAL_BASE + PARTIAL_SLOTS + (uint32_t)(f - 4) * FULL_SLOTS; }
GPW_HD int full_rc(int f) { return f < 4 ? f : f + 22; }

// Sequential form. st: the 12 canonical input values (< p); emit(slot, U192). Leaves the output state in st.
template <class Emit>
GPW_HD void trace_seq(uint64_t st[12], const uint64_t* T, Emit emit) {
  auto full_round = [&](int f) {
    const uint32_t B = full_base(f);
    for (int k = 0; k < 12; k++) st[k] = addconst_emit(st[k], T[T_RC + k + 12 * full_rc(f)], B + 2 * k, emit);
    for (int k = 0; k < 12; k++) st[k] = sbox_emit(st[k], B + 24 + 8 * k, emit);
    uint64_t nx[12];
    for (int r = 0; r < 12; r++) nx[r] = reduce_emit(mds_row(r, st, T), B + 120 + 2 * r, emit);
    for (int r = 0; r < 12; r++) st[r] = nx[r];
  };
  for (int f = 0; f < 4; f++) full_round(f);
  {
    const uint32_t B = PARTIAL_BASE;
    for (int k = 0; k < 12; k++) st[k] = addconst_emit(st[k], T[T_FIRST + k], B + 2 * k, emit);
    uint64_t nx[12];
    for (int d = 0; d < 12; d++) nx[d] = reduce_emit(partial_init_row(d, st, T), B + 24 + 2 * d, emit);
    for (int d = 0; d < 12; d++) st[d] = nx[d];
    for (int i = 0; i < 22; i++) {
      const uint32_t Bi = B + 48 + 36 * i;
      uint64_t s0 = sbox_emit(st[0], Bi, emit);
      s0 = addconst_emit(s0, T[T_PRC + i], Bi + 8, emit);
      U192 d = {{0, 0, 0}};
      for (int j = 1; j < 12; j++) mac(d, st[j], T[T_WHATS + i * 11 + j - 1]);
      mac(d, s0, MDS0TO0);
      const uint64_t dr = reduce_emit(d, Bi + 10, emit);
      nx[0] = reduce_emit(u192(dr), Bi + 12, emit);
      for (int j = 1; j < 12; j++) {
        U192 v = u192(st[j]);
        mac(v, s0, T[T_VS + i * 11 + j - 1]);
        nx[j] = reduce_emit(v, Bi + 12 + 2 * j, emit);
      }
      for (int j = 0; j < 12; j++) st[j] = nx[j];
    }
  }
  for (int f = 4; f < 8; f++) full_round(f);
}

#ifdef __CUDACC__
// Warp form: lane k < 12 owns state element k in a register; the MDS layers mix the state with WARP SHUFFLES (the circulant
// row of lane r reads lane (i + r) mod 12 for i = 0..11; the partial rounds broadcast lanes 1..11 to lane 0 and s0 back),
// no shared-memory exchange and no __syncwarp. Called by all 32 lanes of one warp (lanes >= 12 only take part in the
// shuffles). emit(slot, U192) must be callable concurrently from the lanes. `sh` is unused (kept for the callers' layout).
__device__ __forceinline__ uint64_t shfl64(uint64_t v, int src) { return (uint64_t)__shfl_sync(0xffffffffu, (unsigned long long)v, src); }

template <class Emit>
__device__ void trace_warp(uint64_t x, uint64_t* sh, const uint64_t* __restrict__ T, Emit emit) {
  (void)sh;
  const int lane = (int)(threadIdx.x & 31u);
  const bool on = lane < 12;
  const int r = on ? lane : 0;
  // sum_i st[(i + r) % 12] C[i] + st[r] D[r] with the state spread over lanes 0..11
  auto mds_row_shfl = [&](uint64_t own) {
    U192 acc = {{0, 0, 0}};
    uint64_t v[12];
#pragma unroll
    for (int i = 0; i < 12; i++) {  // twelve independent shuffles in flight, then the multiply-accumulate chain
      int src = i + r;
      src = src >= 12 ? src - 12 : src;
      v[i] = shfl64(own, src);
    }
#pragma unroll
    for (int i = 0; i < 12; i++) mac(acc, v[i], T[T_CIRC + i]);
    mac(acc, own, T[T_DIAG + r]);
    return acc;
  };
  auto full_round = [&](int f) {
    const uint32_t B = full_base(f);
    if (on) {
      x = addconst_emit(x, T[T_RC + lane + 12 * full_rc(f)], B + 2 * lane, emit);
      x = sbox_emit(x, B + 24 + 8 * lane, emit);
    }
    const U192 row = mds_row_shfl(x);
    if (on) x = reduce_emit(row, B + 120 + 2 * lane, emit);
  };
#pragma unroll 1
  for (int f = 0; f < 4; f++) full_round(f);
  {
    const uint32_t B = PARTIAL_BASE;
    if (on) x = addconst_emit(x, T[T_FIRST + lane], B + 2 * lane, emit);
    {  // partial_init_row(lane): lane 0 keeps st[0], lane d sums st[r] INIT[(r - 1) 11 + d - 1] over r = 1..11
      U192 acc = {{0, 0, 0}};
#pragma unroll 1
      for (int rr = 1; rr < 12; rr++) {
        const uint64_t v = shfl64(x, rr);
        if (on && lane > 0) mac(acc, v, T[T_INIT + (rr - 1) * 11 + (lane - 1)]);
      }
      if (on) x = reduce_emit(lane == 0 ? u192(x) : acc, B + 24 + 2 * lane, emit);
    }
#pragma unroll 1
    for (int i = 0; i < 22; i++) {
      const uint32_t Bi = B + 48 + 36 * i;
      U192 d = {{0, 0, 0}};
      uint64_t vs[11];
#pragma unroll
      for (int j = 1; j < 12; j++) vs[j - 1] = shfl64(x, j);
      if (lane == 0) {
#pragma unroll
        for (int j = 1; j < 12; j++) mac(d, vs[j - 1], T[T_WHATS + i * 11 + j - 1]);
      }
      uint64_t s0 = 0;
      if (lane == 0) {
        s0 = sbox_emit(x, Bi, emit);
        s0 = addconst_emit(s0, T[T_PRC + i], Bi + 8, emit);
        mac(d, s0, MDS0TO0);
        x = reduce_emit(d, Bi + 10, emit);
      }
      s0 = shfl64(s0, 0);
      if (on) {
        U192 v = u192(x);  // lane 0: d (already reduced: q = 0)
        if (lane > 0) mac(v, s0, T[T_VS + i * 11 + lane - 1]);
        x = reduce_emit(v, Bi + 12 + 2 * lane, emit);
      }
    }
  }
#pragma unroll 1
  for (int f = 4; f < 8; f++) full_round(f);
}
#endif

}  // namespace glm
}  // namespace gpw
