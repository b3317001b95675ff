// Native evaluation of one Poseidon-BN254 permutation that EMITS, in the order the circuit builder creates them
// (csrc/host/gadgets_core.cc PoseidonBn254Chip::Poseidon, mirroring poseidon/bn254.go:39-208), the value of every
// multiplication wire of the gadget: x^2, x^4, x^5 for each S-box. Used by the solver's OP_POSEIDON_BN254 macro
// instruction (host test interpreter and CUDA executor share this code).
#pragma once
#include "ff.cuh"

namespace gpw {

// tables: C[88], S[392], M[16] (row-major m[j][i] at j*4+i), P[16]; all Fr in Montgomery form.
struct Bn254PoseidonTables {
  const Fr* C;
  const Fr* S;
  const Fr* M;
  const Fr* P;
};

template <class Emit>
GPW_HD void bn254_exp5_emit(Fr& x, bool emit_wires, Emit& emit) {
  Fr x2 = sqr_chain(x);
  Fr x4 = sqr_chain(x2);
  Fr x5 = mul(x4, x);
  if (emit_wires) {
    emit(x2);
    emit(x4);
    emit(x5);
  }
  x = x5;
}

GPW_HD void bn254_mix(Fr st[4], const Fr* m) {
  Fr out[4];
#pragma unroll
  for (int i = 0; i < 4; i++) {
    Fr acc = mul(m[0 * 4 + i], st[0]);
#pragma unroll
    for (int j = 1; j < 4; j++) acc = add(acc, mul(m[j * 4 + i], st[j]));
    out[i] = acc;
  }
#pragma unroll
  for (int i = 0; i < 4; i++) st[i] = out[i];
}

// st: the 4 input state values (Montgomery). input_is_const[k]: the builder folded input k to a constant, so the
// round-0 S-box of lane k creates no wires. Returns the output state in st.
template <class Emit>
GPW_HD void poseidon_bn254_trace(Fr st[4], const bool input_is_const[4], const Bn254PoseidonTables& T, Emit emit) {
#pragma unroll
  for (int k = 0; k < 4; k++) st[k] = add(st[k], T.C[k]);
  // first half of the full rounds
#pragma unroll 1
  for (int i = 0; i < 4; i++) {
#pragma unroll
    for (int k = 0; k < 4; k++) bn254_exp5_emit(st[k], !(i == 0 && input_is_const[k]), emit);
#pragma unroll
    for (int k = 0; k < 4; k++) st[k] = add(st[k], T.C[(i + 1) * 4 + k]);
    bn254_mix(st, i < 3 ? T.M : T.P);
  }
  // partial rounds
#pragma unroll 1
  for (int i = 0; i < 56; i++) {
    bn254_exp5_emit(st[0], true, emit);
    st[0] = add(st[0], T.C[20 + i]);
    Fr n0 = mul(T.S[7 * i], st[0]);
#pragma unroll
    for (int j = 1; j < 4; j++) n0 = add(n0, mul(T.S[7 * i + j], st[j]));
#pragma unroll
    for (int k = 1; k < 4; k++) st[k] = add(st[k], mul(st[0], T.S[7 * i + 4 + k - 1]));
    st[0] = n0;
  }
  // second half of the full rounds
#pragma unroll 1
  for (int i = 0; i < 4; i++) {
#pragma unroll
    for (int k = 0; k < 4; k++) bn254_exp5_emit(st[k], true, emit);
    if (i < 3) {
#pragma unroll
      for (int k = 0; k < 4; k++) st[k] = add(st[k], T.C[20 + 56 + i * 4 + k]);
    }
    bn254_mix(st, T.M);
  }
}

#ifdef __CUDACC__
// Four-lane form of the same permutation for the solver's grid kernel (k_tape_poseidon4): lane q = 0..3 of an aligned group
// of four lanes owns state element q. One thread per permutation walks 784 dependent Montgomery multiplications (a full
// round is 12 S-box + 16 matrix products, a partial round 3 + 7); spread over the four lanes a full round is 3 + 4 and a
// partial round 3 + 1 + 1 multiplication times (S-box; one product per lane of the sparse row, summed by two butterfly
// shuffles; one product per lane of the column update): 336 instead of 784 on the critical path. The warp executes every
// multiplication for all of its lanes anyway, so the split costs no issue slots. Field additions are exact, so summing the
// row in a different order gives the same canonical value; the emitted wires and their order are those of
// poseidon_bn254_trace. Must be called by all 32 lanes of a warp (groups without work pass valid = false).
__device__ __forceinline__ Fr shfl_fr(const Fr& v, int src_lane) {
  Fr r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.l[i] = __shfl_sync(0xffffffffu, v.l[i], src_lane);
  return r;
}
__device__ __forceinline__ Fr shfl_xor_fr(const Fr& v, int m) {
  Fr r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.l[i] = __shfl_xor_sync(0xffffffffu, v.l[i], m);
  return r;
}

// x: this lane's state element; const_mask: bit k set = input k was folded to a constant by the builder (no wires for its
// round-0 S-box). emit(idx, v) is called by the lane owning wire idx (0 .. 264 - 3 popc(const_mask)).
template <class Emit>
__device__ __forceinline__ void poseidon_bn254_trace4(Fr& x, uint32_t const_mask, const Bn254PoseidonTables& T, Emit emit) {
  const int lane = (int)(threadIdx.x & 31u), q = lane & 3, g0 = lane & ~3;
  const uint32_t nskip = (uint32_t)__popc(const_mask & 0xfu);
  const uint32_t rank0 = (uint32_t)__popc(~const_mask & ((1u << q) - 1u) & 0xfu);  // emitting lanes before this one, round 0
  const bool skip0 = (const_mask >> q) & 1u;
  auto sbox = [&](uint32_t idx, bool do_emit) {
    const Fr x2 = sqr_chain(x);
    const Fr x4 = sqr_chain(x2);
    const Fr x5 = mul(x4, x);
    if (do_emit) {
      emit(idx, x2);
      emit(idx + 1, x4);
      emit(idx + 2, x5);
    }
    return x5;
  };
  // out[q] = sum_j m[j * 4 + q] st[j]
  auto mix = [&](const Fr* m) {
    Fr acc = mul(m[q], shfl_fr(x, g0));
#pragma unroll
    for (int j = 1; j < 4; j++) acc = add(acc, mul(m[j * 4 + q], shfl_fr(x, g0 + j)));
    x = acc;
  };
  x = add(x, T.C[q]);
  uint32_t base = 0;
#pragma unroll 1
  for (int i = 0; i < 4; i++) {
    if (i == 0) {
      x = sbox(3 * rank0, !skip0);
      base = 3 * (4 - nskip);
    } else {
      x = sbox(base + 3 * q, true);
      base += 12;
    }
    x = add(x, T.C[(i + 1) * 4 + q]);
    mix(i < 3 ? T.M : T.P);
  }
#pragma unroll 1
  for (int i = 0; i < 56; i++) {
    const Fr s5 = sbox(base, q == 0);  // lanes 1..3 run the same instructions on their own element and drop the result
    base += 3;
    if (q == 0) x = add(s5, T.C[20 + i]);
    const Fr s0 = shfl_fr(x, g0);
    // sparse row: n0 = S[7i] st0 + sum_j S[7i + j] st[j] - one product per lane, then a butterfly sum
    Fr p = mul(T.S[7 * i + q], x);
    p = add(p, shfl_xor_fr(p, 1));
    p = add(p, shfl_xor_fr(p, 2));
    // column: st[k] += st0 S[7i + 4 + k - 1]
    const Fr u = mul(s0, T.S[7 * i + 3 + (q == 0 ? 1 : q)]);
    x = q == 0 ? p : add(x, u);
  }
#pragma unroll 1
  for (int i = 0; i < 4; i++) {
    x = sbox(base + 3 * q, true);
    base += 12;
    if (i < 3) x = add(x, T.C[20 + 56 + i * 4 + q]);
    mix(T.M);
  }
}
#endif


}  // namespace gpw
