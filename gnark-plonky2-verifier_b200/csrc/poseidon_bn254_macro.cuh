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
  Fr x2 = sqr(x);
  Fr x4 = sqr(x2);
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

}  // namespace gpw
