// Host-side group helpers shared by the C-ABI glue (setup, final proof assembly, tests). Off the hot path.
#pragma once
#include <vector>

#include "ec.cuh"

namespace gpw {

template <class F>
inline XYZZ<F> host_scalar_mul(const Affine<F>& p, const uint32_t k[8]) {
  XYZZ<F> r = XYZZ<F>::inf();
  for (int w = 7; w >= 0; w--)
    for (int b = 31; b >= 0; b--) {
      r = dbl(r);
      if ((k[w] >> b) & 1u) add_mixed(r, p, false);
    }
  return r;
}

inline Fp fp_from_u64(uint64_t v) {
  Fp a = Fp::zero();
  a.l[0] = (uint32_t)v;
  a.l[1] = (uint32_t)(v >> 32);
  return to_mont(a);
}

// b' = 3 / (9 + u)
inline Fp2 g2_b() {
  Fp2 d{fp_from_u64(9), fp_from_u64(1)};
  Fp2 i = inv(d);
  Fp three = fp_from_u64(3);
  return {mul(i.c0, three), mul(i.c1, three)};
}

// decimal string -> Fp (Montgomery). Only used for the G2 generator constants below.
inline Fp fp_from_dec(const char* s) {
  Fp acc = Fp::zero();
  Fp ten = fp_from_u64(10);
  for (; *s; s++) acc = add(mul(acc, ten), fp_from_u64((uint64_t)(*s - '0')));
  return acc;
}

template <class F>
inline Affine<F> generator();
template <>
inline Affine<Fp> generator<Fp>() {
  return {fp_from_u64(1), fp_from_u64(2)};
}
template <>
inline Affine<Fp2> generator<Fp2>() {
  return {{fp_from_dec("10857046999023057135944570762232829481370756359578518086990519993285655852781"),
           fp_from_dec("11559732032986387107991004021392285783925812861821192530917403151452391805634")},
          {fp_from_dec("8495653923123431417604973247489272438418190587263600148770280649306958101930"),
           fp_from_dec("4082367875863433681332203403145435568316851327593401208105741076214120093531")}};
}

// batch to-affine with Montgomery's trick (one inversion per call)
template <class F>
inline void batch_to_affine(std::vector<XYZZ<F>>& pts, Affine<F>* out) {
  size_t n = pts.size();
  std::vector<F> pref(n);
  F run = F::one();
  for (size_t i = 0; i < n; i++) {
    pref[i] = run;
    if (!pts[i].is_inf()) run = mul(run, pts[i].ZZZ);
  }
  F invrun = inv(run);
  for (size_t i = n; i-- > 0;) {
    if (pts[i].is_inf()) {
      out[i] = {F::zero(), F::zero()};
      continue;
    }
    F i3 = mul(invrun, pref[i]);
    invrun = mul(invrun, pts[i].ZZZ);
    F i2 = sqr(mul(i3, pts[i].ZZ));
    out[i] = {mul(pts[i].X, i2), mul(pts[i].Y, i3)};
  }
}


inline Fr fr_from_u64_host(uint64_t v) {
  Fr a = Fr::zero();
  a.l[0] = (uint32_t)v;
  a.l[1] = (uint32_t)(v >> 32);
  return to_mont(a);
}

}  // namespace gpw
