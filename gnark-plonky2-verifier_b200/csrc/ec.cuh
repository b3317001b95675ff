// BN254 G1 (over Fp) and G2 (over Fp2) group arithmetic, host + device.
//
// Curves: E: y^2 = x^3 + 3 and the twist E': y^2 = x^3 + 3/(9+u) - both have a = 0, so one formula
// set templated on the coordinate field serves both.
// Bucket accumulators use extended Jacobian "XYZZ" coordinates (x = X/ZZ, y = Y/ZZZ, ZZ^3 = ZZZ^2):
// mixed addition of an affine point costs 8M + 2S with no inversion, which is what Pippenger's
// bucket accumulation does almost exclusively. Memory layout of affine points matches gnark-crypto's
// G1Affine{X,Y fp.Element} (64 B) / G2Affine{X,Y E2{A0,A1}} (128 B), Montgomery form.
#pragma once
#include "ff.cuh"

// Group operations that are NOT on the bucket-accumulation hot path are real function calls
// (__noinline__): their bodies are 10-40 field multiplies, and inlining them at every call site of the
// window-reduction kernels explodes both compile time and register pressure for no measurable gain.
#ifdef __CUDACC__
#define GPW_HD_CALL __host__ __device__ __noinline__
#else
#define GPW_HD_CALL
#endif

namespace gpw {

template <class F>
struct alignas(16) Affine {
  F x, y;  // (0,0) encodes the point at infinity (gnark-crypto convention)
  GPW_HD bool is_inf() const { return x.is_zero() && y.is_zero(); }
};

template <class F>
struct alignas(16) XYZZ {
  F X, Y, ZZ, ZZZ;  // ZZ == 0 encodes infinity
  static GPW_HD XYZZ inf() { return {F::zero(), F::zero(), F::zero(), F::zero()}; }
  GPW_HD bool is_inf() const { return ZZ.is_zero(); }
  static GPW_HD XYZZ from_affine(const Affine<F>& p) {
    if (p.is_inf()) return inf();
    return {p.x, p.y, F::one(), F::one()};
  }
};

// 2*P for affine P (EFD mdbl-2008-s-1, a = 0)
template <class F>
GPW_HD_CALL XYZZ<F> dbl_affine(const Affine<F>& p) {
  if (p.is_inf()) return XYZZ<F>::inf();
  F U = dbl(p.y);
  F V = sqr(U);
  F W = mul(U, V);
  F S = mul(p.x, V);
  F X2 = sqr(p.x);
  F M = add(dbl(X2), X2);
  F X3 = sub(sqr(M), dbl(S));
  F Y3 = mul_sub2(M, sub(S, X3), W, p.y);
  return {X3, Y3, V, W};
}

// 2*P (EFD dbl-2008-s-1, a = 0)
template <class F>
GPW_HD_CALL XYZZ<F> dbl(const XYZZ<F>& p) {
  if (p.is_inf()) return p;
  F U = dbl(p.Y);
  F V = sqr(U);
  F W = mul(U, V);
  F S = mul(p.X, V);
  F X2 = sqr(p.X);
  F M = add(dbl(X2), X2);
  F X3 = sub(sqr(M), dbl(S));
  F Y3 = mul_sub2(M, sub(S, X3), W, p.Y);
  return {X3, Y3, mul(V, p.ZZ), mul(W, p.ZZZ)};
}

// acc += (x2, sign ? -y2 : y2)   (EFD madd-2008-s). Handles infinity, doubling and cancellation.
template <class F>
GPW_HD void add_mixed(XYZZ<F>& acc, const Affine<F>& q, bool negate) {
  if (q.is_inf()) return;
  F y2 = negate ? neg(q.y) : q.y;
  if (acc.is_inf()) {
    acc = {q.x, y2, F::one(), F::one()};
    return;
  }
  F U2 = mul(q.x, acc.ZZ);
  F S2 = mul(y2, acc.ZZZ);
  F Pv = sub(U2, acc.X);
  F R = sub(S2, acc.Y);
  if (Pv.is_zero()) {
    if (R.is_zero()) {
      acc = dbl_affine(Affine<F>{q.x, y2});
    } else {
      acc = XYZZ<F>::inf();
    }
    return;
  }
  F PP = sqr(Pv);
  F PPP = mul(Pv, PP);
  F Q = mul(acc.X, PP);
  F X3 = sub(sub(sqr(R), PPP), dbl(Q));
  F Y3 = mul_sub2(R, sub(Q, X3), acc.Y, PPP);
  acc.X = X3;
  acc.Y = Y3;
  acc.ZZ = mul(acc.ZZ, PP);
  acc.ZZZ = mul(acc.ZZZ, PPP);
}

// acc += q  (EFD add-2008-s)
template <class F>
GPW_HD_CALL void add_full(XYZZ<F>& acc, const XYZZ<F>& q) {
  if (q.is_inf()) return;
  if (acc.is_inf()) {
    acc = q;
    return;
  }
  F U1 = mul(acc.X, q.ZZ);
  F U2 = mul(q.X, acc.ZZ);
  F S1 = mul(acc.Y, q.ZZZ);
  F S2 = mul(q.Y, acc.ZZZ);
  F Pv = sub(U2, U1);
  F R = sub(S2, S1);
  if (Pv.is_zero()) {
    if (R.is_zero()) {
      acc = dbl(acc);
    } else {
      acc = XYZZ<F>::inf();
    }
    return;
  }
  F PP = sqr(Pv);
  F PPP = mul(Pv, PP);
  F Q = mul(U1, PP);
  F X3 = sub(sub(sqr(R), PPP), dbl(Q));
  F Y3 = mul_sub2(R, sub(Q, X3), S1, PPP);
  acc.X = X3;
  acc.Y = Y3;
  acc.ZZ = mul(mul(acc.ZZ, q.ZZ), PP);
  acc.ZZZ = mul(mul(acc.ZZZ, q.ZZZ), PPP);
}

template <class F>
GPW_HD XYZZ<F> neg(const XYZZ<F>& p) {
  return {p.X, neg(p.Y), p.ZZ, p.ZZZ};
}

// XYZZ -> affine (one field inversion; off the hot path)
template <class F>
GPW_HD_CALL Affine<F> to_affine(const XYZZ<F>& p) {
  if (p.is_inf()) return {F::zero(), F::zero()};
  // i3 = 1/ZZZ;  ZZ^3 = ZZZ^2  =>  1/ZZ = ZZ^2 / ZZZ^2 = (ZZ * i3)^2
  F i3 = inv(p.ZZZ);
  F i2 = sqr(mul(i3, p.ZZ));
  return {mul(p.X, i2), mul(p.Y, i3)};
}

// scalar multiplication by a small unsigned integer (double-and-add, MSB first); window-reduction glue
template <class F>
GPW_HD_CALL XYZZ<F> mul_small(const XYZZ<F>& p, uint32_t k) {
  XYZZ<F> r = XYZZ<F>::inf();
  for (int b = 31; b >= 0; b--) {
    r = dbl(r);
    if ((k >> b) & 1u) add_full(r, p);
  }
  return r;
}

using G1Affine = Affine<Fp>;
using G1XYZZ = XYZZ<Fp>;
using G2Affine = Affine<Fp2>;
using G2XYZZ = XYZZ<Fp2>;

}  // namespace gpw
