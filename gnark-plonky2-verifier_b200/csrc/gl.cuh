// Goldilocks field GF(p), p = 2^64 - 2^32 + 1, and the exact integer (quotient, remainder) arithmetic
// of the reference's four solver hints (goldilocks/base.go:223-359), host + device.
//
// The hints are NOT field operations: they return the integer quotient as well as the remainder, and
// the quotient is a witness wire. q = floor(x / p) is obtained without any division: r = x mod p by
// the 2^64 = 2^32 - 1 folding, then q = (x - r) * p^-1 mod 2^256 (x - r is an exact multiple of the
// odd p, so multiplying by the 2-adic inverse of p is exact division).
#pragma once
#include <cstdint>

#include "ff.cuh"

namespace gpw {
namespace gl {

constexpr uint64_t P = 0xffffffff00000001ull;
constexpr uint64_t EPS = 0xffffffffull;  // 2^64 mod p

GPW_HD void mul64(uint64_t a, uint64_t b, uint64_t& lo, uint64_t& hi) {
#ifdef __CUDA_ARCH__
  lo = a * b;
  hi = __umul64hi(a, b);
#else
  unsigned __int128 t = (unsigned __int128)a * b;
  lo = (uint64_t)t;
  hi = (uint64_t)(t >> 64);
#endif
}

// (hi * 2^64 + lo) mod p, canonical
GPW_HD uint64_t reduce128(uint64_t lo, uint64_t hi) {
  uint64_t hi_hi = hi >> 32, hi_lo = hi & EPS;
  uint64_t t0 = lo - hi_hi;
  if (lo < hi_hi) t0 -= EPS;
  uint64_t t1 = hi_lo * EPS;
  uint64_t res = t0 + t1;
  if (res < t1) res += EPS;
  if (res >= P) res -= P;
  return res;
}

GPW_HD uint64_t add(uint64_t a, uint64_t b) {  // a, b canonical
  uint64_t s = a + b;
  if (s < a || s >= P) s -= P;
  return s;
}
GPW_HD uint64_t sub(uint64_t a, uint64_t b) { return a >= b ? a - b : a + (P - b); }
GPW_HD uint64_t mul(uint64_t a, uint64_t b) {
  uint64_t lo, hi;
  mul64(a, b, lo, hi);
  return reduce128(lo, hi);
}
GPW_HD uint64_t pow(uint64_t a, uint64_t e) {
  uint64_t r = 1;
  while (e) {
    if (e & 1) r = mul(r, a);
    a = mul(a, a);
    e >>= 1;
  }
  return r;
}
GPW_HD uint64_t inverse(uint64_t a) { return a ? pow(a, P - 2) : 0; }  // gnark-crypto: Inverse(0) = 0

// ---- exact hints ------------------------------------------------------------------------------
// MulAddHint (base.go:223-243): q = floor((a b + c) / p), r = (a b + c) mod p; a, b, c < p so q < 2^64
GPW_HD void mul_add_hint(uint64_t a, uint64_t b, uint64_t c, uint64_t& q, uint64_t& r) {
  uint64_t lo, hi;
  mul64(a, b, lo, hi);
  uint64_t lo2 = lo + c;
  hi += (lo2 < lo) ? 1u : 0u;
  r = reduce128(lo2, hi);
  q = (lo2 - r) * 0x100000001ull;  // p^-1 mod 2^64 = 2^32 + 1
}

// ReduceHint (base.go:284-294) for x < 2^256 given as 4 LE u64: q (4 LE u64), r
GPW_HD void reduce_hint(const uint64_t x[4], uint64_t q[4], uint64_t& r) {
  uint64_t acc = reduce128(x[3], 0);
  acc = reduce128(x[2], acc);
  acc = reduce128(x[1], acc);
  acc = reduce128(x[0], acc);
  r = acc;
  // d = x - r
  uint64_t d[4];
  uint64_t borrow = x[0] < r ? 1u : 0u;
  d[0] = x[0] - r;
  for (int i = 1; i < 4; i++) {
    uint64_t t = x[i] - borrow;
    borrow = (x[i] < borrow) ? 1u : 0u;
    d[i] = t;
  }
  // q = d * PINV mod 2^256
  const uint64_t PINV[4] = {0x100000001ull, 0xffffffff00000000ull, 0xfffffffffffffffeull, 0x100000000ull};
  uint64_t out[4] = {0, 0, 0, 0};
  for (int i = 0; i < 4; i++) {
    uint64_t carry = 0;
    for (int j = 0; i + j < 4; j++) {
      uint64_t lo, hi;
      mul64(d[i], PINV[j], lo, hi);
      uint64_t s = out[i + j] + lo;
      uint64_t c1 = s < lo ? 1u : 0u;
      uint64_t s2 = s + carry;
      uint64_t c2 = s2 < s ? 1u : 0u;
      out[i + j] = s2;
      carry = hi + c1 + c2;
    }
  }
  for (int i = 0; i < 4; i++) q[i] = out[i];
}

// SplitLimbsHint (base.go:339-359)
GPW_HD void split_limbs_hint(uint64_t x, uint64_t& hi, uint64_t& lo) {
  hi = x >> 32;
  lo = x & EPS;
}

}  // namespace gl
}  // namespace gpw
