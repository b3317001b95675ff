// BN254 prime-field arithmetic (Fr = scalar field, Fp = base field) for host and sm_100a device.
//
// Representation: 8 x 32-bit little-endian limbs, Montgomery form with R = 2^256 - bit-identical in
// memory to gnark-crypto's fr.Element / fp.Element ([4]uint64 LE limbs, Montgomery), so
// buffers cross the C ABI without conversion.
//
// Device multiply: CIOS Montgomery on the 32-bit IMAD pipe. Partial products are accumulated in two
// interleaved 8-word accumulators ("even"/"odd" 64-bit columns) so that every mad.lo.cc/madc.hi.cc
// pair lowers to one IMAD.WIDE.U32(.X) with the carry kept in the CC/predicate chain - no tensor
// cores (pure modular-integer work, BASELINE.json north_star). The same even/odd schedule is
// emulated on the host (explicit carry) so the algorithm is unit-tested on CPU against a plain
// 64-bit CIOS; the GPU self-test compares the PTX path with both.
#pragma once
#include <cstdint>
#include <cstring>

#ifdef __CUDACC__
#define GPW_HD __host__ __device__ __forceinline__
#define GPW_D __device__ __forceinline__
#else
#define GPW_HD inline
#define GPW_D inline
#endif

namespace gpw {

// ---------------------------------------------------------------------------------------------
// Field parameters (numerically re-derived in tests/test_host_ff.py)
// ---------------------------------------------------------------------------------------------
struct FrParams {
  // r = 0x30644e72e131a029b85045b68181585d2833e84879b9709143e1f593f0000001
  static constexpr GPW_HD uint32_t mod(int i) {
    constexpr uint32_t m[8] = {0xf0000001u, 0x43e1f593u, 0x79b97091u, 0x2833e848u,
                               0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u};
    return m[i];
  }
  // R mod r
  static constexpr GPW_HD uint32_t one(int i) {
    constexpr uint32_t m[8] = {0x4ffffffbu, 0xac96341cu, 0x9f60cd29u, 0x36fc7695u,
                               0x7879462eu, 0x666ea36fu, 0x9a07df2fu, 0x0e0a77c1u};
    return m[i];
  }
  // R^2 mod r
  static constexpr GPW_HD uint32_t r2(int i) {
    constexpr uint32_t m[8] = {0xae216da7u, 0x1bb8e645u, 0xe35c59e3u, 0x53fe3ab1u,
                               0x53bb8085u, 0x8c49833du, 0x7f4e44a5u, 0x0216d0b1u};
    return m[i];
  }
  static constexpr uint32_t M0 = 0xefffffffu;  // -r^{-1} mod 2^32
};

struct FpParams {
  // p = 0x30644e72e131a029b85045b68181585d97816a916871ca8d3c208c16d87cfd47
  static constexpr GPW_HD uint32_t mod(int i) {
    constexpr uint32_t m[8] = {0xd87cfd47u, 0x3c208c16u, 0x6871ca8du, 0x97816a91u,
                               0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u};
    return m[i];
  }
  static constexpr GPW_HD uint32_t one(int i) {
    constexpr uint32_t m[8] = {0xc58f0d9du, 0xd35d438du, 0xf5c70b3du, 0x0a78eb28u,
                               0x7879462cu, 0x666ea36fu, 0x9a07df2fu, 0x0e0a77c1u};
    return m[i];
  }
  static constexpr GPW_HD uint32_t r2(int i) {
    constexpr uint32_t m[8] = {0x538afa89u, 0xf32cfc5bu, 0xd44501fbu, 0xb5e71911u,
                               0x0a417ff6u, 0x47ab1effu, 0xcab8351fu, 0x06d89f71u};
    return m[i];
  }
  static constexpr uint32_t M0 = 0xe4866389u;  // -p^{-1} mod 2^32
};

// ---------------------------------------------------------------------------------------------
template <class P>
struct alignas(16) Fe {
  uint32_t l[8];

  static GPW_HD Fe zero() {
    Fe r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.l[i] = 0;
    return r;
  }
  static GPW_HD Fe one() {
    Fe r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.l[i] = P::one(i);
    return r;
  }
  static GPW_HD Fe r2() {
    Fe r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.l[i] = P::r2(i);
    return r;
  }
  GPW_HD bool is_zero() const {
    uint32_t o = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) o |= l[i];
    return o == 0;
  }
  GPW_HD bool operator==(const Fe& b) const {
    uint32_t o = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) o |= (l[i] ^ b.l[i]);
    return o == 0;
  }
  GPW_HD bool operator!=(const Fe& b) const { return !(*this == b); }
};

// ---- raw 256-bit helpers --------------------------------------------------------------------
// r = a + b, returns carry
template <class P>
GPW_HD uint32_t add_raw(Fe<P>& r, const Fe<P>& a, const Fe<P>& b) {
#ifdef __CUDA_ARCH__
  uint32_t c;
  asm("add.cc.u32 %0, %9, %17;\n\t"
      "addc.cc.u32 %1, %10, %18;\n\t"
      "addc.cc.u32 %2, %11, %19;\n\t"
      "addc.cc.u32 %3, %12, %20;\n\t"
      "addc.cc.u32 %4, %13, %21;\n\t"
      "addc.cc.u32 %5, %14, %22;\n\t"
      "addc.cc.u32 %6, %15, %23;\n\t"
      "addc.cc.u32 %7, %16, %24;\n\t"
      "addc.u32 %8, 0, 0;"
      : "=r"(r.l[0]), "=r"(r.l[1]), "=r"(r.l[2]), "=r"(r.l[3]), "=r"(r.l[4]), "=r"(r.l[5]), "=r"(r.l[6]),
        "=r"(r.l[7]), "=r"(c)
      : "r"(a.l[0]), "r"(a.l[1]), "r"(a.l[2]), "r"(a.l[3]), "r"(a.l[4]), "r"(a.l[5]), "r"(a.l[6]), "r"(a.l[7]),
        "r"(b.l[0]), "r"(b.l[1]), "r"(b.l[2]), "r"(b.l[3]), "r"(b.l[4]), "r"(b.l[5]), "r"(b.l[6]), "r"(b.l[7]));
  return c;
#else
  uint64_t c = 0;
  for (int i = 0; i < 8; i++) {
    uint64_t s = (uint64_t)a.l[i] + b.l[i] + c;
    r.l[i] = (uint32_t)s;
    c = s >> 32;
  }
  return (uint32_t)c;
#endif
}

// r = a - b, returns borrow (1 if a < b)
template <class P>
GPW_HD uint32_t sub_raw(Fe<P>& r, const Fe<P>& a, const Fe<P>& b) {
#ifdef __CUDA_ARCH__
  uint32_t c;
  asm("sub.cc.u32 %0, %9, %17;\n\t"
      "subc.cc.u32 %1, %10, %18;\n\t"
      "subc.cc.u32 %2, %11, %19;\n\t"
      "subc.cc.u32 %3, %12, %20;\n\t"
      "subc.cc.u32 %4, %13, %21;\n\t"
      "subc.cc.u32 %5, %14, %22;\n\t"
      "subc.cc.u32 %6, %15, %23;\n\t"
      "subc.cc.u32 %7, %16, %24;\n\t"
      "subc.u32 %8, 0, 0;"
      : "=r"(r.l[0]), "=r"(r.l[1]), "=r"(r.l[2]), "=r"(r.l[3]), "=r"(r.l[4]), "=r"(r.l[5]), "=r"(r.l[6]),
        "=r"(r.l[7]), "=r"(c)
      : "r"(a.l[0]), "r"(a.l[1]), "r"(a.l[2]), "r"(a.l[3]), "r"(a.l[4]), "r"(a.l[5]), "r"(a.l[6]), "r"(a.l[7]),
        "r"(b.l[0]), "r"(b.l[1]), "r"(b.l[2]), "r"(b.l[3]), "r"(b.l[4]), "r"(b.l[5]), "r"(b.l[6]), "r"(b.l[7]));
  return c & 1u;  // subc.u32 0,0 yields 0xffffffff on borrow
#else
  uint64_t brw = 0;
  for (int i = 0; i < 8; i++) {
    uint64_t d = (uint64_t)a.l[i] - b.l[i] - brw;
    r.l[i] = (uint32_t)d;
    brw = (d >> 32) & 1u;
  }
  return (uint32_t)brw;
#endif
}

template <class P>
GPW_HD Fe<P> modulus() {
  Fe<P> m;
#pragma unroll
  for (int i = 0; i < 8; i++) m.l[i] = P::mod(i);
  return m;
}

// conditional final subtraction: a in [0, 2p) -> [0, p)
template <class P>
GPW_HD Fe<P> reduce_once(const Fe<P>& a) {
  Fe<P> t;
  uint32_t borrow = sub_raw(t, a, modulus<P>());
  Fe<P> r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.l[i] = borrow ? a.l[i] : t.l[i];
  return r;
}

template <class P>
GPW_HD Fe<P> add(const Fe<P>& a, const Fe<P>& b) {
  Fe<P> s;
  add_raw(s, a, b);  // a,b < p < 2^254: no carry out
  return reduce_once(s);
}

template <class P>
GPW_HD Fe<P> sub(const Fe<P>& a, const Fe<P>& b) {
  Fe<P> d, t;
  uint32_t borrow = sub_raw(d, a, b);
  add_raw(t, d, modulus<P>());
  Fe<P> r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.l[i] = borrow ? t.l[i] : d.l[i];
  return r;
}

template <class P>
GPW_HD Fe<P> neg(const Fe<P>& a) {
  Fe<P> t;
  sub_raw(t, modulus<P>(), a);
  Fe<P> r;
  bool z = a.is_zero();
#pragma unroll
  for (int i = 0; i < 8; i++) r.l[i] = z ? 0u : t.l[i];
  return r;
}

template <class P>
GPW_HD Fe<P> dbl(const Fe<P>& a) {
  return add(a, a);
}

// ---- Montgomery multiplication: plain 64-bit CIOS (host; also the device cross-check) --------
template <class P>
GPW_HD Fe<P> mont_mul_portable(const Fe<P>& a, const Fe<P>& b) {
  uint32_t t[10];
#pragma unroll
  for (int i = 0; i < 10; i++) t[i] = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    uint64_t c = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) {
      uint64_t s = (uint64_t)a.l[j] * b.l[i] + t[j] + c;
      t[j] = (uint32_t)s;
      c = s >> 32;
    }
    uint64_t s = (uint64_t)t[8] + c;
    t[8] = (uint32_t)s;
    t[9] = (uint32_t)(s >> 32);
    uint32_t m = t[0] * P::M0;
    c = ((uint64_t)m * P::mod(0) + t[0]) >> 32;
#pragma unroll
    for (int j = 1; j < 8; j++) {
      s = (uint64_t)m * P::mod(j) + t[j] + c;
      t[j - 1] = (uint32_t)s;
      c = s >> 32;
    }
    s = (uint64_t)t[8] + c;
    t[7] = (uint32_t)s;
    t[8] = t[9] + (uint32_t)(s >> 32);
  }
  Fe<P> r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.l[i] = t[i];
  return reduce_once(r);
}

// ---- Montgomery multiplication: even/odd IMAD.WIDE schedule -----------------------------------
// Running total T = sum even[k] 2^(32k) + sum odd[k] 2^(32(k+1)).
namespace detail {

// host emulation of the PTX carry-chain primitives (explicit carry flag)
struct CC {
  uint32_t c = 0;
  GPW_HD uint32_t add_cc(uint32_t a, uint32_t b) {
    uint64_t s = (uint64_t)a + b;
    c = (uint32_t)(s >> 32);
    return (uint32_t)s;
  }
  GPW_HD uint32_t addc(uint32_t a, uint32_t b) { return a + b + c; }
  GPW_HD uint32_t mad_lo_cc(uint32_t a, uint32_t b, uint32_t d) {
    uint64_t s = (uint64_t)(uint32_t)((uint64_t)a * b) + d;
    c = (uint32_t)(s >> 32);
    return (uint32_t)s;
  }
  GPW_HD uint32_t madc_lo_cc(uint32_t a, uint32_t b, uint32_t d) {
    uint64_t s = (uint64_t)(uint32_t)((uint64_t)a * b) + d + c;
    c = (uint32_t)(s >> 32);
    return (uint32_t)s;
  }
  GPW_HD uint32_t madc_hi_cc(uint32_t a, uint32_t b, uint32_t d) {
    uint64_t s = (((uint64_t)a * b) >> 32) + d + c;
    c = (uint32_t)(s >> 32);
    return (uint32_t)s;
  }
  GPW_HD uint32_t madc_hi(uint32_t a, uint32_t b, uint32_t d) { return (uint32_t)(((uint64_t)a * b) >> 32) + d + c; }
};

// Block A (iterations 1..7): fold the pending one-limb shift, then T += a * bi.
//   e[0] += o[1]; o = (o >> 64) + a_odd * bi (carry-in from the fold); e += a_even * bi; o[7] += carry
GPW_HD void mul_acc_shift(uint32_t* e, uint32_t* o, const uint32_t* a, uint32_t bi) {
#ifdef __CUDA_ARCH__
  asm("add.cc.u32      %0,  %0,  %9;\n\t"
      "madc.lo.cc.u32  %8,  %17, %24, %10;\n\t"
      "madc.hi.cc.u32  %9,  %17, %24, %11;\n\t"
      "madc.lo.cc.u32  %10, %19, %24, %12;\n\t"
      "madc.hi.cc.u32  %11, %19, %24, %13;\n\t"
      "madc.lo.cc.u32  %12, %21, %24, %14;\n\t"
      "madc.hi.cc.u32  %13, %21, %24, %15;\n\t"
      "madc.lo.cc.u32  %14, %23, %24, 0;\n\t"
      "madc.hi.u32     %15, %23, %24, 0;\n\t"
      "mad.lo.cc.u32   %0,  %16, %24, %0;\n\t"
      "madc.hi.cc.u32  %1,  %16, %24, %1;\n\t"
      "madc.lo.cc.u32  %2,  %18, %24, %2;\n\t"
      "madc.hi.cc.u32  %3,  %18, %24, %3;\n\t"
      "madc.lo.cc.u32  %4,  %20, %24, %4;\n\t"
      "madc.hi.cc.u32  %5,  %20, %24, %5;\n\t"
      "madc.lo.cc.u32  %6,  %22, %24, %6;\n\t"
      "madc.hi.cc.u32  %7,  %22, %24, %7;\n\t"
      "addc.u32        %15, %15, 0;"
      : "+r"(e[0]), "+r"(e[1]), "+r"(e[2]), "+r"(e[3]), "+r"(e[4]), "+r"(e[5]), "+r"(e[6]), "+r"(e[7]),
        "+r"(o[0]), "+r"(o[1]), "+r"(o[2]), "+r"(o[3]), "+r"(o[4]), "+r"(o[5]), "+r"(o[6]), "+r"(o[7])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]), "r"(bi));
#else
  CC k;
  e[0] = k.add_cc(e[0], o[1]);
  uint32_t n0 = k.madc_lo_cc(a[1], bi, o[2]);
  uint32_t n1 = k.madc_hi_cc(a[1], bi, o[3]);
  uint32_t n2 = k.madc_lo_cc(a[3], bi, o[4]);
  uint32_t n3 = k.madc_hi_cc(a[3], bi, o[5]);
  uint32_t n4 = k.madc_lo_cc(a[5], bi, o[6]);
  uint32_t n5 = k.madc_hi_cc(a[5], bi, o[7]);
  uint32_t n6 = k.madc_lo_cc(a[7], bi, 0);
  uint32_t n7 = k.madc_hi(a[7], bi, 0);
  o[0] = n0; o[1] = n1; o[2] = n2; o[3] = n3; o[4] = n4; o[5] = n5; o[6] = n6; o[7] = n7;
  e[0] = k.mad_lo_cc(a[0], bi, e[0]);
  e[1] = k.madc_hi_cc(a[0], bi, e[1]);
  e[2] = k.madc_lo_cc(a[2], bi, e[2]);
  e[3] = k.madc_hi_cc(a[2], bi, e[3]);
  e[4] = k.madc_lo_cc(a[4], bi, e[4]);
  e[5] = k.madc_hi_cc(a[4], bi, e[5]);
  e[6] = k.madc_lo_cc(a[6], bi, e[6]);
  e[7] = k.madc_hi_cc(a[6], bi, e[7]);
  o[7] = k.addc(o[7], 0);
#endif
}

// Block B: T += mi * p  (makes e[0] == 0 mod 2^32)
template <class P>
GPW_HD void redc_step(uint32_t* e, uint32_t* o, uint32_t mi) {
#ifdef __CUDA_ARCH__
  asm("mad.lo.cc.u32   %8,  %17, %24, %8;\n\t"
      "madc.hi.cc.u32  %9,  %17, %24, %9;\n\t"
      "madc.lo.cc.u32  %10, %19, %24, %10;\n\t"
      "madc.hi.cc.u32  %11, %19, %24, %11;\n\t"
      "madc.lo.cc.u32  %12, %21, %24, %12;\n\t"
      "madc.hi.cc.u32  %13, %21, %24, %13;\n\t"
      "madc.lo.cc.u32  %14, %23, %24, %14;\n\t"
      "madc.hi.u32     %15, %23, %24, %15;\n\t"
      "mad.lo.cc.u32   %0,  %16, %24, %0;\n\t"
      "madc.hi.cc.u32  %1,  %16, %24, %1;\n\t"
      "madc.lo.cc.u32  %2,  %18, %24, %2;\n\t"
      "madc.hi.cc.u32  %3,  %18, %24, %3;\n\t"
      "madc.lo.cc.u32  %4,  %20, %24, %4;\n\t"
      "madc.hi.cc.u32  %5,  %20, %24, %5;\n\t"
      "madc.lo.cc.u32  %6,  %22, %24, %6;\n\t"
      "madc.hi.cc.u32  %7,  %22, %24, %7;\n\t"
      "addc.u32        %15, %15, 0;"
      : "+r"(e[0]), "+r"(e[1]), "+r"(e[2]), "+r"(e[3]), "+r"(e[4]), "+r"(e[5]), "+r"(e[6]), "+r"(e[7]),
        "+r"(o[0]), "+r"(o[1]), "+r"(o[2]), "+r"(o[3]), "+r"(o[4]), "+r"(o[5]), "+r"(o[6]), "+r"(o[7])
      : "r"(P::mod(0)), "r"(P::mod(1)), "r"(P::mod(2)), "r"(P::mod(3)), "r"(P::mod(4)), "r"(P::mod(5)),
        "r"(P::mod(6)), "r"(P::mod(7)), "r"(mi));
#else
  CC k;
  o[0] = k.mad_lo_cc(P::mod(1), mi, o[0]);
  o[1] = k.madc_hi_cc(P::mod(1), mi, o[1]);
  o[2] = k.madc_lo_cc(P::mod(3), mi, o[2]);
  o[3] = k.madc_hi_cc(P::mod(3), mi, o[3]);
  o[4] = k.madc_lo_cc(P::mod(5), mi, o[4]);
  o[5] = k.madc_hi_cc(P::mod(5), mi, o[5]);
  o[6] = k.madc_lo_cc(P::mod(7), mi, o[6]);
  o[7] = k.madc_hi(P::mod(7), mi, o[7]);
  e[0] = k.mad_lo_cc(P::mod(0), mi, e[0]);
  e[1] = k.madc_hi_cc(P::mod(0), mi, e[1]);
  e[2] = k.madc_lo_cc(P::mod(2), mi, e[2]);
  e[3] = k.madc_hi_cc(P::mod(2), mi, e[3]);
  e[4] = k.madc_lo_cc(P::mod(4), mi, e[4]);
  e[5] = k.madc_hi_cc(P::mod(4), mi, e[5]);
  e[6] = k.madc_lo_cc(P::mod(6), mi, e[6]);
  e[7] = k.madc_hi_cc(P::mod(6), mi, e[7]);
  o[7] = k.addc(o[7], 0);
#endif
}

// Block C: T += a * bi with no shift (a second product accumulated into the same running total: mont_mul2)
GPW_HD void mul_acc(uint32_t* e, uint32_t* o, const uint32_t* a, uint32_t bi) {
#ifdef __CUDA_ARCH__
  asm("mad.lo.cc.u32   %8,  %17, %24, %8;\n\t"
      "madc.hi.cc.u32  %9,  %17, %24, %9;\n\t"
      "madc.lo.cc.u32  %10, %19, %24, %10;\n\t"
      "madc.hi.cc.u32  %11, %19, %24, %11;\n\t"
      "madc.lo.cc.u32  %12, %21, %24, %12;\n\t"
      "madc.hi.cc.u32  %13, %21, %24, %13;\n\t"
      "madc.lo.cc.u32  %14, %23, %24, %14;\n\t"
      "madc.hi.u32     %15, %23, %24, %15;\n\t"
      "mad.lo.cc.u32   %0,  %16, %24, %0;\n\t"
      "madc.hi.cc.u32  %1,  %16, %24, %1;\n\t"
      "madc.lo.cc.u32  %2,  %18, %24, %2;\n\t"
      "madc.hi.cc.u32  %3,  %18, %24, %3;\n\t"
      "madc.lo.cc.u32  %4,  %20, %24, %4;\n\t"
      "madc.hi.cc.u32  %5,  %20, %24, %5;\n\t"
      "madc.lo.cc.u32  %6,  %22, %24, %6;\n\t"
      "madc.hi.cc.u32  %7,  %22, %24, %7;\n\t"
      "addc.u32        %15, %15, 0;"
      : "+r"(e[0]), "+r"(e[1]), "+r"(e[2]), "+r"(e[3]), "+r"(e[4]), "+r"(e[5]), "+r"(e[6]), "+r"(e[7]),
        "+r"(o[0]), "+r"(o[1]), "+r"(o[2]), "+r"(o[3]), "+r"(o[4]), "+r"(o[5]), "+r"(o[6]), "+r"(o[7])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]), "r"(bi));
#else
  CC k;
  o[0] = k.mad_lo_cc(a[1], bi, o[0]);
  o[1] = k.madc_hi_cc(a[1], bi, o[1]);
  o[2] = k.madc_lo_cc(a[3], bi, o[2]);
  o[3] = k.madc_hi_cc(a[3], bi, o[3]);
  o[4] = k.madc_lo_cc(a[5], bi, o[4]);
  o[5] = k.madc_hi_cc(a[5], bi, o[5]);
  o[6] = k.madc_lo_cc(a[7], bi, o[6]);
  o[7] = k.madc_hi(a[7], bi, o[7]);
  e[0] = k.mad_lo_cc(a[0], bi, e[0]);
  e[1] = k.madc_hi_cc(a[0], bi, e[1]);
  e[2] = k.madc_lo_cc(a[2], bi, e[2]);
  e[3] = k.madc_hi_cc(a[2], bi, e[3]);
  e[4] = k.madc_lo_cc(a[4], bi, e[4]);
  e[5] = k.madc_hi_cc(a[4], bi, e[5]);
  e[6] = k.madc_lo_cc(a[6], bi, e[6]);
  e[7] = k.madc_hi_cc(a[6], bi, e[7]);
  o[7] = k.addc(o[7], 0);
#endif
}

// lo, hi = a * b as ONE 32 x 32 -> 64 multiplication (mul.wide.u32 -> IMAD.WIDE.U32; a separate a * b and __umulhi(a, b)
// compile to an IMAD plus an IMAD.HI.U32 - two trips through the pipe that bounds every kernel of the path)
GPW_HD void mulwide32(uint32_t a, uint32_t b, uint32_t& lo, uint32_t& hi) {
#ifdef __CUDA_ARCH__
  asm("{\n\t.reg .u64 t;\n\tmul.wide.u32 t, %2, %3;\n\tmov.b64 {%0, %1}, t;\n\t}" : "=r"(lo), "=r"(hi) : "r"(a), "r"(b));
#else
  const uint64_t p = (uint64_t)a * b;
  lo = (uint32_t)p;
  hi = (uint32_t)(p >> 32);
#endif
}

GPW_HD uint32_t mulhi32(uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
  return __umulhi(a, b);
#else
  return (uint32_t)(((uint64_t)a * b) >> 32);
#endif
}

}  // namespace detail

template <class P>
GPW_HD Fe<P> mont_mul_wide(const Fe<P>& a, const Fe<P>& b) {
  uint32_t ev[8], od[8];
  // iteration 0: plain products, no accumulation
#pragma unroll
  for (int j = 0; j < 8; j += 2) {
    detail::mulwide32(a.l[j], b.l[0], ev[j], ev[j + 1]);
    detail::mulwide32(a.l[j + 1], b.l[0], od[j], od[j + 1]);
  }
  detail::redc_step<P>(ev, od, ev[0] * P::M0);
#pragma unroll
  for (int i = 1; i < 8; i += 2) {
    // roles swap after every one-limb shift
    detail::mul_acc_shift(od, ev, a.l, b.l[i]);
    detail::redc_step<P>(od, ev, od[0] * P::M0);
    if (i + 1 < 8) {
      detail::mul_acc_shift(ev, od, a.l, b.l[i + 1]);
      detail::redc_step<P>(ev, od, ev[0] * P::M0);
    }
  }
  // last call had (E, O) = (od, ev): result[k] = O[k] + E[k+1]
  Fe<P> r;
#ifdef __CUDA_ARCH__
  asm("add.cc.u32  %0, %8,  %16;\n\t"
      "addc.cc.u32 %1, %9,  %17;\n\t"
      "addc.cc.u32 %2, %10, %18;\n\t"
      "addc.cc.u32 %3, %11, %19;\n\t"
      "addc.cc.u32 %4, %12, %20;\n\t"
      "addc.cc.u32 %5, %13, %21;\n\t"
      "addc.cc.u32 %6, %14, %22;\n\t"
      "addc.u32    %7, %15, 0;"
      : "=r"(r.l[0]), "=r"(r.l[1]), "=r"(r.l[2]), "=r"(r.l[3]), "=r"(r.l[4]), "=r"(r.l[5]), "=r"(r.l[6]),
        "=r"(r.l[7])
      : "r"(ev[0]), "r"(ev[1]), "r"(ev[2]), "r"(ev[3]), "r"(ev[4]), "r"(ev[5]), "r"(ev[6]), "r"(ev[7]),
        "r"(od[1]), "r"(od[2]), "r"(od[3]), "r"(od[4]), "r"(od[5]), "r"(od[6]), "r"(od[7]));
#else
  detail::CC k;
  r.l[0] = k.add_cc(ev[0], od[1]);
  for (int i = 1; i < 7; i++) {
    uint64_t s = (uint64_t)ev[i] + od[i + 1] + k.c;
    r.l[i] = (uint32_t)s;
    k.c = (uint32_t)(s >> 32);
  }
  r.l[7] = ev[7] + k.c;
#endif
  return reduce_once(r);
}

// (a b + c d) R^-1 mod p with ONE Montgomery reduction: both products are accumulated into the same running total, so
// the pair costs 8 x (8 + 8 + 8) + 8 = 200 IMAD.WIDE instead of the 272 of two multiplications. Operands < p; the running
// total stays below 3p < 2^256 and the pre-shift sums below 2^288 (the capacity of the even / odd limb pairs); the result
// is < 1.4 p before the final conditional subtraction. Used for Y3 = R (Q - X3) + (p - Y1) PPP of the mixed addition.
template <class P>
GPW_HD Fe<P> mont_mul2(const Fe<P>& a, const Fe<P>& b, const Fe<P>& c, const Fe<P>& d) {
  uint32_t ev[8], od[8];
#pragma unroll
  for (int j = 0; j < 8; j += 2) {
    detail::mulwide32(a.l[j], b.l[0], ev[j], ev[j + 1]);
    detail::mulwide32(a.l[j + 1], b.l[0], od[j], od[j + 1]);
  }
  detail::mul_acc(ev, od, c.l, d.l[0]);
  detail::redc_step<P>(ev, od, ev[0] * P::M0);
#pragma unroll
  for (int i = 1; i < 8; i += 2) {
    detail::mul_acc_shift(od, ev, a.l, b.l[i]);
    detail::mul_acc(od, ev, c.l, d.l[i]);
    detail::redc_step<P>(od, ev, od[0] * P::M0);
    if (i + 1 < 8) {
      detail::mul_acc_shift(ev, od, a.l, b.l[i + 1]);
      detail::mul_acc(ev, od, c.l, d.l[i + 1]);
      detail::redc_step<P>(ev, od, ev[0] * P::M0);
    }
  }
  Fe<P> r;
#ifdef __CUDA_ARCH__
  asm("add.cc.u32  %0, %8,  %16;\n\t"
      "addc.cc.u32 %1, %9,  %17;\n\t"
      "addc.cc.u32 %2, %10, %18;\n\t"
      "addc.cc.u32 %3, %11, %19;\n\t"
      "addc.cc.u32 %4, %12, %20;\n\t"
      "addc.cc.u32 %5, %13, %21;\n\t"
      "addc.cc.u32 %6, %14, %22;\n\t"
      "addc.u32    %7, %15, 0;"
      : "=r"(r.l[0]), "=r"(r.l[1]), "=r"(r.l[2]), "=r"(r.l[3]), "=r"(r.l[4]), "=r"(r.l[5]), "=r"(r.l[6]),
        "=r"(r.l[7])
      : "r"(ev[0]), "r"(ev[1]), "r"(ev[2]), "r"(ev[3]), "r"(ev[4]), "r"(ev[5]), "r"(ev[6]), "r"(ev[7]),
        "r"(od[1]), "r"(od[2]), "r"(od[3]), "r"(od[4]), "r"(od[5]), "r"(od[6]), "r"(od[7]));
#else
  detail::CC k;
  r.l[0] = k.add_cc(ev[0], od[1]);
  for (int i = 1; i < 7; i++) {
    uint64_t s = (uint64_t)ev[i] + od[i + 1] + k.c;
    r.l[i] = (uint32_t)s;
    k.c = (uint32_t)(s >> 32);
  }
  r.l[7] = ev[7] + k.c;
#endif
  return reduce_once(r);
}


// ---- Montgomery squaring -------------------------------------------------------------------------------------------
// a^2 = D + 2 S with D = sum a_i^2 2^(64 i) and S = sum_{i<j} a_i a_j 2^(32 (i+j)): 28 + 8 = 36 IMAD.WIDE for the 512-bit
// square instead of 64, then a plain REDC of the 16 limbs with the same even/odd reduction rows as mont_mul_wide
// (8 x 8 = 64 IMAD.WIDE): 100 instead of 128 IMAD.WIDE per squaring. The extra work (merging the even / odd cross sums,
// doubling by funnel shifts, feeding the upper limbs into the reduction window) is ~110 carry-chain additions on the ALU
// pipe, which the IMAD-bound group-addition kernels have to spare.
// S is split like the products of mont_mul_wide: cross products a_i a_j with i + j odd (one even, one odd limb: the 4 x 4
// product of the even limbs by the odd limbs) live at odd limb offsets (array o16), those with i + j even at even offsets
// (array e16). Every row below is ordered so that its last pair is fresh (zero) or its carry lands on a fresh limb: no
// carry ever has to ripple further.
namespace detail {

// o-row: x[0..5] += (a0, a2, a4) * m (three pairs), x[6..7] = a6 * m + carry   (x = o16 + 2 v)
GPW_HD void sqr_row4(uint32_t* x, uint32_t a0, uint32_t a2, uint32_t a4, uint32_t a6, uint32_t m) {
#ifdef __CUDA_ARCH__
  asm("mad.lo.cc.u32   %0, %8,  %12, %0;\n\t"
      "madc.hi.cc.u32  %1, %8,  %12, %1;\n\t"
      "madc.lo.cc.u32  %2, %9,  %12, %2;\n\t"
      "madc.hi.cc.u32  %3, %9,  %12, %3;\n\t"
      "madc.lo.cc.u32  %4, %10, %12, %4;\n\t"
      "madc.hi.cc.u32  %5, %10, %12, %5;\n\t"
      "madc.lo.cc.u32  %6, %11, %12, 0;\n\t"
      "madc.hi.u32     %7, %11, %12, 0;"
      : "+r"(x[0]), "+r"(x[1]), "+r"(x[2]), "+r"(x[3]), "+r"(x[4]), "+r"(x[5]), "=r"(x[6]), "=r"(x[7])
      : "r"(a0), "r"(a2), "r"(a4), "r"(a6), "r"(m));
#else
  CC k;
  x[0] = k.mad_lo_cc(a0, m, x[0]);
  x[1] = k.madc_hi_cc(a0, m, x[1]);
  x[2] = k.madc_lo_cc(a2, m, x[2]);
  x[3] = k.madc_hi_cc(a2, m, x[3]);
  x[4] = k.madc_lo_cc(a4, m, x[4]);
  x[5] = k.madc_hi_cc(a4, m, x[5]);
  x[6] = k.madc_lo_cc(a6, m, 0);
  x[7] = k.madc_hi(a6, m, 0);
#endif
}

// x[0..3] += (b0, b1) * m, x[4..5] = b2 * m + carry
GPW_HD void sqr_row3(uint32_t* x, uint32_t b0, uint32_t b1, uint32_t b2, uint32_t m) {
#ifdef __CUDA_ARCH__
  asm("mad.lo.cc.u32   %0, %6, %9, %0;\n\t"
      "madc.hi.cc.u32  %1, %6, %9, %1;\n\t"
      "madc.lo.cc.u32  %2, %7, %9, %2;\n\t"
      "madc.hi.cc.u32  %3, %7, %9, %3;\n\t"
      "madc.lo.cc.u32  %4, %8, %9, 0;\n\t"
      "madc.hi.u32     %5, %8, %9, 0;"
      : "+r"(x[0]), "+r"(x[1]), "+r"(x[2]), "+r"(x[3]), "=r"(x[4]), "=r"(x[5])
      : "r"(b0), "r"(b1), "r"(b2), "r"(m));
#else
  CC k;
  x[0] = k.mad_lo_cc(b0, m, x[0]);
  x[1] = k.madc_hi_cc(b0, m, x[1]);
  x[2] = k.madc_lo_cc(b1, m, x[2]);
  x[3] = k.madc_hi_cc(b1, m, x[3]);
  x[4] = k.madc_lo_cc(b2, m, 0);
  x[5] = k.madc_hi(b2, m, 0);
#endif
}

// x[0..3] += (b0, b1) * m, x[4] = carry
GPW_HD void sqr_row2c(uint32_t* x, uint32_t b0, uint32_t b1, uint32_t m) {
#ifdef __CUDA_ARCH__
  asm("mad.lo.cc.u32   %0, %5, %7, %0;\n\t"
      "madc.hi.cc.u32  %1, %5, %7, %1;\n\t"
      "madc.lo.cc.u32  %2, %6, %7, %2;\n\t"
      "madc.hi.cc.u32  %3, %6, %7, %3;\n\t"
      "addc.u32        %4, 0, 0;"
      : "+r"(x[0]), "+r"(x[1]), "+r"(x[2]), "+r"(x[3]), "=r"(x[4])
      : "r"(b0), "r"(b1), "r"(m));
#else
  CC k;
  x[0] = k.mad_lo_cc(b0, m, x[0]);
  x[1] = k.madc_hi_cc(b0, m, x[1]);
  x[2] = k.madc_lo_cc(b1, m, x[2]);
  x[3] = k.madc_hi_cc(b1, m, x[3]);
  x[4] = k.addc(0, 0);
#endif
}

// x[0..2] += (b0, b1) * m with x[3] fresh: x[0..1] += b0 m, x[2] = lo(b1 m) + x[2] + carry, x[3] = hi(b1 m) + carry
GPW_HD void sqr_row2f(uint32_t* x, uint32_t b0, uint32_t b1, uint32_t m) {
#ifdef __CUDA_ARCH__
  asm("mad.lo.cc.u32   %0, %4, %6, %0;\n\t"
      "madc.hi.cc.u32  %1, %4, %6, %1;\n\t"
      "madc.lo.cc.u32  %2, %5, %6, %2;\n\t"
      "madc.hi.u32     %3, %5, %6, 0;"
      : "+r"(x[0]), "+r"(x[1]), "+r"(x[2]), "=r"(x[3])
      : "r"(b0), "r"(b1), "r"(m));
#else
  CC k;
  x[0] = k.mad_lo_cc(b0, m, x[0]);
  x[1] = k.madc_hi_cc(b0, m, x[1]);
  x[2] = k.madc_lo_cc(b1, m, x[2]);
  x[3] = k.madc_hi(b1, m, 0);
#endif
}

// x[0..1] += b0 * m, x[2] = carry
GPW_HD void sqr_row1c(uint32_t* x, uint32_t b0, uint32_t m) {
#ifdef __CUDA_ARCH__
  asm("mad.lo.cc.u32   %0, %3, %4, %0;\n\t"
      "madc.hi.cc.u32  %1, %3, %4, %1;\n\t"
      "addc.u32        %2, 0, 0;"
      : "+r"(x[0]), "+r"(x[1]), "=r"(x[2])
      : "r"(b0), "r"(m));
#else
  CC k;
  x[0] = k.mad_lo_cc(b0, m, x[0]);
  x[1] = k.madc_hi_cc(b0, m, x[1]);
  x[2] = k.addc(0, 0);
#endif
}

// x[0] += lo(b0 m), x[1] = hi(b0 m) + carry   (x[1] fresh)
GPW_HD void sqr_row1f(uint32_t* x, uint32_t b0, uint32_t m) {
#ifdef __CUDA_ARCH__
  asm("mad.lo.cc.u32   %0, %2, %3, %0;\n\t"
      "madc.hi.u32     %1, %2, %3, 0;"
      : "+r"(x[0]), "=r"(x[1])
      : "r"(b0), "r"(m));
#else
  CC k;
  x[0] = k.mad_lo_cc(b0, m, x[0]);
  x[1] = k.madc_hi(b0, m, 0);
#endif
}

// s[1..15] = e[1..15] + o[0..14] (the odd-offset sum shifted up by one limb); e[0] = e[1] = 0, o[14] = 0 on entry.
// Two asm statements of 7 limbs (operand-count limit); the carry between them is re-created by c + 0xffffffff.
GPW_HD void sqr_merge(uint32_t* s, const uint32_t* e, const uint32_t* o) {
  s[0] = 0;
  s[1] = o[0];
#ifdef __CUDA_ARCH__
  uint32_t c;
  asm("add.cc.u32   %0, %8,  %15;\n\t"
      "addc.cc.u32  %1, %9,  %16;\n\t"
      "addc.cc.u32  %2, %10, %17;\n\t"
      "addc.cc.u32  %3, %11, %18;\n\t"
      "addc.cc.u32  %4, %12, %19;\n\t"
      "addc.cc.u32  %5, %13, %20;\n\t"
      "addc.cc.u32  %6, %14, %21;\n\t"
      "addc.u32     %7, 0, 0;"
      : "=r"(s[2]), "=r"(s[3]), "=r"(s[4]), "=r"(s[5]), "=r"(s[6]), "=r"(s[7]), "=r"(s[8]), "=r"(c)
      : "r"(e[2]), "r"(e[3]), "r"(e[4]), "r"(e[5]), "r"(e[6]), "r"(e[7]), "r"(e[8]),
        "r"(o[1]), "r"(o[2]), "r"(o[3]), "r"(o[4]), "r"(o[5]), "r"(o[6]), "r"(o[7]));
  asm("add.cc.u32   %0, %7,  0xffffffff;\n\t"
      "addc.cc.u32  %0, %8,  %14;\n\t"
      "addc.cc.u32  %1, %9,  %15;\n\t"
      "addc.cc.u32  %2, %10, %16;\n\t"
      "addc.cc.u32  %3, %11, %17;\n\t"
      "addc.cc.u32  %4, %12, %18;\n\t"
      "addc.cc.u32  %5, %13, %19;\n\t"
      "addc.u32     %6, 0, 0;"
      : "=&r"(s[9]), "=&r"(s[10]), "=&r"(s[11]), "=&r"(s[12]), "=&r"(s[13]), "=&r"(s[14]), "=&r"(s[15])
      : "r"(c), "r"(e[9]), "r"(e[10]), "r"(e[11]), "r"(e[12]), "r"(e[13]), "r"(e[14]),
        "r"(o[8]), "r"(o[9]), "r"(o[10]), "r"(o[11]), "r"(o[12]), "r"(o[13]));
#else
  CC k;
  s[2] = k.add_cc(e[2], o[1]);
  for (int i = 3; i <= 14; i++) {
    uint64_t v = (uint64_t)e[i] + o[i - 1] + k.c;
    s[i] = (uint32_t)v;
    k.c = (uint32_t)(v >> 32);
  }
  s[15] = k.c;
#endif
}

// t[0..15] += sum a_i^2 2^(64 i)   (t = 2 S < 2^512 - a^2: the top carry is zero)
GPW_HD void sqr_diag(uint32_t* t, const uint32_t* a) {
#ifdef __CUDA_ARCH__
  asm("mad.lo.cc.u32   %0,  %16, %16, %0;\n\t"
      "madc.hi.cc.u32  %1,  %16, %16, %1;\n\t"
      "madc.lo.cc.u32  %2,  %17, %17, %2;\n\t"
      "madc.hi.cc.u32  %3,  %17, %17, %3;\n\t"
      "madc.lo.cc.u32  %4,  %18, %18, %4;\n\t"
      "madc.hi.cc.u32  %5,  %18, %18, %5;\n\t"
      "madc.lo.cc.u32  %6,  %19, %19, %6;\n\t"
      "madc.hi.cc.u32  %7,  %19, %19, %7;\n\t"
      "madc.lo.cc.u32  %8,  %20, %20, %8;\n\t"
      "madc.hi.cc.u32  %9,  %20, %20, %9;\n\t"
      "madc.lo.cc.u32  %10, %21, %21, %10;\n\t"
      "madc.hi.cc.u32  %11, %21, %21, %11;\n\t"
      "madc.lo.cc.u32  %12, %22, %22, %12;\n\t"
      "madc.hi.cc.u32  %13, %22, %22, %13;\n\t"
      "madc.lo.cc.u32  %14, %23, %23, %14;\n\t"
      "madc.hi.u32     %15, %23, %23, %15;"
      : "+r"(t[0]), "+r"(t[1]), "+r"(t[2]), "+r"(t[3]), "+r"(t[4]), "+r"(t[5]), "+r"(t[6]), "+r"(t[7]), "+r"(t[8]), "+r"(t[9]),
        "+r"(t[10]), "+r"(t[11]), "+r"(t[12]), "+r"(t[13]), "+r"(t[14]), "+r"(t[15])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]));
#else
  CC k;
  for (int i = 0; i < 8; i++) {
    t[2 * i] = i == 0 ? k.mad_lo_cc(a[i], a[i], t[0]) : k.madc_lo_cc(a[i], a[i], t[2 * i]);
    t[2 * i + 1] = i == 7 ? k.madc_hi(a[i], a[i], t[15]) : k.madc_hi_cc(a[i], a[i], t[2 * i + 1]);
  }
#endif
}

// Reduction window moves up by one limb: e[0] += o[1]; o = (o >> 64) with the next limb `tn` of the square entering at
// window position 7 (o[6]) and the last carry at position 8 (o[7]).  (e, o) = (previous odd, previous even) accumulator.
GPW_HD void redc_shift_in(uint32_t* e, uint32_t* o, uint32_t tn) {
#ifdef __CUDA_ARCH__
  asm("add.cc.u32   %0, %0, %2;\n\t"
      "addc.cc.u32  %1, %3, 0;\n\t"
      "addc.cc.u32  %2, %4, 0;\n\t"
      "addc.cc.u32  %3, %5, 0;\n\t"
      "addc.cc.u32  %4, %6, 0;\n\t"
      "addc.cc.u32  %5, %7, 0;\n\t"
      "addc.cc.u32  %6, %8, 0;\n\t"
      "addc.cc.u32  %7, %9, 0;\n\t"
      "addc.u32     %8, 0, 0;"
      : "+r"(e[0]), "+r"(o[0]), "+r"(o[1]), "+r"(o[2]), "+r"(o[3]), "+r"(o[4]), "+r"(o[5]), "+r"(o[6]), "+r"(o[7])
      : "r"(tn));
#else
  CC k;
  e[0] = k.add_cc(e[0], o[1]);
  for (int i = 0; i < 6; i++) {
    uint64_t v = (uint64_t)o[i + 2] + k.c;
    o[i] = (uint32_t)v;
    k.c = (uint32_t)(v >> 32);
  }
  uint64_t v = (uint64_t)tn + k.c;
  o[6] = (uint32_t)v;
  o[7] = (uint32_t)(v >> 32);
#endif
}

}  // namespace detail

template <class P>
GPW_HD Fe<P> mont_sqr_wide(const Fe<P>& a) {
  const uint32_t* x = a.l;
  // odd-offset cross sum: (a0, a2, a4, a6) x (a1, a3, a5, a7); o16[k] has weight 2^(32 (k + 1))
  uint32_t o16[14];
#pragma unroll
  for (int j = 0; j < 4; j++) {
    detail::mulwide32(x[2 * j], x[1], o16[2 * j], o16[2 * j + 1]);
  }
  detail::sqr_row4(o16 + 2, x[0], x[2], x[4], x[6], x[3]);
  detail::sqr_row4(o16 + 4, x[0], x[2], x[4], x[6], x[5]);
  detail::sqr_row4(o16 + 6, x[0], x[2], x[4], x[6], x[7]);
  // even-offset cross sum: even limbs among themselves, odd limbs among themselves; e16[k] has weight 2^(32 k)
  uint32_t e16[15];
  e16[0] = e16[1] = 0;
  e16[14] = 0;
#pragma unroll
  for (int j = 1; j < 4; j++) {  // a0 x (a2, a4, a6) at pairs 1..3
    detail::mulwide32(x[0], x[2 * j], e16[2 * j], e16[2 * j + 1]);
  }
  detail::sqr_row3(e16 + 4, x[3], x[5], x[7], x[1]);  // a1 x (a3, a5, a7) at pairs 2..4   (8, 9 fresh)
  detail::sqr_row2c(e16 + 6, x[4], x[6], x[2]);       // a2 x (a4, a6) at pairs 3, 4       (carry -> 10)
  detail::sqr_row2f(e16 + 8, x[5], x[7], x[3]);       // a3 x (a5, a7) at pairs 4, 5       (11 fresh)
  detail::sqr_row1c(e16 + 10, x[6], x[4]);            // a4 x a6 at pair 5                 (carry -> 12)
  detail::sqr_row1f(e16 + 12, x[7], x[5]);            // a5 x a7 at pair 6                 (13 fresh)
  // S = e16 + (o16 << 32), T = 2 S + D
  uint32_t s[16], t[16];
  {
    uint32_t o15[15];
#pragma unroll
    for (int i = 0; i < 14; i++) o15[i] = o16[i];
    o15[14] = 0;
    detail::sqr_merge(s, e16, o15);
  }
  t[0] = 0;
#pragma unroll
  for (int i = 1; i < 16; i++) {
#ifdef __CUDA_ARCH__
    t[i] = __funnelshift_l(s[i - 1], s[i], 1);
#else
    t[i] = (s[i] << 1) | (s[i - 1] >> 31);
#endif
  }
  detail::sqr_diag(t, x);
  // REDC: window (ev, od) as in mont_mul_wide; t[8] starts at window position 8 (od[7])... see redc_shift_in
  uint32_t ev[8], od[8];
#pragma unroll
  for (int i = 0; i < 8; i++) {
    ev[i] = t[i];
    od[i] = 0;
  }
  detail::redc_step<P>(ev, od, ev[0] * P::M0);
#pragma unroll
  for (int i = 1; i < 8; i += 2) {
    detail::redc_shift_in(od, ev, t[7 + i]);
    detail::redc_step<P>(od, ev, od[0] * P::M0);
    if (i + 1 < 8) {
      detail::redc_shift_in(ev, od, t[8 + i]);
      detail::redc_step<P>(ev, od, ev[0] * P::M0);
    }
  }
  // (E, O) = (od, ev): result[k] = O[k] + E[k + 1], plus the square's top limb t[15] at position 7
  Fe<P> r;
#ifdef __CUDA_ARCH__
  asm("add.cc.u32  %0, %8,  %16;\n\t"
      "addc.cc.u32 %1, %9,  %17;\n\t"
      "addc.cc.u32 %2, %10, %18;\n\t"
      "addc.cc.u32 %3, %11, %19;\n\t"
      "addc.cc.u32 %4, %12, %20;\n\t"
      "addc.cc.u32 %5, %13, %21;\n\t"
      "addc.cc.u32 %6, %14, %22;\n\t"
      "addc.u32    %7, %15, %23;"
      : "=r"(r.l[0]), "=r"(r.l[1]), "=r"(r.l[2]), "=r"(r.l[3]), "=r"(r.l[4]), "=r"(r.l[5]), "=r"(r.l[6]),
        "=r"(r.l[7])
      : "r"(ev[0]), "r"(ev[1]), "r"(ev[2]), "r"(ev[3]), "r"(ev[4]), "r"(ev[5]), "r"(ev[6]), "r"(ev[7]),
        "r"(od[1]), "r"(od[2]), "r"(od[3]), "r"(od[4]), "r"(od[5]), "r"(od[6]), "r"(od[7]), "r"(t[15]));
#else
  detail::CC k;
  r.l[0] = k.add_cc(ev[0], od[1]);
  for (int i = 1; i < 7; i++) {
    uint64_t v = (uint64_t)ev[i] + od[i + 1] + k.c;
    r.l[i] = (uint32_t)v;
    k.c = (uint32_t)(v >> 32);
  }
  r.l[7] = ev[7] + t[15] + k.c;
#endif
  return reduce_once(r);
}

#ifndef GPW_FF_PORTABLE_MUL
template <class P>
GPW_HD Fe<P> mul(const Fe<P>& a, const Fe<P>& b) {
  return mont_mul_wide(a, b);
}
#else
template <class P>
GPW_HD Fe<P> mul(const Fe<P>& a, const Fe<P>& b) {
  return mont_mul_portable(a, b);
}
#endif

template <class P>
GPW_HD Fe<P> sqr(const Fe<P>& a) {
#ifndef GPW_FF_PORTABLE_MUL
  return mont_sqr_wide(a);
#else
  return mul(a, a);
#endif
}

// Squaring for LATENCY-bound code (a lone warp walking a dependency chain: the solve spine, k_tape_poseidon4): the dedicated
// squaring saves 28 IMAD.WIDE but its critical path is longer than the multiplication's (cross sums -> merge -> doubling ->
// diagonal chain -> eight shift-in + reduction rows, ~215 dependent instructions against ~130), and such code waits for
// exactly that path. Measured on the Merkle levels: see DESIGN.md.
template <class P>
GPW_HD Fe<P> sqr_chain(const Fe<P>& a) {
  return mul(a, a);
}

// a b - c d: one Montgomery reduction for both products (mont_mul2 with the second product negated through c)
template <class P>
GPW_HD Fe<P> mul_sub2(const Fe<P>& a, const Fe<P>& b, const Fe<P>& c, const Fe<P>& d) {
#ifndef GPW_FF_PORTABLE_MUL
  return mont_mul2(a, b, neg(c), d);
#else
  return sub(mul(a, b), mul(c, d));
#endif
}

template <class P>
GPW_HD Fe<P> to_mont(const Fe<P>& a) {
  return mul(a, Fe<P>::r2());
}

template <class P>
GPW_HD Fe<P> from_mont(const Fe<P>& a) {
  Fe<P> o = Fe<P>::zero();
  o.l[0] = 1;
  return mul(a, o);
}

// a^e for a 256-bit exponent given as 8 LE u32 words (square-and-multiply, MSB first)
template <class P>
GPW_HD Fe<P> pow_words(const Fe<P>& a, const uint32_t* e, int nwords) {
  Fe<P> r = Fe<P>::one();
  bool started = false;
  for (int w = nwords - 1; w >= 0; w--) {
    for (int b = 31; b >= 0; b--) {
      if (started) r = sqr(r);
      if ((e[w] >> b) & 1u) {
        r = started ? mul(r, a) : a;
        started = true;
      }
    }
  }
  return r;
}

// Fermat inverse (0 -> 0). Used off the hot path and in batched-inverse tails.
template <class P>
GPW_HD Fe<P> inv(const Fe<P>& a) {
  uint32_t e[8];
#pragma unroll
  for (int i = 0; i < 8; i++) e[i] = P::mod(i);
  e[0] -= 2;  // mod(0) >= 2 for both fields
  if (a.is_zero()) return a;
  return pow_words(a, e, 8);
}

// Inverse by the binary extended Euclidean algorithm (0 -> 0): shifts, additions and subtractions only. On the GPU it runs on
// the ALU pipe and leaves the IMAD pipe - the bottleneck of every kernel here - to the other warps; ~770 loop steps of ~30
// simple instructions instead of the ~380 Montgomery multiplications (52 k IMAD.WIDE) of the Fermat inverse. Used for the one
// inversion per CTA batch of the batch-affine bucket accumulation (msm_affine.cuh). Montgomery in, Montgomery out.
template <class P>
GPW_HD Fe<P> inv_euclid(const Fe<P>& a_mont) {
  if (a_mont.is_zero()) return a_mont;
  const Fe<P> p = modulus<P>();
  Fe<P> u = a_mont, v = p, x1 = Fe<P>::zero(), x2 = Fe<P>::zero();
  x1.l[0] = 1;
  // invariants: x1 a = u, x2 a = v (mod p); u, v odd after their even parts are stripped; gcd(a, p) = 1
  auto halve = [&](Fe<P>& x) {  // x <- x / 2 mod p for x in [0, p)
    uint32_t carry = 0;
    if (x.l[0] & 1u) carry = add_raw(x, x, p);  // p < 2^254: the sum fits 255 bits, carry is always 0 (kept for clarity)
#pragma unroll
    for (int i = 0; i < 7; i++) x.l[i] = (x.l[i] >> 1) | (x.l[i + 1] << 31);
    x.l[7] = (x.l[7] >> 1) | (carry << 31);
  };
  auto shr1 = [](Fe<P>& x) {
#pragma unroll
    for (int i = 0; i < 7; i++) x.l[i] = (x.l[i] >> 1) | (x.l[i + 1] << 31);
    x.l[7] >>= 1;
  };
  auto is_one = [](const Fe<P>& x) {
    uint32_t o = x.l[0] ^ 1u;
#pragma unroll
    for (int i = 1; i < 8; i++) o |= x.l[i];
    return o == 0;
  };
  while (!is_one(u) && !is_one(v)) {
    while (!(u.l[0] & 1u)) {
      shr1(u);
      halve(x1);
    }
    while (!(v.l[0] & 1u)) {
      shr1(v);
      halve(x2);
    }
    Fe<P> t;
    if (sub_raw(t, u, v) == 0) {  // u >= v
      u = t;
      x1 = sub(x1, x2);
    } else {
      sub_raw(v, v, u);
      x2 = sub(x2, x1);
    }
  }
  // (a R)^-1 -> a^-1 R: two multiplications by R^2
  const Fe<P> r = is_one(u) ? x1 : x2;
  return mul(mul(r, Fe<P>::r2()), Fe<P>::r2());
}

using Fr = Fe<FrParams>;
using Fp = Fe<FpParams>;

// ---------------------------------------------------------------------------------------------
// Fp2 = Fp[u]/(u^2+1)   - coordinates of G2
// ---------------------------------------------------------------------------------------------
struct alignas(16) Fp2 {
  Fp c0, c1;
  static GPW_HD Fp2 zero() { return {Fp::zero(), Fp::zero()}; }
  static GPW_HD Fp2 one() { return {Fp::one(), Fp::zero()}; }
  GPW_HD bool is_zero() const { return c0.is_zero() && c1.is_zero(); }
  GPW_HD bool operator==(const Fp2& b) const { return c0 == b.c0 && c1 == b.c1; }
  GPW_HD bool operator!=(const Fp2& b) const { return !(*this == b); }
};

GPW_HD Fp2 add(const Fp2& a, const Fp2& b) { return {add(a.c0, b.c0), add(a.c1, b.c1)}; }
GPW_HD Fp2 sub(const Fp2& a, const Fp2& b) { return {sub(a.c0, b.c0), sub(a.c1, b.c1)}; }
GPW_HD Fp2 neg(const Fp2& a) { return {neg(a.c0), neg(a.c1)}; }
GPW_HD Fp2 dbl(const Fp2& a) { return {dbl(a.c0), dbl(a.c1)}; }
// (a0 b0 - a1 b1) + (a0 b1 + a1 b0) u: two dual-product multiplications (one Montgomery reduction each, 2 x 200
// IMAD.WIDE) instead of Karatsuba's three multiplications (3 x 136) plus five 256-bit additions / subtractions and their
// temporaries - the same multiplier work with less register pressure, which is what limits the G2 bucket accumulation
GPW_HD Fp2 mul(const Fp2& a, const Fp2& b) {
#ifndef GPW_FF_PORTABLE_MUL
  return {mul_sub2(a.c0, b.c0, a.c1, b.c1), mont_mul2(a.c0, b.c1, a.c1, b.c0)};
#else
  Fp t0 = mul(a.c0, b.c0);
  Fp t1 = mul(a.c1, b.c1);
  Fp s = mul(add(a.c0, a.c1), add(b.c0, b.c1));
  return {sub(t0, t1), sub(sub(s, t0), t1)};
#endif
}
// (a0+a1 u)^2 = (a0+a1)(a0-a1) + 2 a0 a1 u : 2 Fp muls
GPW_HD Fp2 sqr(const Fp2& a) {
  Fp t = mul(a.c0, a.c1);
  Fp r0 = mul(add(a.c0, a.c1), sub(a.c0, a.c1));
  return {r0, dbl(t)};
}
GPW_HD Fp2 mul_sub2(const Fp2& a, const Fp2& b, const Fp2& c, const Fp2& d) { return sub(mul(a, b), mul(c, d)); }
GPW_HD Fp2 inv(const Fp2& a) {
  Fp n = add(sqr(a.c0), sqr(a.c1));
  Fp ni = inv(n);
  return {mul(a.c0, ni), neg(mul(a.c1, ni))};
}

}  // namespace gpw
