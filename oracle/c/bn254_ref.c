/* ORACLE (test infrastructure + bench.py's cpu_baseline / --impl reference leg ONLY; never linked into
 * libgpw.so and never called by the product path).
 *
 * Plain-C CPU restatement of the BN254 prover primitives that the reference reaches through its
 * un-vendored dependency gnark-crypto v0.12.2-0.20231013160410-1f65e75b6dfb (go.mod:7): the call sites
 * are groth16.Prove / plonk.Prove at benchmark.go:249,162. gnark-crypto's sources
 * are not in /root/reference and there is no Go toolchain here, so this follows its PUBLISHED
 * algorithms: G1Jac/G2Jac.MultiExp = Pippenger bucket method with signed c-bit digits, extended-
 * Jacobian (XYZZ) bucket accumulators, windows processed in parallel; fr/fft = in-place radix-2
 * DIF + bit-reversal with coset generator 5. "parity unpinned" by reference goldens (none exist for
 * MSM/FFT); pinned by mathematics in tests/test_oracle_bn254.py (unique group element /
 * DFT definition). Written independently of csrc/ (4 x 64-bit limbs, unsigned __int128).
 *
 * Build: make -C oracle/c   ->  oracle/c/libbn254_ref.so  (gcc -O3 -march=native -fopenmp)
 */
#include <omp.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef unsigned __int128 u128;
typedef struct { uint64_t l[4]; } fe;
typedef struct { uint64_t p[4]; uint64_t inv; fe one; fe r2; } field_t;

static const field_t FP = {
    {0x3c208c16d87cfd47ull, 0x97816a916871ca8dull, 0xb85045b68181585dull, 0x30644e72e131a029ull},
    0x87d20782e4866389ull,
    {{0xd35d438dc58f0d9dull, 0x0a78eb28f5c70b3dull, 0x666ea36f7879462cull, 0x0e0a77c19a07df2full}},
    {{0xf32cfc5b538afa89ull, 0xb5e71911d44501fbull, 0x47ab1eff0a417ff6ull, 0x06d89f71cab8351full}}};
static const field_t FR = {
    {0x43e1f593f0000001ull, 0x2833e84879b97091ull, 0xb85045b68181585dull, 0x30644e72e131a029ull},
    0xc2e1f593efffffffull,
    {{0xac96341c4ffffffbull, 0x36fc76959f60cd29ull, 0x666ea36f7879462eull, 0x0e0a77c19a07df2full}},
    {{0x1bb8e645ae216da7ull, 0x53fe3ab1e35c59e3ull, 0x8c49833d53bb8085ull, 0x0216d0b17f4e44a5ull}}};

static inline int fe_is_zero(const fe* a) { return (a->l[0] | a->l[1] | a->l[2] | a->l[3]) == 0; }
static inline int fe_eq(const fe* a, const fe* b) { return memcmp(a, b, 32) == 0; }
static inline int fe_geq(const fe* a, const uint64_t* p) {
  for (int i = 3; i >= 0; i--) {
    if (a->l[i] > p[i]) return 1;
    if (a->l[i] < p[i]) return 0;
  }
  return 1;
}
static inline void fe_sub_p(fe* a, const uint64_t* p) {
  u128 b = 0;
  for (int i = 0; i < 4; i++) {
    u128 d = (u128)a->l[i] - p[i] - b;
    a->l[i] = (uint64_t)d;
    b = (d >> 64) & 1;
  }
}
static inline void fe_add(fe* r, const fe* a, const fe* b, const field_t* F) {
  u128 c = 0;
  for (int i = 0; i < 4; i++) {
    c += (u128)a->l[i] + b->l[i];
    r->l[i] = (uint64_t)c;
    c >>= 64;
  }
  if (fe_geq(r, F->p)) fe_sub_p(r, F->p);
}
static inline void fe_sub(fe* r, const fe* a, const fe* b, const field_t* F) {
  u128 brw = 0;
  fe t;
  for (int i = 0; i < 4; i++) {
    u128 d = (u128)a->l[i] - b->l[i] - brw;
    t.l[i] = (uint64_t)d;
    brw = (d >> 64) & 1;
  }
  if (brw) {
    u128 c = 0;
    for (int i = 0; i < 4; i++) {
      c += (u128)t.l[i] + F->p[i];
      t.l[i] = (uint64_t)c;
      c >>= 64;
    }
  }
  *r = t;
}
static inline void fe_neg(fe* r, const fe* a, const field_t* F) {
  fe z = {{0, 0, 0, 0}};
  fe_sub(r, &z, a, F);
}
/* CIOS Montgomery multiplication, 4 x 64-bit limbs */
static inline void fe_mul(fe* r, const fe* a, const fe* b, const field_t* F) {
  uint64_t t[6] = {0, 0, 0, 0, 0, 0};
  for (int i = 0; i < 4; i++) {
    u128 c = 0;
    for (int j = 0; j < 4; j++) {
      c += (u128)a->l[j] * b->l[i] + t[j];
      t[j] = (uint64_t)c;
      c >>= 64;
    }
    c += t[4];
    t[4] = (uint64_t)c;
    t[5] = (uint64_t)(c >> 64);
    uint64_t m = t[0] * F->inv;
    c = ((u128)m * F->p[0] + t[0]) >> 64;
    for (int j = 1; j < 4; j++) {
      c += (u128)m * F->p[j] + t[j];
      t[j - 1] = (uint64_t)c;
      c >>= 64;
    }
    c += t[4];
    t[3] = (uint64_t)c;
    t[4] = t[5] + (uint64_t)(c >> 64);
  }
  fe o = {{t[0], t[1], t[2], t[3]}};
  if (t[4] || fe_geq(&o, F->p)) fe_sub_p(&o, F->p);
  *r = o;
}
static void fe_pow(fe* r, const fe* a, const uint64_t e[4], const field_t* F) {
  fe acc = F->one, base = *a;
  for (int w = 0; w < 4; w++)
    for (int b = 0; b < 64; b++) {
      if ((e[w] >> b) & 1) fe_mul(&acc, &acc, &base, F);
      fe_mul(&base, &base, &base, F);
    }
  *r = acc;
}
static void fe_inv(fe* r, const fe* a, const field_t* F) {
  uint64_t e[4] = {F->p[0] - 2, F->p[1], F->p[2], F->p[3]};
  fe_pow(r, a, e, F);
}
static void fe_from_mont(fe* r, const fe* a, const field_t* F) {
  fe one = {{1, 0, 0, 0}};
  fe_mul(r, a, &one, F);
}
static void fe_from_u64(fe* r, uint64_t v, const field_t* F) {
  fe t = {{v, 0, 0, 0}};
  fe_mul(r, &t, &F->r2, F);
}

/* ---- Fp2 = Fp[u]/(u^2+1) ---- */
typedef struct { fe a, b; } fe2;
static inline int fe2_is_zero(const fe2* x) { return fe_is_zero(&x->a) && fe_is_zero(&x->b); }
static inline void fe2_add(fe2* r, const fe2* x, const fe2* y) { fe_add(&r->a, &x->a, &y->a, &FP); fe_add(&r->b, &x->b, &y->b, &FP); }
static inline void fe2_sub(fe2* r, const fe2* x, const fe2* y) { fe_sub(&r->a, &x->a, &y->a, &FP); fe_sub(&r->b, &x->b, &y->b, &FP); }
static inline void fe2_neg(fe2* r, const fe2* x) { fe_neg(&r->a, &x->a, &FP); fe_neg(&r->b, &x->b, &FP); }
static inline void fe2_mul(fe2* r, const fe2* x, const fe2* y) {
  fe t0, t1, t2, t3;
  fe_mul(&t0, &x->a, &y->a, &FP);
  fe_mul(&t1, &x->b, &y->b, &FP);
  fe_mul(&t2, &x->a, &y->b, &FP);
  fe_mul(&t3, &x->b, &y->a, &FP);
  fe_sub(&r->a, &t0, &t1, &FP);
  fe_add(&r->b, &t2, &t3, &FP);
}
static void fe2_inv(fe2* r, const fe2* x) {
  fe n, t, ni;
  fe_mul(&n, &x->a, &x->a, &FP);
  fe_mul(&t, &x->b, &x->b, &FP);
  fe_add(&n, &n, &t, &FP);
  fe_inv(&ni, &n, &FP);
  fe_mul(&r->a, &x->a, &ni, &FP);
  fe_mul(&t, &x->b, &ni, &FP);
  fe_neg(&r->b, &t, &FP);
}

/* ---- group law, instantiated twice (poor man's template) ---- */
#define T fe
#define NAME(x) g1_##x
#define F_ADD(r, a, b) fe_add(r, a, b, &FP)
#define F_SUB(r, a, b) fe_sub(r, a, b, &FP)
#define F_MUL(r, a, b) fe_mul(r, a, b, &FP)
#define F_NEG(r, a) fe_neg(r, a, &FP)
#define F_ISZERO(a) fe_is_zero(a)
#define F_INV(r, a) fe_inv(r, a, &FP)
#define F_SETONE(r) (*(r) = FP.one)
#include "ec_tmpl.h"
#undef T
#undef NAME
#undef F_ADD
#undef F_SUB
#undef F_MUL
#undef F_NEG
#undef F_ISZERO
#undef F_INV
#undef F_SETONE

#define T fe2
#define NAME(x) g2_##x
#define F_ADD(r, a, b) fe2_add(r, a, b)
#define F_SUB(r, a, b) fe2_sub(r, a, b)
#define F_MUL(r, a, b) fe2_mul(r, a, b)
#define F_NEG(r, a) fe2_neg(r, a)
#define F_ISZERO(a) fe2_is_zero(a)
#define F_INV(r, a) fe2_inv(r, a)
#define F_SETONE(r) do { (r)->a = FP.one; memset(&(r)->b, 0, 32); } while (0)
#include "ec_tmpl.h"

/* scalars: n x 4 u64 (Montgomery Fr if mont, else canonical). points: affine, Montgomery Fp. */
int ref_msm_g1(const uint64_t* scalars, const uint64_t* points, size_t n, int mont, int c, int nthreads, uint64_t* out) {
  return g1_msm((const fe*)scalars, (const g1_aff*)points, n, mont, c, nthreads, (g1_aff*)out);
}
int ref_msm_g2(const uint64_t* scalars, const uint64_t* points, size_t n, int mont, int c, int nthreads, uint64_t* out) {
  return g2_msm((const fe*)scalars, (const g2_aff*)points, n, mont, c, nthreads, (g2_aff*)out);
}

/* ---- radix-2 FFT over Fr: natural in -> natural out (DIF + bit reversal), gnark-crypto conventions ---- */
static uint32_t bitrev32(uint32_t x, int bits) {
  uint32_t r = 0;
  for (int i = 0; i < bits; i++) r |= ((x >> i) & 1u) << (bits - 1 - i);
  return r;
}

int ref_ntt_fr(uint64_t* data, int logn, int inverse, int coset, int nthreads) {
  if (logn < 0 || logn > 28) return -1;
  size_t n = (size_t)1 << logn;
  fe* a = (fe*)data;
  if (nthreads > 0) omp_set_num_threads(nthreads);
  /* w = 5^((r-1)/2^logn) */
  fe g, w;
  fe_from_u64(&g, 5, &FR);
  uint64_t e[4] = {FR.p[0] - 1, FR.p[1], FR.p[2], FR.p[3]};
  for (int s = 0; s < logn; s++) { /* e >>= 1 */
    for (int i = 0; i < 4; i++) e[i] = (e[i] >> 1) | (i < 3 ? e[i + 1] << 63 : 0);
  }
  fe_pow(&w, &g, e, &FR);
  if (inverse) fe_inv(&w, &w, &FR);
  /* twiddles, compacted per stage so that every stage reads them with stride 1: stage s uses w^(j << s), j < n >> (s+1),
   * stored at ctw + (n - (n >> s)). Cached per (logn, direction): gnark-crypto's fft.Domain precomputes them once too. */
  static fe* tw_cache[29][2];
  fe* ctw;
#pragma omp critical(ref_ntt_tw)
  {
    ctw = tw_cache[logn][inverse ? 1 : 0];
    if (!ctw && n > 1) {
      ctw = (fe*)malloc(n * sizeof(fe));
      if (ctw) {
        fe* tw = ctw; /* stage 0 = the plain table w^j, j < n/2 */
        int nt = omp_get_max_threads();
#pragma omp parallel for schedule(static) num_threads(nt)
        for (int t = 0; t < nt; t++) {
          size_t lo = (n / 2) * t / nt, hi = (n / 2) * (t + 1) / nt;
          uint64_t ee[4] = {lo, 0, 0, 0};
          fe cur;
          fe_pow(&cur, &w, ee, &FR);
          for (size_t j = lo; j < hi; j++) {
            tw[j] = cur;
            fe_mul(&cur, &cur, &w, &FR);
          }
        }
        for (int st = 1; st < logn; st++) {
          fe* dst = ctw + (n - (n >> st));
          size_t half = n >> (st + 1);
#pragma omp parallel for schedule(static)
          for (size_t j = 0; j < half; j++) dst[j] = tw[j << st];
        }
        tw_cache[logn][inverse ? 1 : 0] = ctw;
      }
    }
  }
  if (!ctw && n > 1) return -2;
  if (!inverse && coset) {
    /* a_j *= g^j : per-thread chunks seeded by one exponentiation each */
#pragma omp parallel
    {
      int nt = omp_get_num_threads(), t = omp_get_thread_num();
      size_t lo = n * t / nt, hi = n * (t + 1) / nt;
      uint64_t ee[4] = {lo, 0, 0, 0};
      fe cur;
      fe_pow(&cur, &g, ee, &FR);
      for (size_t j = lo; j < hi; j++) {
        fe_mul(&a[j], &a[j], &cur, &FR);
        fe_mul(&cur, &cur, &g, &FR);
      }
    }
  }
  /* DIF stages. The first stages (butterfly span > one cache block) sweep the whole vector; once the span fits a
   * 2^14-element block (512 KB), every block finishes ALL its remaining stages while it is cache resident. */
  const int BLK_LOG = 14;
  int s_split = logn > BLK_LOG ? logn - BLK_LOG : 0;
  for (int s = 0; s < s_split; s++) {
    size_t half = n >> (s + 1);
    const fe* stw = ctw + (n - (n >> s));
#pragma omp parallel for schedule(static)
    for (size_t bf = 0; bf < n / 2; bf++) {
      size_t blk = bf / half, j = bf % half;
      size_t i0 = blk * 2 * half + j, i1 = i0 + half;
      fe u = a[i0], v = a[i1], d;
      fe_add(&a[i0], &u, &v, &FR);
      fe_sub(&d, &u, &v, &FR);
      fe_mul(&a[i1], &d, &stw[j], &FR);
    }
  }
  {
    size_t bsz = n >> s_split; /* block size */
#pragma omp parallel for schedule(static)
    for (size_t b0 = 0; b0 < n; b0 += bsz) {
      for (int s = s_split; s < logn; s++) {
        size_t half = n >> (s + 1);
        const fe* stw = ctw + (n - (n >> s));
        for (size_t g = b0; g < b0 + bsz; g += 2 * half)
          for (size_t j = 0; j < half; j++) {
            size_t i0 = g + j, i1 = i0 + half;
            fe u = a[i0], v = a[i1], d;
            fe_add(&a[i0], &u, &v, &FR);
            fe_sub(&d, &u, &v, &FR);
            fe_mul(&a[i1], &d, &stw[j], &FR);
          }
      }
    }
  }
#pragma omp parallel for schedule(static)
  for (size_t i = 0; i < n; i++) {
    size_t j = bitrev32((uint32_t)i, logn);
    if (i < j) {
      fe t = a[i];
      a[i] = a[j];
      a[j] = t;
    }
  }
  if (inverse) {
    fe ninv, ginv;
    fe_from_u64(&ninv, (uint64_t)n, &FR);
    fe_inv(&ninv, &ninv, &FR);
    fe_inv(&ginv, &g, &FR);
#pragma omp parallel
    {
      int nt = omp_get_num_threads(), t = omp_get_thread_num();
      size_t lo = n * t / nt, hi = n * (t + 1) / nt;
      uint64_t ee[4] = {lo, 0, 0, 0};
      fe cur = ninv;
      if (coset) {
        fe gp;
        fe_pow(&gp, &ginv, ee, &FR);
        fe_mul(&cur, &cur, &gp, &FR);
      }
      for (size_t j = lo; j < hi; j++) {
        fe_mul(&a[j], &a[j], &cur, &FR);
        if (coset) fe_mul(&cur, &cur, &ginv, &FR);
      }
    }
  }
  return 0;
}

/* Fr helpers for tests */
void ref_fr_mul(const uint64_t* a, const uint64_t* b, uint64_t* o, size_t n) {
  for (size_t i = 0; i < n; i++) fe_mul((fe*)(o + 4 * i), (const fe*)(a + 4 * i), (const fe*)(b + 4 * i), &FR);
}
int ref_max_threads(void) { return omp_get_max_threads(); }

/* ---- helpers for the full-size CPU run (oracle/c/wrap_cpu.cc) and the tests ---- */
static const g1_aff G1_GEN = {{{0xd35d438dc58f0d9dull, 0x0a78eb28f5c70b3dull, 0x666ea36f7879462cull, 0x0e0a77c19a07df2full}},   /* 1 */
                              {{0xa6ba871b8b1e1b3aull, 0x14f1d651eb8e167bull, 0xccdd46def0f28c58ull, 0x1c14ef83340fbe5eull}}};  /* 2 */

/* out[i] = [k0 + i] base  (affine, Montgomery), n points, by repeated addition + one inversion each (setup-only) */
void ref_g1_multiples(const uint64_t* base_or_null, uint64_t k0, size_t n, uint64_t* out) {
  g1_aff b = base_or_null ? *(const g1_aff*)base_or_null : G1_GEN;
  g1_xyzz cur;
  g1_set_inf(&cur);
  for (int bit = 63; bit >= 0; bit--) {
    g1_dbl(&cur, &cur);
    if ((k0 >> bit) & 1) g1_madd(&cur, &b, 0);
  }
  for (size_t i = 0; i < n; i++) {
    g1_to_affine((g1_aff*)(out + 8 * i), &cur);
    g1_madd(&cur, &b, 0);
  }
}
void ref_g2_multiples(const uint64_t* base, uint64_t k0, size_t n, uint64_t* out) {
  g2_aff b = *(const g2_aff*)base;
  g2_xyzz cur;
  g2_set_inf(&cur);
  for (int bit = 63; bit >= 0; bit--) {
    g2_dbl(&cur, &cur);
    if ((k0 >> bit) & 1) g2_madd(&cur, &b, 0);
  }
  for (size_t i = 0; i < n; i++) {
    g2_to_affine((g2_aff*)(out + 16 * i), &cur);
    g2_madd(&cur, &b, 0);
  }
}
/* out = [k] P, k canonical 256-bit (double-and-add) */
void ref_g1_scalar_mul(const uint64_t* point, const uint64_t* k, uint64_t* out) {
  g1_xyzz acc;
  g1_set_inf(&acc);
  for (int w = 3; w >= 0; w--)
    for (int bit = 63; bit >= 0; bit--) {
      g1_dbl(&acc, &acc);
      if ((k[w] >> bit) & 1) g1_madd(&acc, (const g1_aff*)point, 0);
    }
  g1_to_affine((g1_aff*)out, &acc);
}
void ref_g2_scalar_mul(const uint64_t* point, const uint64_t* k, uint64_t* out) {
  g2_xyzz acc;
  g2_set_inf(&acc);
  for (int w = 3; w >= 0; w--)
    for (int bit = 63; bit >= 0; bit--) {
      g2_dbl(&acc, &acc);
      if ((k[w] >> bit) & 1) g2_madd(&acc, (const g2_aff*)point, 0);
    }
  g2_to_affine((g2_aff*)out, &acc);
}
/* Fr: o = a + b, o = a - b (Montgomery or canonical alike), o = to/from Montgomery */
void ref_fr_add(const uint64_t* a, const uint64_t* b, uint64_t* o) { fe_add((fe*)o, (const fe*)a, (const fe*)b, &FR); }
void ref_fr_to_mont(const uint64_t* a, uint64_t* o, size_t n) {
  for (size_t i = 0; i < n; i++) fe_mul((fe*)(o + 4 * i), (const fe*)(a + 4 * i), &FR.r2, &FR);
}
void ref_fr_from_mont(const uint64_t* a, uint64_t* o, size_t n) {
  for (size_t i = 0; i < n; i++) fe_from_mont((fe*)(o + 4 * i), (const fe*)(a + 4 * i), &FR);
}
void ref_fp_to_mont(const uint64_t* a, uint64_t* o, size_t n) {
  for (size_t i = 0; i < n; i++) fe_mul((fe*)(o + 4 * i), (const fe*)(a + 4 * i), &FP.r2, &FP);
}
void ref_fp_from_mont(const uint64_t* a, uint64_t* o, size_t n) {
  for (size_t i = 0; i < n; i++) fe_from_mont((fe*)(o + 4 * i), (const fe*)(a + 4 * i), &FP);
}
