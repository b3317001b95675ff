/* ORACLE (test infrastructure only). Group law + Pippenger MSM "template": included twice by bn254_ref.c
 * with T / NAME / F_* bound to Fp (G1) and Fp2 (G2). Formulas: EFD xyzz madd-2008-s, add-2008-s,
 * dbl-2008-s-1 (a = 0) - the extended-Jacobian bucket representation gnark-crypto's MultiExp uses. */
typedef struct { T x, y; } NAME(aff);
typedef struct { T X, Y, ZZ, ZZZ; } NAME(xyzz);

static inline int NAME(aff_is_inf)(const NAME(aff)* p) { return F_ISZERO(&p->x) && F_ISZERO(&p->y); }
static inline void NAME(set_inf)(NAME(xyzz)* p) { memset(p, 0, sizeof(*p)); }
static inline int NAME(is_inf)(const NAME(xyzz)* p) { return F_ISZERO(&p->ZZ); }

static void NAME(dbl)(NAME(xyzz)* r, const NAME(xyzz)* p) {
  if (NAME(is_inf)(p)) { *r = *p; return; }
  T U, V, W, S, M, X2, X3, Y3, t;
  F_ADD(&U, &p->Y, &p->Y);
  F_MUL(&V, &U, &U);
  F_MUL(&W, &U, &V);
  F_MUL(&S, &p->X, &V);
  F_MUL(&X2, &p->X, &p->X);
  F_ADD(&M, &X2, &X2);
  F_ADD(&M, &M, &X2);
  F_MUL(&X3, &M, &M);
  F_SUB(&X3, &X3, &S);
  F_SUB(&X3, &X3, &S);
  F_SUB(&t, &S, &X3);
  F_MUL(&Y3, &M, &t);
  F_MUL(&t, &W, &p->Y);
  F_SUB(&Y3, &Y3, &t);
  NAME(xyzz) o;
  o.X = X3;
  o.Y = Y3;
  F_MUL(&o.ZZ, &V, &p->ZZ);
  F_MUL(&o.ZZZ, &W, &p->ZZZ);
  *r = o;
}

static void NAME(add)(NAME(xyzz)* acc, const NAME(xyzz)* q) {
  if (NAME(is_inf)(q)) return;
  if (NAME(is_inf)(acc)) { *acc = *q; return; }
  T U1, U2, S1, S2, P, R, PP, PPP, Q, X3, Y3, t;
  F_MUL(&U1, &acc->X, &q->ZZ);
  F_MUL(&U2, &q->X, &acc->ZZ);
  F_MUL(&S1, &acc->Y, &q->ZZZ);
  F_MUL(&S2, &q->Y, &acc->ZZZ);
  F_SUB(&P, &U2, &U1);
  F_SUB(&R, &S2, &S1);
  if (F_ISZERO(&P)) {
    if (F_ISZERO(&R)) NAME(dbl)(acc, acc); else NAME(set_inf)(acc);
    return;
  }
  F_MUL(&PP, &P, &P);
  F_MUL(&PPP, &P, &PP);
  F_MUL(&Q, &U1, &PP);
  F_MUL(&X3, &R, &R);
  F_SUB(&X3, &X3, &PPP);
  F_SUB(&X3, &X3, &Q);
  F_SUB(&X3, &X3, &Q);
  F_SUB(&t, &Q, &X3);
  F_MUL(&Y3, &R, &t);
  F_MUL(&t, &S1, &PPP);
  F_SUB(&Y3, &Y3, &t);
  acc->X = X3;
  acc->Y = Y3;
  F_MUL(&t, &acc->ZZ, &q->ZZ);
  F_MUL(&acc->ZZ, &t, &PP);
  F_MUL(&t, &acc->ZZZ, &q->ZZZ);
  F_MUL(&acc->ZZZ, &t, &PPP);
}

static void NAME(madd)(NAME(xyzz)* acc, const NAME(aff)* q, int negate) {
  if (NAME(aff_is_inf)(q)) return;
  T y2 = q->y;
  if (negate) F_NEG(&y2, &q->y);
  if (NAME(is_inf)(acc)) {
    acc->X = q->x;
    acc->Y = y2;
    F_SETONE(&acc->ZZ);
    F_SETONE(&acc->ZZZ);
    return;
  }
  T U2, S2, P, R, PP, PPP, Q, X3, Y3, t;
  F_MUL(&U2, &q->x, &acc->ZZ);
  F_MUL(&S2, &y2, &acc->ZZZ);
  F_SUB(&P, &U2, &acc->X);
  F_SUB(&R, &S2, &acc->Y);
  if (F_ISZERO(&P)) {
    if (F_ISZERO(&R)) NAME(dbl)(acc, acc); else NAME(set_inf)(acc);
    return;
  }
  F_MUL(&PP, &P, &P);
  F_MUL(&PPP, &P, &PP);
  F_MUL(&Q, &acc->X, &PP);
  F_MUL(&X3, &R, &R);
  F_SUB(&X3, &X3, &PPP);
  F_SUB(&X3, &X3, &Q);
  F_SUB(&X3, &X3, &Q);
  F_SUB(&t, &Q, &X3);
  F_MUL(&Y3, &R, &t);
  F_MUL(&t, &acc->Y, &PPP);
  F_SUB(&Y3, &Y3, &t);
  acc->X = X3;
  acc->Y = Y3;
  F_MUL(&acc->ZZ, &acc->ZZ, &PP);
  F_MUL(&acc->ZZZ, &acc->ZZZ, &PPP);
}

static void NAME(to_affine)(NAME(aff)* r, const NAME(xyzz)* p) {
  if (NAME(is_inf)(p)) { memset(r, 0, sizeof(*r)); return; }
  T i3, i2, t;
  F_INV(&i3, &p->ZZZ);
  F_MUL(&t, &i3, &p->ZZ);
  F_MUL(&i2, &t, &t);
  F_MUL(&r->x, &p->X, &i2);
  F_MUL(&r->y, &p->Y, &i3);
}

static inline uint32_t NAME(bits)(const uint64_t* s, int lo, int c) {
  int w = lo >> 6, off = lo & 63;
  if (w >= 4) return 0;
  uint64_t v = s[w] >> off;
  if (off && w + 1 < 4) v |= s[w + 1] << (64 - off);
  return (uint32_t)(v & ((1ull << c) - 1));
}

/* Pippenger, signed digits; tasks = windows x point-ranges so that all cores are busy */
static int NAME(msm)(const fe* scalars, const NAME(aff)* points, size_t n, int mont, int c, int nthreads,
                     NAME(aff)* out) {
  if (c <= 0) {
    int lg = 0;
    while (((size_t)1 << (lg + 1)) <= (n ? n : 1)) lg++;
    c = lg - 3;
    if (c > 16) c = 16;
    if (c < 4) c = 4;
  }
  if (c > 16) return -1;
  if (nthreads <= 0) nthreads = omp_get_max_threads();
  int nwin = (254 + c) / c;
  uint32_t half = 1u << (c - 1);
  /* signed digits for every scalar: int32 digits[n][nwin] */
  int32_t* digits = (int32_t*)malloc(n * (size_t)nwin * sizeof(int32_t) + 16);
  if (!digits) return -2;
#pragma omp parallel for num_threads(nthreads) schedule(static)
  for (size_t i = 0; i < n; i++) {
    fe s = scalars[i];
    if (mont) fe_from_mont(&s, &s, &FR);
    uint32_t carry = 0;
    for (int w = 0; w < nwin; w++) {
      uint32_t raw = NAME(bits)(s.l, w * c, c) + carry;
      if (raw > half) { digits[i * nwin + w] = (int32_t)raw - (int32_t)(1u << c); carry = 1; }
      else { digits[i * nwin + w] = (int32_t)raw; carry = 0; }
    }
  }
  int splits = (nthreads + nwin - 1) / nwin;
  if (splits < 1) splits = 1;
  int ntasks = nwin * splits;
  NAME(xyzz)* partial = (NAME(xyzz)*)malloc((size_t)ntasks * sizeof(NAME(xyzz)));
  int fail = 0;
#pragma omp parallel for num_threads(nthreads) schedule(dynamic, 1)
  for (int task = 0; task < ntasks; task++) {
    int w = task / splits, sp = task % splits;
    size_t lo = n * sp / splits, hi = n * (sp + 1) / splits;
    NAME(xyzz)* buckets = (NAME(xyzz)*)calloc(half, sizeof(NAME(xyzz)));
    if (!buckets) { fail = 1; continue; }
    for (size_t i = lo; i < hi; i++) {
      int32_t d = digits[i * nwin + w];
      if (d > 0) NAME(madd)(&buckets[d - 1], &points[i], 0);
      else if (d < 0) NAME(madd)(&buckets[-d - 1], &points[i], 1);
    }
    NAME(xyzz) run, sum;
    NAME(set_inf)(&run);
    NAME(set_inf)(&sum);
    for (int b = (int)half - 1; b >= 0; b--) {
      NAME(add)(&run, &buckets[b]);
      NAME(add)(&sum, &run);
    }
    partial[task] = sum;
    free(buckets);
  }
  free(digits);
  if (fail) { free(partial); return -2; }
  NAME(xyzz) res;
  NAME(set_inf)(&res);
  for (int w = nwin - 1; w >= 0; w--) {
    for (int k = 0; k < c; k++) NAME(dbl)(&res, &res);
    for (int sp = 0; sp < splits; sp++) NAME(add)(&res, &partial[w * splits + sp]);
  }
  free(partial);
  NAME(to_affine)(out, &res);
  return 0;
}
