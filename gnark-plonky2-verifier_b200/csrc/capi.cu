// Context management, error plumbing, host-side arithmetic entry points and the GPU field self-test.
#include <cstdarg>
#include <cstring>
#include <random>

#include "common.cuh"
#include "ec.cuh"
#include "host_ec.cuh"

namespace gpw {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

}  // namespace gpw

using namespace gpw;

int gpw_ctx::get_scratch(const char* name, size_t bytes, void** out) {
  Scratch& s = scratch[name];
  if (s.cap < bytes) {
    if (s.p) {
      GPW_CUDA(cudaStreamSynchronize(stream));
      GPW_CUDA(cudaFree(s.p));
      s.p = nullptr;
      s.cap = 0;
    }
    size_t want = bytes + bytes / 8 + 256;
    cudaError_t e = cudaMalloc(&s.p, want);
    if (e != cudaSuccess) {
      set_error("cudaMalloc(%zu bytes) for scratch '%s' failed: %s", want, name, cudaGetErrorString(e));
      s.p = nullptr;
      return GPW_ENOMEM;
    }
    s.cap = want;
  }
  *out = s.p;
  return GPW_OK;
}

extern "C" int gpw_version(void) { return 100; }

extern "C" const char* gpw_last_error(void) { return g_err; }

extern "C" int gpw_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}

extern "C" int gpw_ctx_create(int device, gpw_ctx** out) {
  if (!out) {
    set_error("ctx_create: null out");
    return GPW_EINVAL;
  }
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    set_error("no usable CUDA device (%s); libgpw has no CPU fallback", e == cudaSuccess ? "count = 0" : cudaGetErrorString(e));
    return GPW_ENODEV;
  }
  if (device < 0 || device >= n) {
    set_error("device %d out of range (have %d)", device, n);
    return GPW_EINVAL;
  }
  GPW_CUDA(cudaSetDevice(device));
  gpw_ctx* c = new gpw_ctx();
  c->device = device;
  GPW_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  {
    int least = 0, greatest = 0;
    GPW_CUDA(cudaDeviceGetStreamPriorityRange(&least, &greatest));
    GPW_CUDA(cudaStreamCreateWithPriority(&c->stream_hi, cudaStreamNonBlocking, greatest));
    GPW_CUDA(cudaEventCreateWithFlags(&c->ev_hop, cudaEventDisableTiming));
    for (int i = 0; i < 2; i++) GPW_CUDA(cudaEventCreateWithFlags(&c->slot_done[i], cudaEventDisableTiming));
    for (auto& p : c->pend)
      for (auto& e : p.ev) GPW_CUDA(cudaEventCreate(&e));
  }
  cudaDeviceProp prop;
  GPW_CUDA(cudaGetDeviceProperties(&prop, device));
  c->sm_count = prop.multiProcessorCount;
  for (int i = 0; i < 4; i++) GPW_CUDA(cudaEventCreate(&c->ev[i]));
  GPW_CUDA(cudaHostAlloc((void**)&c->pin, gpw_ctx::PIN_CAP, cudaHostAllocDefault));
  *out = c;
  return GPW_OK;
}

extern "C" int gpw_comm_destroy(gpw_ctx* ctx);
extern "C" void gpw_ctx_destroy(gpw_ctx* ctx) {
  if (!ctx) return;
  gpw_comm_destroy(ctx);
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  if (ctx->stream_hi) {
    cudaStreamSynchronize(ctx->stream_hi);
    cudaStreamDestroy(ctx->stream_hi);
  }
  if (ctx->ev_hop) cudaEventDestroy(ctx->ev_hop);
  for (auto& e : ctx->slot_done)
    if (e) cudaEventDestroy(e);
  for (auto& p : ctx->pend)
    for (auto& e : p.ev)
      if (e) cudaEventDestroy(e);
  for (auto& kv : ctx->scratch)
    if (kv.second.p) cudaFree(kv.second.p);
  for (auto& kv : ctx->ntt) {
    if (kv.second.shared) continue;
    cudaFree(kv.second.tw);
    cudaFree(kv.second.coset);
    cudaFree(kv.second.coset_inv);
  }
  for (int i = 0; i < 4; i++)
    if (ctx->ev[i]) cudaEventDestroy(ctx->ev[i]);
  if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
  if (ctx->pin) cudaFreeHost(ctx->pin);
  delete ctx;
}

extern "C" int gpw_ctx_set_stream(gpw_ctx* ctx, void* cuda_stream) {
  if (!ctx) return GPW_EINVAL;
  GPW_CUDA(cudaSetDevice(ctx->device));
  GPW_CUDA(cudaStreamSynchronize(ctx->stream));
  if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
  if (cuda_stream) {
    ctx->stream = (cudaStream_t)cuda_stream;
    ctx->own_stream = false;
  } else {
    GPW_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    ctx->own_stream = true;
  }
  return GPW_OK;
}

// Tunables of a context. "msm_affine_rounds": batch-affine pair-reduction rounds before the XYZZ bucket accumulation
// (0 = off, the default; see csrc/msm_affine.cuh).
extern "C" int gpw_ctx_set_option(gpw_ctx* ctx, const char* key, int64_t value) {
  if (!ctx || !key) return GPW_EINVAL;
  if (!strcmp(key, "msm_affine_rounds")) {
    if (value < 0 || value > 8) {
      set_error("msm_affine_rounds must be in [0, 8]");
      return GPW_EINVAL;
    }
    ctx->msm_affine_rounds = (int)value;
    return GPW_OK;
  }
  if (!strcmp(key, "msm_overlap")) {  // the wrap prover's deferred MSMs (common.cuh): -1 automatic (default), 0 off, 1 on
    if (value < -1 || value > 1) {
      set_error("msm_overlap must be -1, 0 or 1");
      return GPW_EINVAL;
    }
    ctx->msm_overlap_mode = (int)value;
    return GPW_OK;
  }
  set_error("ctx_set_option: unknown option '%s'", key);
  return GPW_EINVAL;
}

extern "C" int gpw_ctx_sync(gpw_ctx* ctx) {
  if (!ctx) return GPW_EINVAL;
  GPW_CUDA(cudaSetDevice(ctx->device));
  GPW_CUDA(cudaStreamSynchronize(ctx->stream));
  return GPW_OK;
}

extern "C" uint64_t gpw_ctx_launch_count(const gpw_ctx* ctx) { return ctx ? ctx->launches : 0; }

// ---- host arithmetic ------------------------------------------------------------------------------
template <class P>
static void host_mul(int impl, const uint64_t* a, const uint64_t* b, uint64_t* o, size_t n) {
  for (size_t i = 0; i < n; i++) {
    Fe<P> x, y;
    memcpy(&x, a + 4 * i, 32);
    memcpy(&y, b + 4 * i, 32);
    // impl 0: even/odd IMAD.WIDE schedule (host emulation), 1: plain CIOS, 2: dedicated squaring of a (b ignored)
    Fe<P> z = impl == 0 ? mont_mul_wide(x, y) : impl == 2 ? mont_sqr_wide(x) : mont_mul_portable(x, y);
    memcpy(o + 4 * i, &z, 32);
  }
}

extern "C" int gpw_host_ff_mul(int field, int impl, const uint64_t* a, const uint64_t* b, uint64_t* out, size_t n) {
  if ((!a || !b || !out) && n) return GPW_EINVAL;
  if (field == 0) host_mul<FrParams>(impl, a, b, out, n);
  else if (field == 1) host_mul<FpParams>(impl, a, b, out, n);
  else return GPW_EINVAL;
  return GPW_OK;
}

// out = a b - c d through mont_mul2 (host emulation of the schedule the GPU runs): one reduction for both products
template <class P>
static void host_mul_sub2(const uint64_t* a, const uint64_t* b, const uint64_t* c, const uint64_t* d, uint64_t* out, size_t n) {
  for (size_t i = 0; i < n; i++) {
    Fe<P> x, y, z, w;
    memcpy(&x, a + 4 * i, 32);
    memcpy(&y, b + 4 * i, 32);
    memcpy(&z, c + 4 * i, 32);
    memcpy(&w, d + 4 * i, 32);
    Fe<P> r = mul_sub2(x, y, z, w);
    memcpy(out + 4 * i, &r, 32);
  }
}

extern "C" int gpw_host_ff_mul_sub2(int field, const uint64_t* a, const uint64_t* b, const uint64_t* c, const uint64_t* d, uint64_t* out,
                                    size_t n) {
  if ((!a || !b || !c || !d || !out) && n) return GPW_EINVAL;
  if (field == 0) host_mul_sub2<FrParams>(a, b, c, d, out, n);
  else if (field == 1) host_mul_sub2<FpParams>(a, b, c, d, out, n);
  else return GPW_EINVAL;
  return GPW_OK;
}

template <class P, class Fn>
static void host_map(const uint64_t* a, uint64_t* o, size_t n, Fn f) {
  for (size_t i = 0; i < n; i++) {
    Fe<P> x;
    memcpy(&x, a + 4 * i, 32);
    Fe<P> z = f(x);
    memcpy(o + 4 * i, &z, 32);
  }
}

extern "C" int gpw_host_ff_to_mont(int field, const uint64_t* a, uint64_t* out, size_t n) {
  if ((!a || !out) && n) return GPW_EINVAL;
  if (field == 0) host_map<FrParams>(a, out, n, [](const Fr& x) { return to_mont(x); });
  else if (field == 1) host_map<FpParams>(a, out, n, [](const Fp& x) { return to_mont(x); });
  else return GPW_EINVAL;
  return GPW_OK;
}

extern "C" int gpw_host_ff_from_mont(int field, const uint64_t* a, uint64_t* out, size_t n) {
  if ((!a || !out) && n) return GPW_EINVAL;
  if (field == 0) host_map<FrParams>(a, out, n, [](const Fr& x) { return from_mont(x); });
  else if (field == 1) host_map<FpParams>(a, out, n, [](const Fp& x) { return from_mont(x); });
  else return GPW_EINVAL;
  return GPW_OK;
}

extern "C" int gpw_host_ff_inv(int field, const uint64_t* a, uint64_t* out, size_t n) {
  if ((!a || !out) && n) return GPW_EINVAL;
  if (field == 0) host_map<FrParams>(a, out, n, [](const Fr& x) { return inv(x); });
  else if (field == 1) host_map<FpParams>(a, out, n, [](const Fp& x) { return inv(x); });
  else return GPW_EINVAL;
  return GPW_OK;
}

// the shift-and-subtract inverse the batch-affine bucket accumulation uses on the device (ff.cuh inv_euclid)
extern "C" int gpw_host_ff_inv_euclid(int field, const uint64_t* a, uint64_t* out, size_t n) {
  if ((!a || !out) && n) return GPW_EINVAL;
  if (field == 0) host_map<FrParams>(a, out, n, [](const Fr& x) { return inv_euclid(x); });
  else if (field == 1) host_map<FpParams>(a, out, n, [](const Fp& x) { return inv_euclid(x); });
  else return GPW_EINVAL;
  return GPW_OK;
}

template <class F>
static int ec_scalar_mul_t(const uint64_t* p, const uint64_t* k, uint64_t* out) {
  Affine<F> a;
  memcpy(&a, p, sizeof(a));
  uint32_t kw[8];
  memcpy(kw, k, 32);
  Affine<F> r = to_affine(host_scalar_mul(a, kw));
  memcpy(out, &r, sizeof(r));
  return GPW_OK;
}

extern "C" int gpw_host_ec_scalar_mul(int group, const uint64_t* p, const uint64_t* k, uint64_t* out) {
  if (!p || !k || !out) return GPW_EINVAL;
  if (group == 1) return ec_scalar_mul_t<Fp>(p, k, out);
  if (group == 2) return ec_scalar_mul_t<Fp2>(p, k, out);
  return GPW_EINVAL;
}

template <class F>
static int ec_add_t(const uint64_t* p, const uint64_t* q, uint64_t* out) {
  Affine<F> a, b;
  memcpy(&a, p, sizeof(a));
  memcpy(&b, q, sizeof(b));
  XYZZ<F> r = XYZZ<F>::from_affine(a);
  add_mixed(r, b, false);
  Affine<F> o = to_affine(r);
  memcpy(out, &o, sizeof(o));
  return GPW_OK;
}

extern "C" int gpw_host_ec_add(int group, const uint64_t* p, const uint64_t* q, uint64_t* out) {
  if (!p || !q || !out) return GPW_EINVAL;
  if (group == 1) return ec_add_t<Fp>(p, q, out);
  if (group == 2) return ec_add_t<Fp2>(p, q, out);
  return GPW_EINVAL;
}

extern "C" int gpw_host_ec_is_on_curve(int group, const uint64_t* p) {
  if (!p) return GPW_EINVAL;
  if (group == 1) {
    G1Affine a;
    memcpy(&a, p, sizeof(a));
    if (a.is_inf()) return 1;
    Fp lhs = sqr(a.y), rhs = add(mul(sqr(a.x), a.x), fp_from_u64(3));
    return lhs == rhs ? 1 : 0;
  }
  if (group == 2) {
    G2Affine a;
    memcpy(&a, p, sizeof(a));
    if (a.is_inf()) return 1;
    Fp2 lhs = sqr(a.y), rhs = add(mul(sqr(a.x), a.x), g2_b());
    return lhs == rhs ? 1 : 0;
  }
  return GPW_EINVAL;
}

template <class F>
static int gen_multiples_t(uint64_t k0, size_t n, uint64_t* out) {
  Affine<F> g = generator<F>();
  uint32_t kw[8] = {(uint32_t)k0, (uint32_t)(k0 >> 32), 0, 0, 0, 0, 0, 0};
  XYZZ<F> cur = host_scalar_mul(g, kw);
  std::vector<XYZZ<F>> pts(n);
  for (size_t i = 0; i < n; i++) {
    pts[i] = cur;
    add_mixed(cur, g, false);
  }
  batch_to_affine(pts, reinterpret_cast<Affine<F>*>(out));
  return GPW_OK;
}

extern "C" int gpw_host_ec_generator_multiples(int group, uint64_t k0, size_t n, uint64_t* out) {
  if (!out && n) return GPW_EINVAL;
  if (group == 1) return gen_multiples_t<Fp>(k0, n, out);
  if (group == 2) return gen_multiples_t<Fp2>(k0, n, out);
  return GPW_EINVAL;
}

// ---- GPU self-test ----------------------------------------------------------------------------------
template <class P>
__global__ void k_selftest_mul(const Fe<P>* a, const Fe<P>* b, Fe<P>* o_wide, Fe<P>* o_port, Fe<P>* o_addsub, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  o_wide[i] = mont_mul_wide(a[i], b[i]);
  o_port[i] = mont_mul_portable(a[i], b[i]);
  o_addsub[i] = sub(add(a[i], b[i]), neg(b[i]));  // a + 2b
  // dedicated squaring against the general multiplication (both operands): a mismatch poisons the wide result
  if (mont_sqr_wide(a[i]) != mont_mul_wide(a[i], a[i]) || mont_sqr_wide(b[i]) != mont_mul_wide(b[i], b[i])) o_wide[i] = Fe<P>::zero();
}

template <class P>
static int selftest_field(gpw_ctx* ctx, size_t n, uint64_t seed, const char* name) {
  std::mt19937_64 rng(seed);
  std::vector<Fe<P>> a(n), b(n), w(n), p(n), s(n);
  for (size_t i = 0; i < n; i++) {
    for (int k = 0; k < 8; k++) {
      a[i].l[k] = (uint32_t)rng();
      b[i].l[k] = (uint32_t)rng();
    }
    a[i].l[7] &= 0x0fffffffu;  // < 2^252 < modulus
    b[i].l[7] &= 0x0fffffffu;
    if (i == 0) a[i] = Fe<P>::zero();
    if (i == 1) { a[i] = modulus<P>(); a[i].l[0] -= 1; b[i] = a[i]; }  // (p-1)^2
    if (i == 2) a[i] = Fe<P>::one();
  }
  Fe<P>*da, *db, *dw, *dp, *ds;
  size_t bytes = n * sizeof(Fe<P>);
  GPW_TRY(ctx->get_scratch("st.a", bytes, (void**)&da));
  GPW_TRY(ctx->get_scratch("st.b", bytes, (void**)&db));
  GPW_TRY(ctx->get_scratch("st.w", bytes, (void**)&dw));
  GPW_TRY(ctx->get_scratch("st.p", bytes, (void**)&dp));
  GPW_TRY(ctx->get_scratch("st.s", bytes, (void**)&ds));
  GPW_CUDA(cudaMemcpyAsync(da, a.data(), bytes, cudaMemcpyHostToDevice, ctx->stream));
  GPW_CUDA(cudaMemcpyAsync(db, b.data(), bytes, cudaMemcpyHostToDevice, ctx->stream));
  k_selftest_mul<P><<<div_up(n, 128), 128, 0, ctx->stream>>>(da, db, dw, dp, ds, n);
  GPW_CHECK_LAUNCH();
  ctx->launches++;
  GPW_CUDA(cudaMemcpyAsync(w.data(), dw, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  GPW_CUDA(cudaMemcpyAsync(p.data(), dp, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  GPW_CUDA(cudaMemcpyAsync(s.data(), ds, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  GPW_CUDA(cudaStreamSynchronize(ctx->stream));
  for (size_t i = 0; i < n; i++) {
    Fe<P> ref = mont_mul_portable(a[i], b[i]);
    Fe<P> ras = add(add(a[i], b[i]), b[i]);
    if (w[i] != ref || p[i] != ref || s[i] != ras) {
      set_error("selftest_ff(%s): mismatch at %zu (wide %s, portable %s, addsub %s)", name, i,
                w[i] == ref ? "ok" : "BAD", p[i] == ref ? "ok" : "BAD", s[i] == ras ? "ok" : "BAD");
      return GPW_ECUDA;
    }
  }
  return GPW_OK;
}

extern "C" int gpw_selftest_ff(gpw_ctx* ctx, size_t n, uint64_t seed) {
  if (!ctx) return GPW_EINVAL;
  GPW_CUDA(cudaSetDevice(ctx->device));
  GPW_TRY(selftest_field<FrParams>(ctx, n, seed, "Fr"));
  GPW_TRY(selftest_field<FpParams>(ctx, n, seed + 1, "Fp"));
  return GPW_OK;
}
