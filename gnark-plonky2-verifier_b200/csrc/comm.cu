// One large MSM over several GPUs (BASELINE.json configs[4]): NCCL communicator per context and the sharded
// MSM entry points. The reference has no multi-GPU path (gnark-crypto's MultiExp splits its windows over CPU cores,
// reached from groth16.Prove at benchmark.go:249); this is that split across B200s.
//
//   split = windows: every rank holds all n scalars and bases and computes the Pippenger windows [lo, hi) it owns
//                    (digits of the other windows are not even sorted), result sum_{w in range} 2^(c w) W_w;
//   split = points:  rank r takes the points [n r / N, n (r + 1) / N) and runs a complete MSM over them.
// Either way ONE all-gather of a single affine point per rank (64 B G1 / 128 B G2) on the context's stream over
// NCCL / NVLink, then the N points are added on the device (EC addition is not an NCCL reduction operator). The result
// is the unique group element, bit-identical to the single-GPU MSM, on every rank.
//
// libnccl is loaded lazily (dlopen "libnccl.so.2": the copy torch has already loaded when the host is a torch process,
// the system one otherwise), so libgpw.so itself has no link-time dependency on NCCL; GPW_ENCCL if it cannot be had.
#include <dlfcn.h>
#include <nccl.h>

#include <mutex>

#include "common.cuh"
#include "ec.cuh"

extern "C" {
int gpw_msm_g1_dev(gpw_ctx* ctx, uint64_t s, uint64_t p, size_t n, int mont, int c, int lo, int hi, uint64_t* out);
int gpw_msm_g2_dev(gpw_ctx* ctx, uint64_t s, uint64_t p, size_t n, int mont, int c, int lo, int hi, uint64_t* out);
}

namespace gpw {

struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  ncclResult_t (*GetVersion)(int*) = nullptr;
};

static NcclApi* nccl_api() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
      api.lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
      if (api.lib) break;
    }
    if (!api.lib) return;
    api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(api.lib, "ncclGetUniqueId");
    api.CommInitRank = (decltype(api.CommInitRank))dlsym(api.lib, "ncclCommInitRank");
    api.CommDestroy = (decltype(api.CommDestroy))dlsym(api.lib, "ncclCommDestroy");
    api.AllGather = (decltype(api.AllGather))dlsym(api.lib, "ncclAllGather");
    api.GetErrorString = (decltype(api.GetErrorString))dlsym(api.lib, "ncclGetErrorString");
    api.GetVersion = (decltype(api.GetVersion))dlsym(api.lib, "ncclGetVersion");
    if (!api.GetUniqueId || !api.CommInitRank || !api.CommDestroy || !api.AllGather || !api.GetErrorString) api.lib = nullptr;
  });
  return api.lib ? &api : nullptr;
}

#define GPW_NCCL(expr)                                                                                   \
  do {                                                                                                   \
    ncclResult_t _r = (expr);                                                                            \
    if (_r != ncclSuccess) {                                                                             \
      gpw::set_error("NCCL error at %s:%d: %s", __FILE__, __LINE__, gpw::nccl_api()->GetErrorString(_r)); \
      return GPW_ENCCL;                                                                                  \
    }                                                                                                    \
  } while (0)

// out = sum of the n affine points (one thread: n is the number of ranks)
template <class F>
__global__ void k_ec_sum_points(const Affine<F>* __restrict__ pts, int n, Affine<F>* __restrict__ out) {
  if (blockIdx.x || threadIdx.x) return;
  XYZZ<F> acc = XYZZ<F>::inf();
  for (int i = 0; i < n; i++) add_mixed(acc, pts[i], false);
  *out = to_affine(acc);
}

// windows [lo, hi) of rank r out of N for nwin windows: contiguous, the first (nwin mod N) ranks get one more
static void window_range(int nwin, int N, int r, int* lo, int* hi) {
  const int base = nwin / N, extra = nwin % N;
  *lo = r * base + (r < extra ? r : extra);
  *hi = *lo + base + (r < extra ? 1 : 0);
}

static int pick_window(size_t n) {  // same rule as msm_dev_impl's choose_window
  int lg = 0;
  while ((1ull << (lg + 1)) <= n) lg++;
  int c = lg - 3;
  return c > 16 ? 16 : c < 4 ? 4 : c;
}

// This rank's share of the MSM as one affine point (host). virt_rank / virt_n: the communicator's rank and size, or - for
// the single-GPU tests of the splitting logic - any (rank, size) pair.
template <class F>
static int sharded_partial(gpw_ctx* ctx, uint64_t scalars_dev, uint64_t points_dev, size_t n, int mont, int c, int split, int rank,
                           int N, uint64_t* out_affine) {
  constexpr int WORDS = (int)(sizeof(Affine<F>) / 8);
  auto msm = sizeof(F) == sizeof(Fp) ? gpw_msm_g1_dev : gpw_msm_g2_dev;
  for (int i = 0; i < WORDS; i++) out_affine[i] = 0;
  if (n == 0) return GPW_OK;
  if (split == 1) {
    if (c == 0) c = pick_window(n);
    const int nwin = (254 + c) / c;
    int lo, hi;
    window_range(nwin, N, rank, &lo, &hi);
    if (lo >= hi) return GPW_OK;  // more ranks than windows: this one contributes the point at infinity
    return msm(ctx, scalars_dev, points_dev, n, mont, c, lo, hi, out_affine);
  }
  const size_t lo = n * (size_t)rank / N, hi = n * (size_t)(rank + 1) / N;
  if (lo >= hi) return GPW_OK;
  return msm(ctx, scalars_dev + lo * sizeof(Fr), points_dev + lo * sizeof(Affine<F>), hi - lo, mont, c, 0, 0, out_affine);
}

template <class F>
static int msm_sharded_impl(gpw_ctx* ctx, uint64_t scalars_dev, uint64_t points_dev, size_t n, int mont, int c, int split,
                            uint64_t* out_affine) {
  constexpr int WORDS = (int)(sizeof(Affine<F>) / 8);
  if (!ctx || !out_affine || (n && (!scalars_dev || !points_dev)) || split < 0 || split > 2) {
    set_error("msm_sharded: bad argument");
    return GPW_EINVAL;
  }
  if (!ctx->nccl_comm) {
    set_error("msm_sharded: no communicator on this context (gpw_comm_init first)");
    return GPW_ENCCL;
  }
  NcclApi* api = nccl_api();
  const int N = ctx->comm_size, rank = ctx->comm_rank;
  // auto: the window split while every rank gets at least two windows (its sort and bucket reduction shrink with the share),
  // the point split beyond that
  if (split == 0) split = ((254 + (c ? c : pick_window(n ? n : 1))) / (c ? c : pick_window(n ? n : 1))) >= 2 * N ? 1 : 2;
  GPW_CUDA(cudaSetDevice(ctx->device));
  uint64_t mine[WORDS];
  GPW_TRY(sharded_partial<F>(ctx, scalars_dev, points_dev, n, mont, c, split, rank, N, mine));
  Affine<F>* buf = nullptr;  // [0, N): gathered partials, [N]: mine, [N + 1]: the sum
  GPW_TRY(ctx->get_scratch("comm.gather", (size_t)(N + 2) * sizeof(Affine<F>), (void**)&buf));
  uint64_t* stage = (uint64_t*)ctx->pin_take(sizeof(Affine<F>));
  memcpy(stage, mine, sizeof(Affine<F>));
  GPW_CUDA(cudaMemcpyAsync(buf + N, stage, sizeof(Affine<F>), cudaMemcpyHostToDevice, ctx->stream));
  GPW_NCCL(api->AllGather(buf + N, buf, sizeof(Affine<F>), ncclUint8, (ncclComm_t)ctx->nccl_comm, ctx->stream));
  k_ec_sum_points<F><<<1, 32, 0, ctx->stream>>>(buf, N, buf + N + 1);
  GPW_CHECK_LAUNCH();
  ctx->launches += 1;
  uint64_t* res = (uint64_t*)ctx->pin_take(sizeof(Affine<F>));
  GPW_CUDA(cudaMemcpyAsync(res, buf + N + 1, sizeof(Affine<F>), cudaMemcpyDeviceToHost, ctx->stream));
  GPW_CUDA(cudaStreamSynchronize(ctx->stream));
  memcpy(out_affine, res, sizeof(Affine<F>));
  return GPW_OK;
}

}  // namespace gpw

using namespace gpw;

// ncclGetUniqueId: rank 0 calls this and hands the 128 bytes to every rank through whatever the host program uses to
// talk between its processes (torch.distributed broadcast in bench.py, MPI, the Go side's own RPC).
extern "C" int gpw_comm_unique_id(uint8_t* out128) {
  if (!out128) return GPW_EINVAL;
  NcclApi* api = nccl_api();
  if (!api) {
    set_error("libnccl.so.2 could not be loaded: %s", dlerror());
    return GPW_ENCCL;
  }
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  ncclUniqueId id;
  GPW_NCCL(api->GetUniqueId(&id));
  memcpy(out128, &id, 128);
  return GPW_OK;
}

// ncclCommInitRank on the context's device. Collective: every rank of the job calls it with the same id.
extern "C" int gpw_comm_init(gpw_ctx* ctx, int nranks, int rank, const uint8_t* id128) {
  if (!ctx || !id128 || nranks < 1 || rank < 0 || rank >= nranks) {
    set_error("comm_init: bad argument");
    return GPW_EINVAL;
  }
  NcclApi* api = nccl_api();
  if (!api) {
    set_error("libnccl.so.2 could not be loaded: %s", dlerror());
    return GPW_ENCCL;
  }
  if (ctx->nccl_comm) {
    set_error("comm_init: the context already has a communicator");
    return GPW_EINVAL;
  }
  GPW_CUDA(cudaSetDevice(ctx->device));
  ncclUniqueId id;
  memcpy(&id, id128, 128);
  ncclComm_t comm;
  GPW_NCCL(api->CommInitRank(&comm, nranks, id, rank));
  ctx->nccl_comm = comm;
  ctx->comm_size = nranks;
  ctx->comm_rank = rank;
  return GPW_OK;
}

extern "C" int gpw_comm_destroy(gpw_ctx* ctx) {
  if (!ctx) return GPW_EINVAL;
  if (!ctx->nccl_comm) return GPW_OK;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  NcclApi* api = nccl_api();
  if (api) api->CommDestroy((ncclComm_t)ctx->nccl_comm);
  ctx->nccl_comm = nullptr;
  ctx->comm_size = 1;
  ctx->comm_rank = 0;
  return GPW_OK;
}

// info3: {ranks, this rank, NCCL version code (0 if unknown)}
extern "C" int gpw_comm_info(const gpw_ctx* ctx, int* info3) {
  if (!ctx || !info3) return GPW_EINVAL;
  info3[0] = ctx->nccl_comm ? ctx->comm_size : 0;
  info3[1] = ctx->comm_rank;
  int v = 0;
  NcclApi* api = nccl_api();
  if (api && api->GetVersion) api->GetVersion(&v);
  info3[2] = v;
  return GPW_OK;
}

extern "C" int gpw_msm_g1_sharded(gpw_ctx* ctx, uint64_t scalars_dev, uint64_t points_dev, size_t n, int scalars_mont, int window_bits,
                                  int split, uint64_t* out_affine) {
  return msm_sharded_impl<Fp>(ctx, scalars_dev, points_dev, n, scalars_mont, window_bits, split, out_affine);
}
extern "C" int gpw_msm_g2_sharded(gpw_ctx* ctx, uint64_t scalars_dev, uint64_t points_dev, size_t n, int scalars_mont, int window_bits,
                                  int split, uint64_t* out_affine) {
  return msm_sharded_impl<Fp2>(ctx, scalars_dev, points_dev, n, scalars_mont, window_bits, split, out_affine);
}

// The share rank `rank` of `nranks` would contribute, computed on THIS context without any communicator: lets one GPU
// check that the shares of a split add up to the whole MSM (tests) and lets a host-side scheduler run shares wherever it likes.
extern "C" int gpw_msm_sharded_partial(gpw_ctx* ctx, int group, uint64_t scalars_dev, uint64_t points_dev, size_t n, int scalars_mont,
                                       int window_bits, int split, int rank, int nranks, uint64_t* out_affine) {
  if (!ctx || !out_affine || nranks < 1 || rank < 0 || rank >= nranks || (split != 1 && split != 2) || (group != 1 && group != 2)) {
    set_error("msm_sharded_partial: bad argument");
    return GPW_EINVAL;
  }
  GPW_CUDA(cudaSetDevice(ctx->device));
  if (group == 1) return sharded_partial<Fp>(ctx, scalars_dev, points_dev, n, scalars_mont, window_bits, split, rank, nranks, out_affine);
  return sharded_partial<Fp2>(ctx, scalars_dev, points_dev, n, scalars_mont, window_bits, split, rank, nranks, out_affine);
}
