// Context, error plumbing and scratch-memory management shared by all libgpw translation units.
#pragma once
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <map>
#include <string>
#include <vector>

#include "../../include/gpw.h"

namespace gpw {

void set_error(const char* fmt, ...);

#define GPW_CUDA(expr)                                                                        \
  do {                                                                                        \
    cudaError_t _e = (expr);                                                                  \
    if (_e != cudaSuccess) {                                                                  \
      gpw::set_error("CUDA error %s at %s:%d: %s", cudaGetErrorName(_e), __FILE__, __LINE__, \
                     cudaGetErrorString(_e));                                                 \
      return GPW_ECUDA;                                                                       \
    }                                                                                         \
  } while (0)

#define GPW_CHECK_LAUNCH() GPW_CUDA(cudaGetLastError())

#define GPW_TRY(expr)          \
  do {                         \
    int _r = (expr);           \
    if (_r != GPW_OK) return _r; \
  } while (0)

// Grow-only device scratch buffer (one per named slot) so steady-state calls allocate nothing.
struct Scratch {
  void* p = nullptr;
  size_t cap = 0;
};

struct NttTables {
  void* tw = nullptr;         // w^e, e < N/2
  void* coset = nullptr;      // g^j, j < N
  void* coset_inv = nullptr;  // g^-j / N, j < N
  bool shared = false;        // borrowed from another context of the same device (gpw_ntt_share_tables): not freed here
};

}  // namespace gpw

struct gpw_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = true;
  int sm_count = 148;
  uint64_t launches = 0;
  std::map<std::string, gpw::Scratch> scratch;
  std::map<int, gpw::NttTables> ntt;  // by logn
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  float msm_acc_ms = 0.f, msm_total_ms = 0.f;
  uint64_t msm_digits = 0;
  // cumulative statistics of the G1 bucket-accumulation kernel (bench.py roofline): [0] G1, [1] G2
  double msm_acc_ms_sum[2] = {0, 0}, msm_total_ms_sum[2] = {0, 0};
  uint64_t msm_points_sum[2] = {0, 0}, msm_digits_sum[2] = {0, 0}, msm_calls[2] = {0, 0};
  // what a shared bucket sort (msm_dev_impl's sort_tag) was built for: reuse by an MSM of any other shape is refused
  struct SortDesc {
    size_t n = 0;
    int c = 0, win_lo = 0, win_hi = 0, fixed = 0, rounds = 0;
    const void* scalars = nullptr;
    bool operator==(const SortDesc& o) const {
      return n == o.n && c == o.c && win_lo == o.win_lo && win_hi == o.win_hi && fixed == o.fixed && rounds == o.rounds && scalars == o.scalars;
    }
  };
  std::map<std::string, SortDesc> sort_desc;
  // NCCL communicator of the sharded MSM (csrc/comm.cu); ncclComm_t kept opaque here
  void* nccl_comm = nullptr;
  int comm_size = 1, comm_rank = 0;
  int msm_affine_rounds = -1;  // -1: environment / default (off); see msm_dev_impl
  bool poseidon_consts_loaded = false;
  // Pinned host staging for the small host<->device transfers of the proving path (window sums, status words,
  // challenges). A cudaMemcpyAsync to or from PAGEABLE memory blocks inside the driver until the stream has drained -
  // behind a 190 ms solve spine that stalls every other host thread's launches - so these go through pinned memory
  // and wait with cudaStreamSynchronize instead.
  uint8_t* pin = nullptr;
  size_t pin_off = 0;
  static constexpr size_t PIN_CAP = 1 << 16;
  void* pin_take(size_t bytes) {
    bytes = (bytes + 15) & ~(size_t)15;
    if (pin_off + bytes > PIN_CAP) {  // wrap around: everything staged so far must have been consumed
      cudaStreamSynchronize(stream);
      pin_off = 0;
    }
    void* p = pin + pin_off;
    pin_off += bytes;
    return p;
  }

  int get_scratch(const char* name, size_t bytes, void** out);
};

namespace gpw {
inline int div_up(size_t a, size_t b) { return (int)((a + b - 1) / b); }
}  // namespace gpw
