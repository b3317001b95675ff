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
  // ---- deferred MSMs (msm_impl.cuh, used by the wrap prover) ------------------------------------------------------------
  // An MSM ends in short latency-bound grids (bucket fix-up, two-level window reduction, sums) and a host-side Horner over
  // <= 32 window sums. Run back to back, every MSM of a proof leaves the device nearly idle for that tail and then waits for
  // the host. In deferred mode (msm_defer_begin .. msm_finish_all) an MSM call only ENQUEUES: sort + bucket accumulation on
  // `stream`, the tail on `stream_hi` - a second stream of the context at the device's greatest priority, so that its small
  // CTAs take the next free slots while the NEXT MSM's accumulation fills the rest of the device - and the window sums land
  // in pinned memory; msm_finish() waits for the MSM's event and folds them on the host. Scratch that the tail still reads
  // while the next MSM starts is double-buffered (tag suffix a / b by call parity).
  cudaStream_t stream_hi = nullptr;
  cudaEvent_t ev_hop = nullptr;
  struct MsmPending {
    int group = 0;  // 1 = G1, 2 = G2
    const void* hw = nullptr;      // pinned: 2 nw window sums (XYZZ)
    const uint32_t* Mp = nullptr;  // pinned: number of sorted entries
    int nw = 0, c = 0, win_lo = 0;
    uint32_t chunk = 0;
    size_t n = 0;
    uint64_t* out = nullptr;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};  // start, sorted, accumulated, done
    float acc_ms = 0, total_ms = 0;
    uint64_t digits = 0;
    bool open = false;
  };
  static constexpr int MAX_PENDING = 16;
  MsmPending pend[MAX_PENDING];
  int n_pend = 0;          // MSMs enqueued since msm_defer_begin
  bool msm_defer = false;  // msm_dev_impl leaves the result to msm_finish()
  int msm_parity = 0;      // scratch double buffer of the next deferred MSM
  cudaEvent_t slot_done[2] = {nullptr, nullptr};  // tail of the last MSM that used scratch set a / b
  bool slot_used[2] = {false, false};
  bool msm_overlap = true;  // this call: deferred MSMs allowed (set by the wrap entry points from msm_overlap_mode)
  int msm_overlap_mode = -1;  // option "msm_overlap": -1 = automatic (a lone proof overlaps, a stream of proofs does not), 0 / 1 = forced
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
  static constexpr size_t PIN_CAP = 1 << 18;
  void* pin_take(size_t bytes) {
    bytes = (bytes + 15) & ~(size_t)15;
    if (pin_off + bytes > PIN_CAP) {  // wrap around: everything staged so far must have been consumed
      cudaStreamSynchronize(stream);
      if (stream_hi) cudaStreamSynchronize(stream_hi);
      pin_off = 0;
    }
    void* p = pin + pin_off;
    pin_off += bytes;
    return p;
  }

  // makes sure the next `bytes` of staging come without a wrap-around (results of deferred MSMs stay put until finished)
  void pin_reserve(size_t bytes) {
    if (pin_off + bytes > PIN_CAP) {
      cudaStreamSynchronize(stream);
      if (stream_hi) cudaStreamSynchronize(stream_hi);
      pin_off = 0;
    }
  }

  int get_scratch(const char* name, size_t bytes, void** out);
};

namespace gpw {
inline int div_up(size_t a, size_t b) { return (int)((a + b - 1) / b); }
}  // namespace gpw
