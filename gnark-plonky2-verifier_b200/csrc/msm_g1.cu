#include "msm_impl.cuh"
using namespace gpw;
extern "C" int gpw_msm_g1(gpw_ctx* ctx, const uint64_t* scalars, const uint64_t* points, size_t n, int scalars_mont,
                          int window_bits, uint64_t* out_affine) {
  return msm_host_impl<Fp>(ctx, scalars, points, n, scalars_mont, window_bits, out_affine, "msm1");
}

extern "C" int gpw_msm_g1_dev(gpw_ctx* ctx, uint64_t scalars_dev, uint64_t points_dev, size_t n, int scalars_mont,
                              int window_bits, int win_lo, int win_hi, uint64_t* out_affine) {
  if (!ctx || !out_affine) {
    set_error("msm: null argument");
    return GPW_EINVAL;
  }
  return msm_dev_impl<Fp>(ctx, (const Fr*)scalars_dev, (const Affine<Fp>*)points_dev, n, scalars_mont, window_bits,
                          win_lo, win_hi, out_affine, "msm1");
}

// Fixed-base MSM (bases known at setup, e.g. a proving key): gpw_msm_g1_fixed_table precomputes 2^(c w) P_i for the
// W = ceil(255 / c) windows (table_dev: W * n affine points), gpw_msm_g1_fixed_dev then needs ceil(255 / c) bucket
// additions per full-width scalar into a single bucket set.
extern "C" int gpw_msm_g1_fixed_table(gpw_ctx* ctx, uint64_t points_dev, size_t n, int window_bits, int n_windows, uint64_t table_dev) {
  return msm_fixed_table_impl<Fp>(ctx, (const Affine<Fp>*)points_dev, n, window_bits, n_windows, (Affine<Fp>*)table_dev);
}

extern "C" int gpw_msm_g1_fixed_dev(gpw_ctx* ctx, uint64_t scalars_dev, uint64_t table_dev, size_t n, int scalars_mont, int window_bits,
                                    int n_windows, uint64_t* out_affine) {
  if (!ctx || !out_affine || window_bits < 2 || n_windows < 1) {
    set_error("msm_fixed: bad argument");
    return GPW_EINVAL;
  }
  return msm_dev_impl<Fp>(ctx, (const Fr*)scalars_dev, (const Affine<Fp>*)table_dev, n, scalars_mont, window_bits, 0, 0, out_affine,
                          "msm1f", n_windows);
}

// Internal (wrap.cu): G1 MSM that shares its digit decomposition + bucket sort with other MSMs over the same scalars.
// fixed_windows != 0: points_dev is a fixed-base table. reuse != 0: the sort named sort_tag is already in place.
extern "C" int gpw_msm_g1_shared_dev(gpw_ctx* ctx, uint64_t scalars_dev, uint64_t points_dev, size_t n, int scalars_mont, int window_bits,
                                     int fixed_windows, const char* sort_tag, int reuse, uint64_t* out_affine) {
  if (!ctx || !out_affine || !sort_tag) {
    set_error("msm: null argument");
    return GPW_EINVAL;
  }
  return msm_dev_impl<Fp>(ctx, (const Fr*)scalars_dev, (const Affine<Fp>*)points_dev, n, scalars_mont, window_bits, 0, 0, out_affine,
                          fixed_windows ? "msm1f" : "msm1", fixed_windows, sort_tag, reuse != 0);
}

extern "C" int gpw_msm_last_stats(gpw_ctx* ctx, float* accumulate_ms, float* total_ms, uint64_t* nonzero_digits) {
  if (!ctx) return GPW_EINVAL;
  if (accumulate_ms) *accumulate_ms = ctx->msm_acc_ms;
  if (total_ms) *total_ms = ctx->msm_total_ms;
  if (nonzero_digits) *nonzero_digits = ctx->msm_digits;
  return GPW_OK;
}

// Cumulative per-group MSM statistics since the last reset: out8 = {acc_ms_sum, total_ms_sum, points, digits, calls}
// for group (1 = G1, 2 = G2). reset != 0 clears both groups afterwards.
extern "C" int gpw_msm_cumulative_stats(gpw_ctx* ctx, int group, int reset, double* out5) {
  if (!ctx || group < 1 || group > 2) return GPW_EINVAL;
  const int g = group - 1;
  if (out5) {
    out5[0] = ctx->msm_acc_ms_sum[g];
    out5[1] = ctx->msm_total_ms_sum[g];
    out5[2] = (double)ctx->msm_points_sum[g];
    out5[3] = (double)ctx->msm_digits_sum[g];
    out5[4] = (double)ctx->msm_calls[g];
  }
  if (reset)
    for (int i = 0; i < 2; i++) {
      ctx->msm_acc_ms_sum[i] = ctx->msm_total_ms_sum[i] = 0;
      ctx->msm_points_sum[i] = ctx->msm_digits_sum[i] = ctx->msm_calls[i] = 0;
    }
  return GPW_OK;
}

// ---- deferred MSMs (common.cuh): internal to libgpw, used by the wrap prover ---------------------------------------------
extern "C" int gpwi_msm_finish_g2(gpw_ctx* ctx, int idx);

// msm_defer_begin: from here on gpw_msm_*_dev calls on this context only enqueue their work and return at once
extern "C" int gpwi_msm_defer_begin(gpw_ctx* ctx) {
  if (!ctx) return GPW_EINVAL;
  if (!ctx->msm_overlap) return GPW_OK;  // synchronous MSMs: every call finishes itself
  GPW_CUDA(cudaSetDevice(ctx->device));
  ctx->pin_reserve(96 * 1024);  // window sums of a whole proof's MSMs stay in the staging area until finished
  ctx->msm_defer = true;
  ctx->n_pend = 0;
  ctx->msm_parity = 0;  // every proof walks the two scratch sets in the same order: their sizes settle after the first proof
  return GPW_OK;
}

// result of the idx-th MSM since msm_defer_begin (waits for it; idempotent)
extern "C" int gpwi_msm_finish(gpw_ctx* ctx, int idx) {
  if (!ctx || idx < 0 || idx >= gpw_ctx::MAX_PENDING) return GPW_EINVAL;
  if (!ctx->msm_defer || idx >= ctx->n_pend) return GPW_OK;
  GPW_CUDA(cudaSetDevice(ctx->device));
  gpw_ctx::MsmPending& P = ctx->pend[idx];
  return P.group == 2 ? gpwi_msm_finish_g2(ctx, idx) : msm_finish_impl<Fp>(ctx, P);
}

// finishes every MSM still open and leaves deferred mode; `stream` is ordered after all of their work
extern "C" int gpwi_msm_defer_end(gpw_ctx* ctx) {
  if (!ctx) return GPW_EINVAL;
  if (!ctx->msm_defer) return GPW_OK;
  int rc = GPW_OK;
  for (int i = 0; i < ctx->n_pend; i++) {
    const int r = gpwi_msm_finish(ctx, i);
    if (rc == GPW_OK) rc = r;
  }
  ctx->msm_defer = false;
  ctx->n_pend = 0;
  for (int p = 0; p < 2; p++)
    if (ctx->slot_used[p]) {
      cudaStreamWaitEvent(ctx->stream, ctx->slot_done[p], 0);
      ctx->slot_used[p] = false;
    }
  return rc;
}
