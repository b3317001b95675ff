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

extern "C" int gpw_msm_last_stats(gpw_ctx* ctx, float* accumulate_ms, float* total_ms, uint64_t* nonzero_digits) {
  if (!ctx) return GPW_EINVAL;
  if (accumulate_ms) *accumulate_ms = ctx->msm_acc_ms;
  if (total_ms) *total_ms = ctx->msm_total_ms;
  if (nonzero_digits) *nonzero_digits = ctx->msm_digits;
  return GPW_OK;
}
