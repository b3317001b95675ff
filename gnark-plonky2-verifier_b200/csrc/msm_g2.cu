#include "msm_impl.cuh"
using namespace gpw;
extern "C" int gpw_msm_g2(gpw_ctx* ctx, const uint64_t* scalars, const uint64_t* points, size_t n, int scalars_mont,
                          int window_bits, uint64_t* out_affine) {
  return msm_host_impl<Fp2>(ctx, scalars, points, n, scalars_mont, window_bits, out_affine, "msm2");
}

extern "C" int gpw_msm_g2_dev(gpw_ctx* ctx, uint64_t scalars_dev, uint64_t points_dev, size_t n, int scalars_mont,
                              int window_bits, int win_lo, int win_hi, uint64_t* out_affine) {
  if (!ctx || !out_affine) {
    set_error("msm: null argument");
    return GPW_EINVAL;
  }
  return msm_dev_impl<Fp2>(ctx, (const Fr*)scalars_dev, (const Affine<Fp2>*)points_dev, n, scalars_mont, window_bits,
                           win_lo, win_hi, out_affine, "msm2");
}


// Internal (wrap.cu): G2 MSM sharing the bucket sort of a G1 MSM over the same scalars (see gpw_msm_g1_shared_dev).
extern "C" int gpw_msm_g2_shared_dev(gpw_ctx* ctx, uint64_t scalars_dev, uint64_t points_dev, size_t n, int scalars_mont, int window_bits,
                                     const char* sort_tag, int reuse, uint64_t* out_affine) {
  if (!ctx || !out_affine || !sort_tag) {
    set_error("msm: null argument");
    return GPW_EINVAL;
  }
  return msm_dev_impl<Fp2>(ctx, (const Fr*)scalars_dev, (const Affine<Fp2>*)points_dev, n, scalars_mont, window_bits, 0, 0, out_affine,
                           "msm2", 0, sort_tag, reuse != 0);
}

extern "C" int gpwi_msm_finish_g2(gpw_ctx* ctx, int idx) {
  if (!ctx || idx < 0 || idx >= gpw_ctx::MAX_PENDING) return GPW_EINVAL;
  return msm_finish_impl<Fp2>(ctx, ctx->pend[idx]);
}
