// gmql_cover.cu — COVER / FLAT / SUMMIT / HISTOGRAM on the device (placeholder until the kernels land).
#include "common.cuh"
extern "C" int gmql_cover(gmql_ctx* ctx, const gmql_regions*, const int32_t*, int32_t, int32_t, const int32_t*, const int32_t*, int64_t,
                          const gmql_agg*, int32_t, gmql_result** out) {
  if (out) *out = nullptr;
  if (ctx) ctx->err = "gmql_cover: not built yet";
  return GMQL_ERR_UNSUPPORTED;
}
extern "C" int gmql_cover_count_samples(gmql_ctx* ctx, const gmql_regions*, int32_t*) {
  if (ctx) ctx->err = "gmql_cover_count_samples: not built yet";
  return GMQL_ERR_UNSUPPORTED;
}
extern "C" int gmql_result_as_regions(gmql_ctx* ctx, gmql_result*, gmql_regions*) {
  if (ctx) ctx->err = "gmql_result_as_regions: not built yet";
  return GMQL_ERR_UNSUPPORTED;
}
