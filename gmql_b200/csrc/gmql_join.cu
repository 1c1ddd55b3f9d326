// gmql_join.cu — genometric JOIN on the device (placeholder until the kernels land).
#include "common.cuh"
extern "C" int gmql_join(gmql_ctx* ctx, const gmql_regions*, const gmql_regions*, const gmql_pair*, int64_t, const gmql_atom*, int32_t, int32_t,
                         int64_t, int64_t, const int64_t*, const int64_t*, gmql_result** out) {
  if (out) *out = nullptr;
  if (ctx) ctx->err = "gmql_join: not built yet";
  return GMQL_ERR_UNSUPPORTED;
}
