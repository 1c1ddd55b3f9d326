// gmql_profile.cu — dataset statistics of the reference's Profiler on the device.
//
// Replaces the two foldByKey passes of Profiler.profile (GMQL-Profiler/.../Profilers/Profiler.scala:124-153; called on every
// materialised result, GMQLSparkExecutor.scala:158-173): per sample min start, max stop, min / max / sum of lengths, sum of
// squared lengths and the region count (ProfilerValue :65, reduceFunc :71-82).  All integer; sums wrap like Java longs.
// One read of (start, stop, sample); a warp whose 32 regions belong to one sample (the usual case: samples arrive one
// after the other) reduces in registers and issues one set of atomics.
#include "common.cuh"
#include <algorithm>

namespace {

struct ProfAcc {
  long long* left; long long* right; long long* minlen; long long* maxlen;
  unsigned long long* sumlen; unsigned long long* sumsq; unsigned long long* count;
};

__global__ void __launch_bounds__(256) k_profile(const void* __restrict__ start, const void* __restrict__ stop, int bits, const int32_t* __restrict__ sample,
                                                 int64_t n, int n_samples, ProfAcc acc, int* __restrict__ bad) {
  const int lane = threadIdx.x & 31;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  // whole warps iterate together (the loop bound is rounded up to a warp), so the shuffles below are convergent
  for (int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x - lane; i0 < n; i0 += stride) {
    const int64_t i = i0 + lane;
    const bool ok = i < n;
    int sm = ok ? (sample ? sample[i] : 0) : -1;
    long long s = ok ? (long long)ld_coord(start, bits, i) : 0, e = ok ? (long long)ld_coord(stop, bits, i) : 0;
    if (ok && (sm < 0 || sm >= n_samples)) { *bad = 1; sm = -1; }
    const bool live = sm >= 0;
    const long long len = e - s;
    const int first = __shfl_sync(0xffffffffu, sm, 0);
    if (__all_sync(0xffffffffu, sm == first)) {
      if (first < 0) continue;
      long long l = s, r = e, mn = len, mx = len;
      unsigned long long sl = (unsigned long long)len, sq = (unsigned long long)len * (unsigned long long)len, c = 1;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        long long l2 = __shfl_xor_sync(0xffffffffu, l, o), r2 = __shfl_xor_sync(0xffffffffu, r, o);
        long long mn2 = __shfl_xor_sync(0xffffffffu, mn, o), mx2 = __shfl_xor_sync(0xffffffffu, mx, o);
        l = l2 < l ? l2 : l; r = r2 > r ? r2 : r; mn = mn2 < mn ? mn2 : mn; mx = mx2 > mx ? mx2 : mx;
        sl += __shfl_xor_sync(0xffffffffu, sl, o); sq += __shfl_xor_sync(0xffffffffu, sq, o); c += __shfl_xor_sync(0xffffffffu, c, o);
      }
      if (lane == 0) {
        atomicMin(&acc.left[first], l); atomicMax(&acc.right[first], r);
        atomicMin(&acc.minlen[first], mn); atomicMax(&acc.maxlen[first], mx);
        atomicAdd(&acc.sumlen[first], sl); atomicAdd(&acc.sumsq[first], sq); atomicAdd(&acc.count[first], c);
      }
    } else if (live) {
      atomicMin(&acc.left[sm], s); atomicMax(&acc.right[sm], e);
      atomicMin(&acc.minlen[sm], len); atomicMax(&acc.maxlen[sm], len);
      atomicAdd(&acc.sumlen[sm], (unsigned long long)len); atomicAdd(&acc.sumsq[sm], (unsigned long long)len * (unsigned long long)len);
      atomicAdd(&acc.count[sm], 1ull);
    }
  }
}

__global__ void k_profile_init(ProfAcc acc, int n_samples) {
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n_samples; k += gridDim.x * blockDim.x) {
    acc.left[k] = 0x7fffffffffffffffLL; acc.right[k] = (long long)0x8000000000000000ULL;      // zeroProfilerValue (:69)
    acc.minlen[k] = 0x7fffffffffffffffLL; acc.maxlen[k] = (long long)0x8000000000000000ULL;
    acc.sumlen[k] = 0; acc.sumsq[k] = 0; acc.count[k] = 0;
  }
}

}  // namespace

extern "C" int gmql_profile(gmql_ctx* ctx, const gmql_regions* in, gmql_result** out) {
  if (!ctx || !out) return GMQL_ERR_INVALID;
  *out = nullptr;
  gmql_result* res = nullptr;
  try {
    CUDA_TRY(cudaSetDevice(ctx->device));
    ctx_begin_call(ctx);
    GMQL_REQUIRE(in != nullptr, GMQL_ERR_INVALID, "null dataset");
    cudaEventRecord(ctx->ev_h0, ctx->stream);
    DevRegions d;
    gmql_upload_regions(ctx, in, std::vector<int>(), &d);
    cudaEventRecord(ctx->ev_h1, ctx->stream);
    cudaEventRecord(ctx->ev_k0, ctx->stream);
    const int ns = d.n_samples;
    DevBuf<long long> left(ctx, ns), right(ctx, ns), minlen(ctx, ns), maxlen(ctx, ns);
    DevBuf<unsigned long long> sumlen(ctx, ns), sumsq(ctx, ns), count(ctx, ns);
    DevBuf<int> bad(ctx, 1);
    bad.zero();
    ProfAcc acc;
    acc.left = left.p; acc.right = right.p; acc.minlen = minlen.p; acc.maxlen = maxlen.p; acc.sumlen = sumlen.p; acc.sumsq = sumsq.p; acc.count = count.p;
    LAUNCH(ctx, k_profile_init, std::max(1, std::min(ns / 256 + 1, ctx->num_sms)), 256, 0, acc, ns);
    if (d.n) {
      int g = (int)std::min<int64_t>(ceil_div64(d.n, 256 * 8), (int64_t)ctx->num_sms * 16);
      LAUNCH(ctx, k_profile, g, 256, 0, d.start, d.stop, d.coord_bits, d.sample, d.n, ns, acc, bad.p);
    }
    cudaEventRecord(ctx->ev_k1, ctx->stream);
    int h_bad = 0;
    CUDA_TRY(cudaMemcpyAsync(&h_bad, bad.p, sizeof h_bad, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(GMQL_SYNC(ctx));
    GMQL_REQUIRE(!h_bad, GMQL_ERR_INVALID, "sample index out of range");
    res = new gmql_result();
    res->owner = ctx;
    res->rows = ns;
    res->kind = 5;
    result_add(res, GMQL_COL_PROF_LEFTMOST, left);
    result_add(res, GMQL_COL_PROF_RIGHTMOST, right);
    result_add(res, GMQL_COL_PROF_MIN_LENGTH, minlen);
    result_add(res, GMQL_COL_PROF_MAX_LENGTH, maxlen);
    result_add(res, GMQL_COL_PROF_SUM_LENGTH, sumlen);
    result_add(res, GMQL_COL_PROF_SUM_LENGTH_SQ, sumsq);
    result_add(res, GMQL_COL_PROF_COUNT, count);
    *out = res;
    return GMQL_OK;
  } catch (const gmql_error& ex) {
    ctx->err = ex.what();
    if (res) gmql_result_free(ctx, res);
    cudaStreamSynchronize(ctx->stream);
    return ex.code;
  } catch (const std::exception& ex) {
    ctx->err = ex.what();
    if (res) gmql_result_free(ctx, res);
    return GMQL_ERR_INVALID;
  }
}
