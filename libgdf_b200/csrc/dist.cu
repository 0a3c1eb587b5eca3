// Device helpers of the multi-GPU layer (libgdf_b200/dist.py).  The reference has no counterpart:
// it is a single-GPU library.
//
// gdfx_remap_indices: after the all-to-all, a rank joins the rows it RECEIVED; the join's int32
// outputs index those received rows.  Every received row carries the global row id it had in the
// caller's table (the "payload" column that travelled with the key), so the final result is
// payload[index] - one gather per output side, done in place on the library-owned index column.
#include "common.cuh"

namespace b200 {
namespace {

constexpr int kThreads = 256;

__global__ void __launch_bounds__(kThreads)
remap_kernel(int32_t* __restrict__ idx, size_t n, const int32_t* __restrict__ payload, size_t payload_rows) {
  const size_t stride = (size_t)gridDim.x * kThreads;
  constexpr int U = 4;  // independent gathers in flight per thread
  for (size_t i0 = (size_t)blockIdx.x * kThreads + threadIdx.x; i0 < n; i0 += stride * U) {
    int32_t v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const size_t i = i0 + (size_t)u * stride;
      v[u] = i < n ? idx[i] : -1;
    }
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (v[u] >= 0 && (size_t)v[u] < payload_rows) v[u] = __ldg(payload + v[u]);
      else v[u] = -1;  // JoinNoneValue stays -1 (ref join_kernels.cuh:17)
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const size_t i = i0 + (size_t)u * stride;
      if (i < n) idx[i] = v[u];
    }
  }
}

}  // namespace
}  // namespace b200

using namespace b200;

extern "C" gdf_error gdfx_remap_indices(gdf_column* indices, const int32_t* payload, size_t payload_rows) {
  B200_REQUIRE(indices != nullptr, GDF_DATASET_EMPTY);
  if (indices->size == 0) return GDF_SUCCESS;
  B200_REQUIRE(indices->dtype == GDF_INT32, GDF_UNSUPPORTED_DTYPE);
  B200_REQUIRE(indices->data != nullptr && payload != nullptr, GDF_DATASET_EMPTY);
  const size_t n = indices->size;
  size_t want = (n + kThreads * 4 - 1) / (kThreads * 4);
  const size_t cap = (size_t)sm_count() * 8;
  const int blocks = (int)(want < cap ? (want ? want : 1) : cap);
  B200_TIMED("dist_remap");
  remap_kernel<<<blocks, kThreads>>>(static_cast<int32_t*>(indices->data), n, payload, payload_rows);
  B200_CHECK_LAST();
  return GDF_SUCCESS;
}
