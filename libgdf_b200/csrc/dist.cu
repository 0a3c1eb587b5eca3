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

// ---- row validity as a travelling byte column (composite-key joins with NULLs across ranks, C5) ----
// gdfx_rows_valid_to_bytes: out[i] = 1 if every given column's validity bit i is set (columns without
// a mask count as all valid) - the row-valid rule of the reference's gdf_table (gdf_table.cuh:63-98).
// gdfx_bytes_to_valid: the inverse on the receiving rank: Arrow LSB-first bitmask from a byte column.
namespace b200 {
namespace {
struct MaskSet {
  const gdf_valid_type* m[16];
  int n;
};
__global__ void __launch_bounds__(256) rows_valid_kernel(MaskSet ms, size_t rows, int8_t* __restrict__ out) {
  const size_t nbytes = (rows + 7) / 8;
  const size_t stride = (size_t)gridDim.x * 256;
  for (size_t b = (size_t)blockIdx.x * 256 + threadIdx.x; b < nbytes; b += stride) {
    unsigned bits = 0xffu;
    for (int c = 0; c < ms.n; ++c) bits &= ms.m[c][b];
    const size_t base = b * 8;
    const int cnt = rows - base < 8 ? (int)(rows - base) : 8;
    for (int j = 0; j < cnt; ++j) out[base + j] = (int8_t)((bits >> j) & 1u);
  }
}
__global__ void __launch_bounds__(256) bytes_to_valid_kernel(const int8_t* __restrict__ in, size_t rows,
                                                             gdf_valid_type* __restrict__ out) {
  const size_t nbytes = (rows + 7) / 8;
  const size_t stride = (size_t)gridDim.x * 256;
  for (size_t b = (size_t)blockIdx.x * 256 + threadIdx.x; b < nbytes; b += stride) {
    unsigned bits = 0;
    const size_t base = b * 8;
    const int cnt = rows - base < 8 ? (int)(rows - base) : 8;
    for (int j = 0; j < cnt; ++j) bits |= (unsigned)(in[base + j] != 0) << j;
    out[b] = (gdf_valid_type)bits;
  }
}
int blocks_for(size_t items) {
  size_t want = (items + 255) / 256;
  const size_t cap = (size_t)sm_count() * 8;
  return (int)(want < 1 ? 1 : (want < cap ? want : cap));
}
}  // namespace
}  // namespace b200

extern "C" gdf_error gdfx_rows_valid_to_bytes(gdf_column** cols, int num_cols, int8_t* out) {
  B200_REQUIRE(cols != nullptr && out != nullptr && num_cols >= 1, GDF_DATASET_EMPTY);
  B200_REQUIRE(num_cols <= 16, GDF_JOIN_TOO_MANY_COLUMNS);
  const size_t rows = cols[0]->size;
  if (rows == 0) return GDF_SUCCESS;
  MaskSet ms;
  ms.n = 0;
  for (int c = 0; c < num_cols; ++c) {
    B200_REQUIRE(cols[c]->size == rows, GDF_COLUMN_SIZE_MISMATCH);
    if (cols[c]->valid) ms.m[ms.n++] = cols[c]->valid;
  }
  rows_valid_kernel<<<blocks_for((rows + 7) / 8), 256>>>(ms, rows, out);
  B200_CHECK_LAST();
  return GDF_SUCCESS;
}

extern "C" gdf_error gdfx_bytes_to_valid(const int8_t* in, size_t rows, gdf_valid_type* out) {
  if (rows == 0) return GDF_SUCCESS;
  B200_REQUIRE(in != nullptr && out != nullptr, GDF_DATASET_EMPTY);
  bytes_to_valid_kernel<<<blocks_for((rows + 7) / 8), 256>>>(in, rows, out);
  B200_CHECK_LAST();
  return GDF_SUCCESS;
}
