// Device-side "row view over N columns" shared by gdf_hash, gdf_hash_partition, group-by and join.
//
// The reference keeps a gdf_table object in unified memory, passes it to kernels by reference and
// precomputes a row-validity byte array with a Thrust launch in its constructor
// (ref src/gdf_table.cuh:249-322).  Here the view is a small POD passed BY VALUE in kernel
// parameter space (constant bank, broadcast to all threads, no UVM page faults, no extra launch),
// and row validity is the AND of the column bits evaluated inside the consuming scan.
//
// Semantics reproduced:
//   row_valid   AND over columns of the LSB-first validity bit; a column without mask is all valid
//               (ref gdf_table.cuh:63-98)
//   row_hash    per-column hash of the value's bytes (MurmurHash3_x86_32 seed 0, or the identity
//               cast), first column taken as is, later ones folded with hash_combine
//               (ref gdf_table.cuh:705-854, hashmap/hash_functions.cuh:31-161)
//   rows_equal  typed `==` on every column (floats: NaN != NaN, -0.0 == 0.0)
//               (ref gdf_table.cuh:581-691); validity is checked separately by the callers
#pragma once
#include "common.cuh"

namespace b200 {

constexpr int kMaxCols = 16;

struct TableView {
  const void* data[kMaxCols];
  const gdf_valid_type* valid[kMaxCols];
  unsigned char dtype[kMaxCols];
  int ncols;
  size_t rows;
  bool any_valid;  // true if at least one column carries a mask
};

// Host: build a view from gdf_column pointers.  Returns false if ncols > kMaxCols.
inline bool make_view(TableView& tv, gdf_column* const* cols, int ncols) {
  if (ncols > kMaxCols || ncols < 0) return false;
  tv.ncols = ncols;
  tv.rows = ncols ? cols[0]->size : 0;
  tv.any_valid = false;
  for (int c = 0; c < kMaxCols; ++c) {
    tv.data[c] = nullptr;
    tv.valid[c] = nullptr;
    tv.dtype[c] = 0;
  }
  for (int c = 0; c < ncols; ++c) {
    tv.data[c] = cols[c]->data;
    tv.valid[c] = cols[c]->valid;
    tv.dtype[c] = (unsigned char)cols[c]->dtype;
    if (cols[c]->valid) tv.any_valid = true;
  }
  return true;
}

// dtypes the row hash / row compare understand (ref gdf_table.cuh:720-850)
inline bool hashable_dtype(int t) { return dtype_width(t) != 0; }

#ifdef __CUDACC__

static __device__ __forceinline__ bool row_valid(const TableView& tv, size_t row) {
  if (!tv.any_valid) return true;
  bool ok = true;
#pragma unroll 1
  for (int c = 0; c < tv.ncols; ++c) ok = ok && bit_valid(tv.valid[c], row);
  return ok;
}

// Raw value bits of (column c, row), zero-extended to 64 bits.
static __device__ __forceinline__ uint64_t load_bits(const TableView& tv, int c, size_t row) {
  switch (dtype_width(tv.dtype[c])) {
    case 1: return static_cast<const uint8_t*>(tv.data[c])[row];
    case 2: return static_cast<const uint16_t*>(tv.data[c])[row];
    case 4: return static_cast<const uint32_t*>(tv.data[c])[row];
    default: return static_cast<const uint64_t*>(tv.data[c])[row];
  }
}

static __device__ __forceinline__ uint32_t murmur_bits(int width, uint64_t bits) {
  switch (width) {
    case 1: return murmur3_32<1>(bits);
    case 2: return murmur3_32<2>(bits);
    case 4: return murmur3_32<4>(bits);
    default: return murmur3_32<8>(bits);
  }
}

// IdentityHash: static_cast<uint32_t>(typed value) (ref hash_functions.cuh:156-160).  Integers are
// sign-extended then truncated; floating values use the device's saturating float->uint32 convert.
static __device__ __forceinline__ uint32_t identity_bits(int dtype, uint64_t bits) {
  switch (dtype) {
    case GDF_INT8: return (uint32_t)(int32_t)(int8_t)bits;
    case GDF_INT16: return (uint32_t)(int32_t)(int16_t)bits;
    case GDF_FLOAT32: return (uint32_t)__uint_as_float((uint32_t)bits);
    case GDF_FLOAT64: return (uint32_t)__longlong_as_double((long long)bits);
    default: return (uint32_t)bits;
  }
}

template <bool IDENTITY>
static __device__ __forceinline__ uint32_t row_hash(const TableView& tv, size_t row) {
  uint32_t h = 0;
#pragma unroll 1
  for (int c = 0; c < tv.ncols; ++c) {
    const uint64_t bits = load_bits(tv, c, row);
    const uint32_t hc = IDENTITY ? identity_bits(tv.dtype[c], bits) : murmur_bits(dtype_width(tv.dtype[c]), bits);
    h = c ? hash_combine(h, hc) : hc;
  }
  return h;
}

static __device__ __forceinline__ bool value_equal(int dtype, uint64_t a, uint64_t b) {
  if (dtype == GDF_FLOAT32) return __uint_as_float((uint32_t)a) == __uint_as_float((uint32_t)b);
  if (dtype == GDF_FLOAT64) return __longlong_as_double((long long)a) == __longlong_as_double((long long)b);
  return a == b;
}

static __device__ __forceinline__ bool rows_equal(const TableView& a, size_t ra, const TableView& b, size_t rb) {
#pragma unroll 1
  for (int c = 0; c < a.ncols; ++c)
    if (!value_equal(a.dtype[c], load_bits(a, c, ra), load_bits(b, c, rb))) return false;
  return true;
}

#endif  // __CUDACC__
}  // namespace b200
