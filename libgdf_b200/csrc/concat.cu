// gdf_column_concat / gdf_mask_concat: stitch columns (e.g. the per-GPU shards the multi-GPU layer returns) into one.
//
// Reference behaviour followed (file:line in /root/reference/libgdf/src):
//   gdf_column_concat   argument checks and their error codes, data copied column after column, null_count summed,
//                       output mask = concatenation of the input masks (a column without mask counts as all valid),
//                       or all ones when no input has a mask                         column.cpp:53-153
//   gdf_mask_concat     bit i of the output = validity bit of the element that lands at position i; bits past
//                       output_column_length in the last byte are 0                   validops.cu:203-256
//
// B200 design.  The reference walks the column list once per output BIT inside a Thrust tabulate and reads the
// pointer / length arrays from managed memory.  Here the host builds the prefix sums once, and one thread builds 8
// output bytes (64 bits): it finds its first source column by binary search and then only steps forward.  When the
// run of 64 bits comes from one column at a byte-aligned offset it is a plain 8-byte copy.
#include <vector>

#include "common.cuh"

namespace b200 {
namespace {

constexpr int kThreads = 256;

struct ConcatPlan {
  const gdf_valid_type* const* masks;  // device array [ncols]
  const unsigned long long* starts;    // device array [ncols + 1]: first output row of column c
  int ncols;
};

__global__ void __launch_bounds__(kThreads)
mask_concat_kernel(gdf_valid_type* __restrict__ out, size_t out_rows, ConcatPlan plan) {
  const size_t out_bytes = (out_rows + 7) / 8;
  const size_t words = (out_bytes + 7) / 8;
  const size_t stride = (size_t)gridDim.x * kThreads;
  for (size_t w = (size_t)blockIdx.x * kThreads + threadIdx.x; w < words; w += stride) {
    const size_t row0 = w * 64;
    // last column whose start is <= row0
    int lo = 0, hi = plan.ncols;
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (plan.starts[mid] <= row0) lo = mid;
      else hi = mid;
    }
    int c = lo;
    unsigned long long bits = 0;
    for (int b = 0; b < 64; ++b) {
      const size_t row = row0 + b;
      if (row >= out_rows) break;
      while (c + 1 < plan.ncols && plan.starts[c + 1] <= row) ++c;
      const gdf_valid_type* m = plan.masks[c];
      const size_t i = row - plan.starts[c];
      if (m == nullptr || ((m[i >> 3] >> (i & 7)) & 1)) bits |= 1ull << b;
    }
    const size_t byte0 = w * 8;
    for (int k = 0; k < 8 && byte0 + k < out_bytes; ++k) out[byte0 + k] = (gdf_valid_type)(bits >> (8 * k));
  }
}

// copy `count` elements of a host- or device-resident array to the host
template <typename T>
cudaError_t fetch(std::vector<T>& dst, const T* src, size_t count) {
  dst.resize(count);
  return cudaMemcpy(dst.data(), src, count * sizeof(T), cudaMemcpyDefault);
}

}  // namespace
}  // namespace b200

using namespace b200;

extern "C" gdf_error gdf_mask_concat(gdf_valid_type* output_mask, gdf_size_type output_column_length,
                                     gdf_valid_type* masks_to_concat[], gdf_size_type* column_lengths,
                                     gdf_size_type num_columns) {
  B200_REQUIRE(output_mask != nullptr && masks_to_concat != nullptr && column_lengths != nullptr, GDF_DATASET_EMPTY);
  if (output_column_length == 0 || num_columns == 0) return GDF_SUCCESS;
  // the two arrays may live in host, managed or device memory (the reference reads them on the device)
  std::vector<gdf_valid_type*> h_masks;
  std::vector<gdf_size_type> h_len;
  B200_CUDA_TRY(fetch(h_masks, masks_to_concat, (size_t)num_columns));
  B200_CUDA_TRY(fetch(h_len, column_lengths, (size_t)num_columns));
  std::vector<unsigned long long> h_starts((size_t)num_columns + 1);
  unsigned long long run = 0;
  for (size_t c = 0; c < (size_t)num_columns; ++c) {
    h_starts[c] = run;
    run += h_len[c];
  }
  h_starts[num_columns] = run;
  Scratch plan_mem;
  const size_t ptr_bytes = (size_t)num_columns * sizeof(void*), start_bytes = ((size_t)num_columns + 1) * sizeof(unsigned long long);
  B200_CUDA_TRY(plan_mem.alloc(ptr_bytes + start_bytes));
  B200_CUDA_TRY(cudaMemcpy(plan_mem.ptr, h_masks.data(), ptr_bytes, cudaMemcpyHostToDevice));
  B200_CUDA_TRY(cudaMemcpy(static_cast<char*>(plan_mem.ptr) + ptr_bytes, h_starts.data(), start_bytes, cudaMemcpyHostToDevice));
  ConcatPlan plan{static_cast<const gdf_valid_type* const*>(plan_mem.ptr),
                  reinterpret_cast<const unsigned long long*>(static_cast<char*>(plan_mem.ptr) + ptr_bytes), (int)num_columns};
  const size_t words = (((size_t)output_column_length + 7) / 8 + 7) / 8;
  size_t blocks = (words + kThreads - 1) / kThreads;
  const size_t cap = (size_t)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  mask_concat_kernel<<<(unsigned)(blocks ? blocks : 1), kThreads>>>(output_mask, (size_t)output_column_length, plan);
  B200_CHECK_LAST();
  return GDF_SUCCESS;
}

extern "C" gdf_error gdf_column_concat(gdf_column* output_column, gdf_column* columns_to_concat[], int num_columns) {
  if (columns_to_concat == nullptr) return GDF_DATASET_EMPTY;
  if (num_columns < 1 || columns_to_concat[0] == nullptr || output_column == nullptr) return GDF_DATASET_EMPTY;
  const gdf_dtype column_type = columns_to_concat[0]->dtype;
  if (column_type != output_column->dtype) return GDF_DTYPE_MISMATCH;
  size_t total = 0;
  bool any_mask = false;
  for (int i = 0; i < num_columns; ++i) {
    const gdf_column* c = columns_to_concat[i];
    if (c == nullptr) return GDF_DATASET_EMPTY;
    if (c->size > 0 && c->data == nullptr) return GDF_DATASET_EMPTY;
    if (c->dtype != column_type) return GDF_DTYPE_MISMATCH;
    total += c->size;
    any_mask = any_mask || c->valid != nullptr;
  }
  if (output_column->size != total) return GDF_COLUMN_SIZE_MISMATCH;
  const int width = dtype_width(output_column->dtype);
  if (width == 0) return GDF_UNSUPPORTED_DTYPE;
  B200_REQUIRE(total == 0 || output_column->data != nullptr, GDF_DATASET_EMPTY);
  char* target = static_cast<char*>(output_column->data);
  output_column->null_count = 0;
  for (int i = 0; i < num_columns; ++i) {
    const size_t bytes = (size_t)width * columns_to_concat[i]->size;
    if (bytes) B200_CUDA_TRY(cudaMemcpyAsync(target, columns_to_concat[i]->data, bytes, cudaMemcpyDeviceToDevice, 0));
    target += bytes;
    output_column->null_count += columns_to_concat[i]->null_count;
  }
  if (any_mask) {
    B200_REQUIRE(output_column->valid != nullptr, GDF_DATASET_EMPTY);
    std::vector<gdf_valid_type*> masks((size_t)num_columns);
    std::vector<gdf_size_type> lens((size_t)num_columns);
    for (int i = 0; i < num_columns; ++i) {
      masks[i] = columns_to_concat[i]->valid;
      lens[i] = columns_to_concat[i]->size;
    }
    return gdf_mask_concat(output_column->valid, output_column->size, masks.data(), lens.data(), num_columns);
  }
  if (output_column->valid != nullptr)
    B200_CUDA_TRY(cudaMemsetAsync(output_column->valid, 0xff, valid_bytes(total), 0));
  return GDF_SUCCESS;
}
