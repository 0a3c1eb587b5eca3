// Host-only plumbing of the gdf_* ABI: column/context views, error names, CUDA error accessors,
// NVTX ranges, plus the library-internal scratch allocator.  Reference behaviour followed:
//   libgdf/src/column.cpp:160-275, src/context.cpp:3-11, src/errorhandling.cpp:5-35,
//   src/cudautils.cu:4-14, src/nvtx_utils.cpp:19-71.
#include <mutex>

#include <nvtx3/nvToolsExt.h>

#include "common.cuh"

namespace b200 {

int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

static cudaError_t tune_default_pool() {
  static std::mutex mu;
  static bool tuned[64] = {false};
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  std::lock_guard<std::mutex> g(mu);
  if (dev >= 0 && dev < 64 && !tuned[dev]) {
    cudaMemPool_t pool;
    e = cudaDeviceGetDefaultMemPool(&pool, dev);
    if (e != cudaSuccess) return e;
    uint64_t never = UINT64_MAX;
    e = cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &never);
    if (e != cudaSuccess) return e;
    tuned[dev] = true;
  }
  return cudaSuccess;
}

cudaError_t scratch_alloc(void** p, size_t bytes, cudaStream_t s) {
  cudaError_t e = tune_default_pool();
  if (e != cudaSuccess) return e;
  return cudaMallocAsync(p, bytes, s);
}
cudaError_t scratch_free(void* p, cudaStream_t s) { return cudaFreeAsync(p, s); }

void* pinned_mailbox() {
  static thread_local void* box = nullptr;
  if (!box) {
    if (cudaHostAlloc(&box, 256, cudaHostAllocDefault) != cudaSuccess) box = nullptr;
  }
  return box;
}

}  // namespace b200

extern "C" {

gdf_size_type gdf_column_sizeof() { return sizeof(gdf_column); }

gdf_error gdf_column_view(gdf_column* column, void* data, gdf_valid_type* valid, gdf_size_type size,
                          gdf_dtype dtype) {
  column->data = data;
  column->valid = valid;
  column->size = size;
  column->dtype = dtype;
  column->null_count = 0;
  return GDF_SUCCESS;
}

gdf_error gdf_column_view_augmented(gdf_column* column, void* data, gdf_valid_type* valid,
                                    gdf_size_type size, gdf_dtype dtype, gdf_size_type null_count) {
  gdf_column_view(column, data, valid, size, dtype);
  column->null_count = null_count;
  return GDF_SUCCESS;
}

// Releases buffers that the library handed out through rmmAlloc (join outputs).
gdf_error gdf_column_free(gdf_column* column) {
  B200_RMM_TRY(rmmFree(column->data, 0));
  B200_RMM_TRY(rmmFree(column->valid, 0));
  return GDF_SUCCESS;
}

gdf_error get_column_byte_width(gdf_column* col, int* width) {
  int w = b200::dtype_width(col->dtype);
  if (w == 0) {
    *width = -1;
    return GDF_UNSUPPORTED_DTYPE;
  }
  *width = w;
  return GDF_SUCCESS;
}

gdf_error gdf_context_view(gdf_context* context, int flag_sorted, gdf_method flag_method,
                           int flag_distinct, int flag_sort_result, int flag_sort_inplace) {
  context->flag_sorted = flag_sorted;
  context->flag_method = flag_method;
  context->flag_distinct = flag_distinct;
  context->flag_sort_result = flag_sort_result;
  context->flag_sort_inplace = flag_sort_inplace;
  return GDF_SUCCESS;
}

const char* gdf_error_get_name(gdf_error errcode) {
#define B200_NAME(x) \
  case x: return #x;
  switch (errcode) {
    B200_NAME(GDF_SUCCESS)
    B200_NAME(GDF_CUDA_ERROR)
    B200_NAME(GDF_UNSUPPORTED_DTYPE)
    B200_NAME(GDF_COLUMN_SIZE_MISMATCH)
    B200_NAME(GDF_COLUMN_SIZE_TOO_BIG)
    B200_NAME(GDF_DATASET_EMPTY)
    B200_NAME(GDF_VALIDITY_MISSING)
    B200_NAME(GDF_VALIDITY_UNSUPPORTED)
    B200_NAME(GDF_INVALID_API_CALL)
    B200_NAME(GDF_JOIN_DTYPE_MISMATCH)
    B200_NAME(GDF_JOIN_TOO_MANY_COLUMNS)
    B200_NAME(GDF_DTYPE_MISMATCH)
    B200_NAME(GDF_UNSUPPORTED_METHOD)
    B200_NAME(GDF_INVALID_AGGREGATOR)
    B200_NAME(GDF_INVALID_HASH_FUNCTION)
    B200_NAME(GDF_PARTITION_DTYPE_MISMATCH)
    B200_NAME(GDF_HASH_TABLE_INSERT_FAILURE)
    B200_NAME(GDF_UNSUPPORTED_JOIN_TYPE)
    B200_NAME(GDF_C_ERROR)
    B200_NAME(GDF_FILE_ERROR)
    B200_NAME(GDF_MEMORYMANAGER_ERROR)
    B200_NAME(GDF_UNDEFINED_NVTX_COLOR)
    B200_NAME(GDF_NULL_NVTX_NAME)
    default: return "Internal error. Unknown error code.";
  }
#undef B200_NAME
}

int gdf_cuda_last_error() { return cudaGetLastError(); }
const char* gdf_cuda_error_string(int cuda_error) { return cudaGetErrorString((cudaError_t)cuda_error); }
const char* gdf_cuda_error_name(int cuda_error) { return cudaGetErrorName((cudaError_t)cuda_error); }

// ---- NVTX (ref src/nvtx_utils.h:18, nvtx_utils.cpp:19-71): 9 fixed ARGB colours ----
static const uint32_t kNvtxColors[GDF_NUM_COLORS] = {0xff00ff00u, 0xff0000ffu, 0xffffff00u,
                                                     0xffff00ffu, 0xff00ffffu, 0xffff0000u,
                                                     0xffffffffu, 0xff006600u, 0xffffa500u};

gdf_error gdf_nvtx_range_push_hex(char const* const name, unsigned int color) {
  if (name == nullptr) return GDF_NULL_NVTX_NAME;
  nvtxEventAttributes_t attr = {};
  attr.version = NVTX_VERSION;
  attr.size = NVTX_EVENT_ATTRIB_STRUCT_SIZE;
  attr.colorType = NVTX_COLOR_ARGB;
  attr.color = color;
  attr.messageType = NVTX_MESSAGE_TYPE_ASCII;
  attr.message.ascii = name;
  nvtxRangePushEx(&attr);
  return GDF_SUCCESS;
}

gdf_error gdf_nvtx_range_push(char const* const name, gdf_color color) {
  if ((int)color < 0 || color >= GDF_NUM_COLORS) return GDF_UNDEFINED_NVTX_COLOR;
  return gdf_nvtx_range_push_hex(name, kNvtxColors[color]);
}

gdf_error gdf_nvtx_range_pop() {
  nvtxRangePop();
  return GDF_SUCCESS;
}

}  // extern "C"
