// Host-only plumbing of the gdf_* ABI: column/context views, error names, CUDA error accessors,
// NVTX ranges, plus the library-internal scratch allocator.  Reference behaviour followed:
//   libgdf/src/column.cpp:160-275, src/context.cpp:3-11, src/errorhandling.cpp:5-35,
//   src/cudautils.cu:4-14, src/nvtx_utils.cpp:19-71.
#include <map>
#include <mutex>
#include <string>
#include <vector>
#include <cstdio>
#include <cstring>
#include <cstdlib>

#include <nvtx3/nvToolsExt.h>

#include "block_cache.h"
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

static int g_select_dealing = 0;
int select_dealing_mode() { return g_select_dealing; }

bool cooperative_launch_ok() {
  static int cached[64] = {0};  // 0 unknown, 1 yes, 2 no
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return false;
  if (cached[dev] == 0) {
    int v = 0;
    cached[dev] = (cudaDeviceGetAttribute(&v, cudaDevAttrCooperativeLaunch, dev) == cudaSuccess && v) ? 1 : 2;
  }
  return cached[dev] == 1;
}

static BlockCache& scratch_cache() {
  static BlockCache* c = new BlockCache();  // leaked on purpose: no teardown-order issues at exit
  return *c;
}
extern "C" void rmmxTrimPool();  // librmm.so (rmm.cpp)

cudaError_t scratch_alloc(void** p, size_t bytes, cudaStream_t /*s*/) {
  cudaError_t e = scratch_cache().alloc(p, bytes);  // trims its own cache and retries on out-of-memory
  if (e == cudaErrorMemoryAllocation) {             // still no room: blocks parked in the rmm pool are the other candidate
    cudaGetLastError();
    rmmxTrimPool();
    e = scratch_cache().alloc(p, bytes);
  }
  return e;
}
cudaError_t scratch_free(void* p, cudaStream_t /*s*/) {
  return scratch_cache().release(p) ? cudaSuccess : cudaFree(p);
}

// Library-owned outputs (join index columns, result_cols): rmmAlloc on the default stream.  In rmm's default mode
// that is a plain cudaMalloc, which can fail while gigabytes sit in this library's scratch cache - so on failure
// the cache is handed back to the driver and the request repeated once.
rmmError_t output_alloc(void** p, size_t bytes) {
  rmmError_t r = rmmAlloc(p, bytes, 0);
  if (r == RMM_ERROR_OUT_OF_MEMORY || r == RMM_ERROR_CUDA_ERROR) {
    cudaGetLastError();
    scratch_cache().trim();
    r = rmmAlloc(p, bytes, 0);
  }
  return r;
}

#ifdef B200_LAB
int lab_knob(const char* name, int dflt) {
  const char* v = getenv(name);
  return v ? atoi(v) : dflt;
}
#endif

void* pinned_mailbox() {
  static thread_local void* box = nullptr;
  if (!box) {
    if (cudaHostAlloc(&box, 256, cudaHostAllocDefault) != cudaSuccess) box = nullptr;
  }
  return box;
}

// ---- per-kernel timing ----
namespace {
struct Pending {
  const char* name;
  cudaEvent_t start, stop;
};
struct Profiler {
  std::mutex mu;
  bool enabled = false;
  std::vector<Pending> pending;
  std::vector<cudaEvent_t> pool;
  std::map<std::string, std::pair<long long, double>> totals;  // name -> {launches, ms}
  cudaEvent_t get() {
    if (!pool.empty()) {
      cudaEvent_t e = pool.back();
      pool.pop_back();
      return e;
    }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
  }
};
Profiler& prof() {
  static Profiler p;
  return p;
}
}  // namespace

KernelTimer::KernelTimer(const char* name, cudaStream_t s) : slot(-1), stream(s) {
  Profiler& p = prof();
  if (!p.enabled) return;
  std::lock_guard<std::mutex> g(p.mu);
  Pending pe{name, p.get(), p.get()};
  if (!pe.start || !pe.stop) return;
  cudaEventRecord(pe.start, stream);
  slot = (int)p.pending.size();
  p.pending.push_back(pe);
}
KernelTimer::~KernelTimer() {
  if (slot < 0) return;
  Profiler& p = prof();
  std::lock_guard<std::mutex> g(p.mu);
  if (slot < (int)p.pending.size()) cudaEventRecord(p.pending[slot].stop, stream);
}

}  // namespace b200

extern "C" size_t gdfx_trim_scratch() {
  const size_t had = b200::scratch_cache().cached_bytes();
  b200::scratch_cache().trim();
  return had;
}
extern "C" size_t gdfx_scratch_cached_bytes() { return b200::scratch_cache().cached_bytes(); }
extern "C" void gdfx_set_scratch_limit(size_t bytes) { b200::scratch_cache().set_limit(bytes); }

extern "C" int gdfx_set_select_dealing(int mode) {
  const int prev = b200::g_select_dealing;
  b200::g_select_dealing = mode ? 1 : 0;
  return prev;
}

// TEST helper (include/gdf_b200_ext.h): CTAs that fill one SM each, parked for a while on a private non-blocking stream
static __global__ void __launch_bounds__(1024, 1) occupy_kernel(unsigned long long ns) {
  extern __shared__ unsigned char occupy_smem[];
  unsigned long long t0, t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  do {
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  } while (t - t0 < ns);
  if (ns == ~0ull) occupy_smem[threadIdx.x] = 0;  // keeps the shared-memory allocation alive
}
extern "C" gdf_error gdfx_debug_occupy_sms(int blocks, unsigned microseconds) {
  if (blocks <= 0) return GDF_SUCCESS;
  static cudaStream_t side = nullptr;
  if (side == nullptr) B200_CUDA_TRY(cudaStreamCreateWithFlags(&side, cudaStreamNonBlocking));
  const int smem = 200 * 1024;
  B200_CUDA_TRY(cudaFuncSetAttribute(occupy_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  occupy_kernel<<<blocks, 1024, smem, side>>>((unsigned long long)microseconds * 1000ull);
  B200_CHECK_LAST();
  return GDF_SUCCESS;
}

extern "C" int gdfx_profile_enable(int on) {
  b200::Profiler& p = b200::prof();
  std::lock_guard<std::mutex> g(p.mu);
  const int was = p.enabled;
  p.enabled = on != 0;
  return was;
}

extern "C" size_t gdfx_profile_report(char* buf, size_t capacity) {
  b200::Profiler& p = b200::prof();
  std::lock_guard<std::mutex> g(p.mu);
  cudaDeviceSynchronize();
  for (const b200::Pending& pe : p.pending) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, pe.start, pe.stop) == cudaSuccess) {
      auto& t = p.totals[pe.name];
      t.first += 1;
      t.second += ms;
    }
    p.pool.push_back(pe.start);
    p.pool.push_back(pe.stop);
  }
  cudaGetLastError();
  p.pending.clear();
  std::string out = "{";
  bool first = true;
  for (const auto& kv : p.totals) {
    char line[256];
    snprintf(line, sizeof line, "%s\"%s\": {\"launches\": %lld, \"ms\": %.6f}", first ? "" : ", ", kv.first.c_str(),
             kv.second.first, kv.second.second);
    out += line;
    first = false;
  }
  out += "}";
  p.totals.clear();
  if (buf && capacity) {
    const size_t n = out.size() < capacity - 1 ? out.size() : capacity - 1;
    memcpy(buf, out.data(), n);
    buf[n] = 0;
  }
  return out.size() + 1;
}

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
