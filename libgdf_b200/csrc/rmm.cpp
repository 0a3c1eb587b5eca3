// librmm.so - the rmm* allocator ABI (reference: libgdf/include/memory.h:65-184,
// libgdf/src/memory/memory.cpp:120-299) re-designed for CUDA 12 / B200.
//
// The reference sub-allocates one big cudaMalloc with the vendored cnmem pool.  On CUDA 12 the
// driver's stream-ordered allocator does that job natively, so:
//   * CudaDefaultAllocation  -> cudaMalloc / cudaFree (same as the reference's default mode)
//   * PoolAllocation         -> a caching allocator (block_cache.h): freed blocks are kept and handed
//                               back for the next request of about the same size, so warm operator
//                               calls never enter the driver - the property cnmem gives the
//                               reference.  An initial_pool_size > 0 pre-warms one block of that size.
//                               Reuse is stream-aware like cnmem's per-stream child pools: a block freed on
//                               stream A goes to a later request on A at once and to a request on another
//                               stream only behind an event recorded on A at free time.
// A pointer from either mode may be released in either mode (rmmFree falls back to cudaFree for
// pointers the cache does not own), so changing mode between alloc and free is safe.
// The optional event log keeps the reference's CSV schema (memory_manager.cpp:46-64).
#include <cuda_runtime_api.h>

#include <chrono>
#include <cstring>
#include <fstream>
#include <mutex>
#include <set>
#include <sstream>
#include <vector>

#include <rmm.h>

#include "block_cache.h"

namespace {

using Clock = std::chrono::system_clock;

struct Event {
  int kind;  // 0 Alloc, 1 Realloc, 2 Free
  int device;
  void* ptr;
  size_t size;
  cudaStream_t stream;
  size_t live;
  Clock::time_point start, end;
};

struct Manager {
  std::mutex mu;
  rmmOptions_t opts{CudaDefaultAllocation, 0, false};
  std::vector<Event> events;
  std::set<void*> live;
  Clock::time_point base = Clock::now();
};

Manager& mgr() {
  static Manager m;
  return m;
}

bool pool_mode() { return mgr().opts.allocation_mode == PoolAllocation; }

rmmError_t from_cuda(cudaError_t e) {
  if (e == cudaSuccess) return RMM_SUCCESS;
  if (e == cudaErrorMemoryAllocation) {
    cudaGetLastError();  // OOM is reported through the return code, keep the error state clean
    return RMM_ERROR_OUT_OF_MEMORY;
  }
  return RMM_ERROR_CUDA_ERROR;
}

b200::BlockCache& pool() {
  static b200::BlockCache* c = new b200::BlockCache(/*track_streams=*/true);  // leaked on purpose (exit-order safety)
  return *c;
}

struct LogScope {
  int kind;
  void* ptr;
  size_t size;
  cudaStream_t stream;
  bool on;
  Clock::time_point start;
  LogScope(int k, void* p, size_t s, cudaStream_t st) : kind(k), ptr(p), size(s), stream(st) {
    on = mgr().opts.enable_logging;
    if (on) start = Clock::now();
  }
  ~LogScope() {
    if (!on) return;
    Manager& m = mgr();
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> g(m.mu);
    if (kind == 0) m.live.insert(ptr);
    if (kind == 2) m.live.erase(ptr);
    m.events.push_back({kind, dev, ptr, size, stream, m.live.size(), start, Clock::now()});
  }
};

void write_csv(std::ostream& os) {
  Manager& m = mgr();
  std::lock_guard<std::mutex> g(m.mu);
  os << "Event Type,Device ID,Address,Stream,Size (bytes),Free Memory,Total Memory,Current Allocs,"
        "Start,End,Elapsed\n";
  static const char* names[] = {"Alloc", "Realloc", "Free"};
  for (const Event& e : m.events) {
    std::chrono::duration<double> t0 = e.start - m.base, t1 = e.end - m.base, dt = e.end - e.start;
    os << names[e.kind] << ',' << e.device << ',' << e.ptr << ',' << (void*)e.stream << ',' << e.size
       << ",0,0," << e.live << ',' << t0.count() << ',' << t1.count() << ',' << dt.count() << '\n';
  }
}

}  // namespace

extern "C" {

const char* rmmGetErrorString(rmmError_t errcode) {
  switch (errcode) {
    case RMM_SUCCESS: return "RMM_SUCCESS";
    case RMM_ERROR_CUDA_ERROR: return "RMM_ERROR_CUDA_ERROR";
    case RMM_ERROR_INVALID_ARGUMENT: return "RMM_ERROR_INVALID_ARGUMENT";
    case RMM_ERROR_NOT_INITIALIZED: return "RMM_ERROR_NOT_INITIALIZED";
    case RMM_ERROR_OUT_OF_MEMORY: return "RMM_ERROR_OUT_OF_MEMORY";
    case RMM_ERROR_UNKNOWN: return "RMM_ERROR_UNKNOWN";
    case RMM_ERROR_IO: return "RMM_ERROR_IO";
    default: return "Internal error. Unknown error code.";
  }
}

rmmError_t rmmInitialize(rmmOptions_t* options) {
  if (options) {
    std::lock_guard<std::mutex> g(mgr().mu);
    mgr().opts = *options;
  }
  if (pool_mode()) {
    size_t warm = mgr().opts.initial_pool_size;
    if (warm) {  // reserve the requested amount up front and park it in the cache
      void* p = nullptr;
      cudaError_t e = pool().alloc(&p, warm);
      if (e != cudaSuccess) return from_cuda(e);
      pool().release(p);
    }
  }
  return RMM_SUCCESS;
}

rmmError_t rmmFinalize() {
  Manager& m = mgr();
  std::lock_guard<std::mutex> g(m.mu);
  pool().trim();  // give cached blocks back to the driver
  m.events.clear();
  m.live.clear();
  m.opts = rmmOptions_t{CudaDefaultAllocation, 0, false};
  return RMM_SUCCESS;
}

rmmError_t rmmAlloc(void** ptr, size_t size, cudaStream_t stream) {
  if (!ptr && !size) return RMM_SUCCESS;
  if (!ptr) return RMM_ERROR_INVALID_ARGUMENT;
  LogScope log(0, nullptr, size, stream);
  const cudaError_t e = pool_mode() ? pool().alloc(ptr, size, stream) : cudaMalloc(ptr, size);
  if (e != cudaSuccess) return from_cuda(e);
  log.ptr = *ptr;
  return RMM_SUCCESS;
}

rmmError_t rmmFree(void* ptr, cudaStream_t stream) {
  LogScope log(2, ptr, 0, stream);
  if (!ptr) return RMM_SUCCESS;  // cudaFree(nullptr) is a no-op in the reference as well
  if (pool().release(ptr, stream)) return RMM_SUCCESS;  // block came from the cache (whatever the mode is now)
  return from_cuda(cudaFree(ptr));
}

// Same contract as the reference (memory.cpp:172-194): the old block is released, contents are
// NOT preserved.
rmmError_t rmmRealloc(void** ptr, size_t new_size, cudaStream_t stream) {
  if (!ptr && !new_size) return RMM_SUCCESS;
  if (!ptr) return RMM_ERROR_INVALID_ARGUMENT;
  LogScope log(1, nullptr, new_size, stream);
  rmmError_t r = RMM_SUCCESS;
  if (*ptr && !pool().release(*ptr, stream)) {
    if ((r = from_cuda(cudaFree(*ptr))) != RMM_SUCCESS) return r;
  }
  const cudaError_t e = pool_mode() ? pool().alloc(ptr, new_size, stream) : cudaMalloc(ptr, new_size);
  if ((r = from_cuda(e)) != RMM_SUCCESS) return r;
  log.ptr = *ptr;
  return RMM_SUCCESS;
}

// Offset of ptr inside its underlying allocation (used for IPC handles, memory.cpp:206-221).
// Without a sub-allocating pool every rmm block is its own allocation as far as the runtime API can
// tell, so the offset is computed from the allocation base reported by the driver through
// cudaPointerGetAttributes' devicePointer of the range start when available; otherwise 0.
rmmError_t rmmGetAllocationOffset(offset_t* offset, void* ptr, cudaStream_t /*stream*/) {
  if (!offset || !ptr) return RMM_ERROR_INVALID_ARGUMENT;
  cudaPointerAttributes attr;
  if (cudaPointerGetAttributes(&attr, ptr) != cudaSuccess || attr.type != cudaMemoryTypeDevice) {
    cudaGetLastError();
    return RMM_ERROR_INVALID_ARGUMENT;
  }
  *offset = 0;
  return RMM_SUCCESS;
}

rmmError_t rmmGetInfo(size_t* freeSize, size_t* totalSize, cudaStream_t /*stream*/) {
  if (!freeSize || !totalSize) return RMM_ERROR_INVALID_ARGUMENT;
  return from_cuda(cudaMemGetInfo(freeSize, totalSize));
}

// Extension (not in the reference ABI): give every cached pool block back to the driver.  libgdf.so calls it
// before it reports an out-of-memory condition for its own scratch.
void rmmxTrimPool() { pool().trim(); }
size_t rmmxPoolCachedBytes() { return pool().cached_bytes(); }

rmmError_t rmmWriteLog(const char* filename) {
  if (!filename) return RMM_ERROR_INVALID_ARGUMENT;
  std::ofstream csv(filename);
  if (!csv) return RMM_ERROR_IO;
  write_csv(csv);
  return csv ? RMM_SUCCESS : RMM_ERROR_IO;
}

size_t rmmLogSize() {
  std::ostringstream csv;
  write_csv(csv);
  return csv.str().size();
}

rmmError_t rmmGetLog(char* buffer, size_t buffer_size) {
  if (!buffer) return RMM_ERROR_INVALID_ARGUMENT;
  std::ostringstream csv;
  write_csv(csv);
  const std::string s = csv.str();
  std::memcpy(buffer, s.data(), s.size() < buffer_size ? s.size() : buffer_size);
  return RMM_SUCCESS;
}

}  // extern "C"
