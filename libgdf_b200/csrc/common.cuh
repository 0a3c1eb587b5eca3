// Shared host/device helpers for the sm_100a gdf_* hot path.
//
// Everything here is written for one target: B200 (sm_100a, 148 SMs, HBM3e).  The kernels in this
// directory are HBM-bound integer / indexing scans, so the helpers are about (a) 128-bit coalesced
// streaming loads that bypass L1 allocation, (b) Arrow LSB-first validity bits consumed inside the
// scan, (c) grid sizing in multiples of the SM count and (d) stream-ordered scratch memory so that
// no API call pays a cudaMalloc/cudaFree device synchronisation.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>

#include <gdf/gdf.h>
#include <rmm.h>

namespace b200 {

// ----------------------------------------------------------------------------------------------
// error plumbing (same observable behaviour as reference include/gdf/errorutils.h:8-29: a failing
// CUDA call turns into GDF_CUDA_ERROR and the CUDA error stays retrievable through
// gdf_cuda_last_error()).
// ----------------------------------------------------------------------------------------------
#define B200_CUDA_TRY(call)                                   \
  do {                                                        \
    cudaError_t b200_status__ = (call);                       \
    if (b200_status__ != cudaSuccess) return GDF_CUDA_ERROR;  \
  } while (0)
#define B200_RMM_TRY(call)                                        \
  do {                                                            \
    if ((call) != RMM_SUCCESS) return GDF_MEMORYMANAGER_ERROR;    \
  } while (0)
#define B200_REQUIRE(cond, err) \
  do {                          \
    if (!(cond)) return (err);  \
  } while (0)
#define B200_CHECK_LAST() B200_CUDA_TRY(cudaPeekAtLastError())

// ----------------------------------------------------------------------------------------------
// device properties (cached per device)
// ----------------------------------------------------------------------------------------------
int sm_count();  // 148 on B200; queried once per device
int select_dealing_mode();      // gdfx_set_select_dealing (include/gdf_b200_ext.h): 0 cooperative + static, 1 ticket
bool cooperative_launch_ok();  // cudaDevAttrCooperativeLaunch: a cooperative grid is only started once ALL its CTAs fit

// Width in bytes of a gdf dtype, 0 if the dtype has no fixed width on this path
// (ref src/column.cpp:237-275).
__host__ __device__ inline int dtype_width(int dtype) {
  switch (dtype) {
    case GDF_INT8: return 1;
    case GDF_INT16: return 2;
    case GDF_INT32: case GDF_FLOAT32: case GDF_DATE32: return 4;
    case GDF_INT64: case GDF_FLOAT64: case GDF_DATE64: case GDF_TIMESTAMP: return 8;
    default: return 0;
  }
}

inline size_t valid_bytes(size_t rows) { return (rows + 7) / 8; }  // ref include/gdf/utils.h:21-23

__host__ __device__ inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// ----------------------------------------------------------------------------------------------
// Scratch.  Library-internal temporaries (hash tables, partition buffers, look-back descriptors) come from a
// caching block allocator (block_cache.h): a freed block is kept and handed back for the next request of about
// the same size, so after the first call of a given shape allocation never enters the driver and no API call
// pays a cudaMalloc/cudaFree device synchronisation.  The cache is bounded (half of the device's memory by
// default, gdfx_set_scratch_limit), is handed back to the driver when any allocation of this library or of
// librmm's pool runs out of memory, and can be emptied by the caller (gdfx_trim_scratch).  Everything the library
// launches runs on the legacy default stream, so reuse is ordered by that stream.
// Library-OWNED OUTPUTS (join index columns, result_cols) go through rmmAlloc instead (output_alloc), because
// the caller releases them with gdf_column_free -> rmmFree (ref src/column.cpp:222-227).
// ----------------------------------------------------------------------------------------------
cudaError_t scratch_alloc(void** p, size_t bytes, cudaStream_t s);
cudaError_t scratch_free(void* p, cudaStream_t s);
rmmError_t output_alloc(void** p, size_t bytes);  // rmmAlloc(stream 0); trims the scratch cache and retries on failure

struct Scratch {  // RAII wrapper; frees in stream order
  void* ptr = nullptr;
  cudaStream_t stream = 0;
  Scratch() = default;
  Scratch(const Scratch&) = delete;
  Scratch& operator=(const Scratch&) = delete;
  cudaError_t alloc(size_t bytes, cudaStream_t s = 0) {
    release();
    stream = s;
    return scratch_alloc(&ptr, bytes ? bytes : 1, s);
  }
  void release() {
    if (ptr) scratch_free(ptr, stream);
    ptr = nullptr;
  }
  template <typename T> T* as() const { return static_cast<T*>(ptr); }
  ~Scratch() { release(); }
};

// Optional per-kernel timing (gdfx_profile_enable / gdfx_profile_report, include/gdf_b200_ext.h).
// B200_TIMED("name") brackets the launches in the enclosing scope with events on stream 0 (B200_TIMED_ON: on the
// given stream - work that runs on the library's private stream beside the legacy stream).
struct KernelTimer {
  explicit KernelTimer(const char* name, cudaStream_t s = 0);
  ~KernelTimer();
  int slot;
  cudaStream_t stream;
};
#define B200_TIMED(name) ::b200::KernelTimer b200_kernel_timer__(name)
#define B200_TIMED_ON(name, s) ::b200::KernelTimer b200_kernel_timer__(name, s)

// Lab knobs: A/B switches of kernel variants for tools/lab_*.py.  The shipped library has ONE behaviour: unless
// it is built with -DB200_LAB (make LAB=1 -> lib_lab/), a knob is its compile-time default and no environment
// variable is ever read.
#ifdef B200_LAB
int lab_knob(const char* name, int dflt);  // atoi(getenv(name)) or dflt
#else
inline int lab_knob(const char*, int dflt) { return dflt; }
#endif

// Small pinned host mailbox for "count" read-backs (one per host thread).
void* pinned_mailbox();  // >= 256 bytes, cudaHostAlloc'd

#ifdef __CUDACC__
// ----------------------------------------------------------------------------------------------
// device-side building blocks
// ----------------------------------------------------------------------------------------------
static __device__ __forceinline__ unsigned lane_id() { return threadIdx.x & 31u; }
static __device__ __forceinline__ unsigned lanemask_lt() {
  unsigned m;
  asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
  return m;
}

// 128-bit streaming load: read-only path, do not allocate in L1 (each byte is used exactly once).
static __device__ __forceinline__ uint4 ldg_stream(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
static __device__ __forceinline__ void stg_stream(void* p, const uint4& v) {
  asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y),
               "r"(v.z), "r"(v.w)
               : "memory");
}

// Arrow validity: bit i of the mask, LSB first (ref include/gdf/utils.h:10-15). nullptr == all valid.
static __device__ __forceinline__ bool bit_valid(const gdf_valid_type* m, size_t i) {
  return m == nullptr || ((m[i >> 3] >> (i & 7)) & 1);
}

// MurmurHash3_x86_32, seed 0, over the little-endian bytes of one fixed-width value; this is the
// published algorithm the reference instantiates per column type
// (ref src/hashmap/hash_functions.cuh:31-121).  W = value width in bytes (1, 2, 4 or 8).
static __host__ __device__ __forceinline__ uint32_t rotl32(uint32_t x, int r) {
  return (x << r) | (x >> (32 - r));
}
static __host__ __device__ __forceinline__ uint32_t fmix32(uint32_t h) {
  h ^= h >> 16; h *= 0x85ebca6bu; h ^= h >> 13; h *= 0xc2b2ae35u; h ^= h >> 16;
  return h;
}
static __host__ __device__ __forceinline__ uint32_t murmur_block(uint32_t h1, uint32_t k1) {
  k1 *= 0xcc9e2d51u; k1 = rotl32(k1, 15); k1 *= 0x1b873593u;
  h1 ^= k1; h1 = rotl32(h1, 13); h1 = h1 * 5u + 0xe6546b64u;
  return h1;
}
template <int W>
static __host__ __device__ __forceinline__ uint32_t murmur3_32(uint64_t bits) {
  uint32_t h1 = 0;
  if (W == 8) {
    h1 = murmur_block(h1, (uint32_t)bits);
    h1 = murmur_block(h1, (uint32_t)(bits >> 32));
  } else if (W == 4) {
    h1 = murmur_block(h1, (uint32_t)bits);
  } else {  // 1- or 2-byte tail
    uint32_t k1 = (uint32_t)bits & (W == 2 ? 0xffffu : 0xffu);
    k1 *= 0xcc9e2d51u; k1 = rotl32(k1, 15); k1 *= 0x1b873593u;
    h1 ^= k1;
  }
  h1 ^= (uint32_t)W;
  return fmix32(h1);
}
// boost-style combine used for columns after the first (ref hash_functions.cuh:66-72).
static __host__ __device__ __forceinline__ uint32_t hash_combine(uint32_t lhs, uint32_t rhs) {
  return lhs ^ (rhs + 0x9e3779b9u + (lhs << 6) + (lhs >> 2));
}
#endif  // __CUDACC__

}  // namespace b200
