// Columnar reductions: sum / product / sum_squared / min / max over the valid rows of one column.
// Contract followed (reference libgdf/src/reductions.cu):
//   * null rows contribute the op's identity (:44-47); min identity = numeric_limits::max(),
//     max identity = numeric_limits::lowest() (:255-269); an empty / all-null column returns it
//   * arithmetic is carried out in the column's own C type - int8 sums wrap in int8 (:236-241)
//   * the answer is written to dev_result[0]; dev_result is caller scratch of dev_result_size
//     elements, gdf_reduce_optimal_output_size() says how many the caller should provide (:224)
//
// B200 design.  The reference launches at most 128 blocks x 128 threads with scalar loads and `int`
// indexing (:33-45,:107-109) - about a tenth of a B200's thread capacity.  Here: a persistent grid
// of sm_count x 4 CTAs x 512 threads, 128-bit no-allocate loads with 4 loads in flight per thread,
// validity bytes consumed inside the scan, warp shuffle + one smem stage per CTA, and a single
// launch: the last CTA to finish (atomic ticket) folds the per-CTA partials and writes
// dev_result[0].  Integer results are bit-exact (wrapping add/mul is associative); floating-point
// results differ from the reference only by summation order.
#include <limits>
#include <mutex>

#include "common.cuh"

namespace b200 {
namespace {

constexpr int kThreads = 512;
constexpr int kCtasPerSm = 4;
constexpr int kMaxPartials = 4096;

struct RSum { template <typename T> static __device__ T apply(T a, T b) { return a + b; } };
struct RProd { template <typename T> static __device__ T apply(T a, T b) { return a * b; } };
struct RMin { template <typename T> static __device__ T apply(T a, T b) { return a <= b ? a : b; } };
struct RMax { template <typename T> static __device__ T apply(T a, T b) { return a >= b ? a : b; } };

template <typename T, int N> struct alignas(16) Pack { T v[N]; };

// int8 arithmetic must wrap exactly like the reference's int8_t accumulators: do it in int and
// truncate on every step (identical mod 2^8).
template <typename T> struct Acc { using type = T; };

template <typename T, typename Op>
static __device__ __forceinline__ T warp_fold(T v) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    T o = __shfl_xor_sync(0xffffffffu, v, d);
    v = Op::apply(v, o);
  }
  return v;
}
template <typename Op>
static __device__ __forceinline__ int8_t warp_fold_i8(int8_t v) {
  int x = v;
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    int o = __shfl_xor_sync(0xffffffffu, x, d);
    x = (int8_t)Op::apply((int8_t)x, (int8_t)o);
  }
  return (int8_t)x;
}
template <typename T, typename Op> struct WarpFold { static __device__ T run(T v) { return warp_fold<T, Op>(v); } };
template <typename Op> struct WarpFold<int8_t, Op> { static __device__ int8_t run(int8_t v) { return warp_fold_i8<Op>(v); } };

template <typename T, typename Op>
static __device__ __forceinline__ T block_fold(T v, T identity, T* smem) {
  v = WarpFold<T, Op>::run(v);
  const int w = threadIdx.x >> 5;
  if (lane_id() == 0) smem[w] = v;
  __syncthreads();
  T r = identity;
  if (w == 0) {
    r = lane_id() < (kThreads >> 5) ? smem[lane_id()] : identity;
    r = WarpFold<T, Op>::run(r);
  }
  __syncthreads();
  return r;  // valid in warp 0
}

template <typename T, typename Op, bool SQUARE, bool VECTOR>
__global__ void __launch_bounds__(kThreads) reduce_kernel(const T* __restrict__ data,
                                                          const gdf_valid_type* __restrict__ mask,
                                                          size_t n, T identity, T* __restrict__ partials,
                                                          unsigned* __restrict__ ticket,
                                                          T* __restrict__ result) {
  constexpr int VEC = 16 / sizeof(T);
  __shared__ T smem[kThreads / 32];
  __shared__ bool is_last;
  T acc = identity;
  auto fold_elem = [&](T x) {
    if (SQUARE) x = x * x;
    acc = Op::apply(acc, x);
  };
  const size_t stride = (size_t)gridDim.x * kThreads;
  const size_t gtid = (size_t)blockIdx.x * kThreads + threadIdx.x;
  if (VECTOR) {
    const size_t nvec = n / VEC;
    const uint4* d4 = reinterpret_cast<const uint4*>(data);
    auto fold_vec = [&](const uint4& raw, size_t vi) {
      const Pack<T, VEC>& p = reinterpret_cast<const Pack<T, VEC>&>(raw);
      unsigned bits = 0xffffffffu;
      if (mask) {
        const size_t row = vi * VEC;  // VEC in {2,4,16}: rows of one vector never straddle oddly
        if (VEC <= 8) bits = (unsigned)mask[row >> 3] >> (row & 7);
        else bits = (unsigned)mask[row >> 3] | ((unsigned)mask[(row >> 3) + 1] << 8);
      }
#pragma unroll
      for (int e = 0; e < VEC; ++e)
        if ((bits >> e) & 1u) fold_elem(p.v[e]);
    };
    size_t i = gtid;
    for (; i + 3 * stride < nvec; i += 4 * stride) {
      uint4 a = ldg_stream(d4 + i), b = ldg_stream(d4 + i + stride), c = ldg_stream(d4 + i + 2 * stride),
            d = ldg_stream(d4 + i + 3 * stride);
      fold_vec(a, i); fold_vec(b, i + stride); fold_vec(c, i + 2 * stride); fold_vec(d, i + 3 * stride);
    }
    for (; i < nvec; i += stride) {
      uint4 a = ldg_stream(d4 + i);
      fold_vec(a, i);
    }
    if (gtid == 0)
      for (size_t k = nvec * VEC; k < n; ++k)
        if (bit_valid(mask, k)) fold_elem(data[k]);
  } else {
    for (size_t i = gtid; i < n; i += stride)
      if (bit_valid(mask, i)) fold_elem(data[i]);
  }

  T block_total = block_fold<T, Op>(acc, identity, smem);
  if (threadIdx.x == 0) {
    partials[blockIdx.x] = block_total;
    __threadfence();
    unsigned t = atomicAdd(ticket, 1u);
    is_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  // second stage: plain fold (sum_squared's second stage is a plain sum, ref reductions.cu:166-167)
  T v = identity;
  for (unsigned i = threadIdx.x; i < gridDim.x; i += kThreads) v = Op::apply(v, ((volatile T*)partials)[i]);
  T total = block_fold<T, Op>(v, identity, smem);
  if (threadIdx.x == 0) {
    *result = total;
    *ticket = 0;  // re-arm for the next call on this device
  }
}

// Per-device scratch for partials + ticket, allocated once.  Calls are stream-ordered on the
// default stream, so consecutive reductions reuse it safely.
struct DeviceScratch {
  void* partials = nullptr;  // kMaxPartials * 8 bytes
  unsigned* ticket = nullptr;
};
cudaError_t get_scratch(DeviceScratch** out) {
  static std::mutex mu;
  static DeviceScratch per_dev[64];
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
  std::lock_guard<std::mutex> g(mu);
  DeviceScratch& s = per_dev[dev];
  if (!s.partials) {
    void* p = nullptr;
    e = cudaMalloc(&p, kMaxPartials * 8 + 256);
    if (e != cudaSuccess) return e;
    e = cudaMemset(p, 0, kMaxPartials * 8 + 256);
    if (e != cudaSuccess) return e;
    s.partials = p;
    s.ticket = reinterpret_cast<unsigned*>(static_cast<char*>(p) + kMaxPartials * 8);
  }
  *out = &s;
  return cudaSuccess;
}

template <typename T, typename Op, bool SQUARE>
gdf_error launch_reduce(gdf_column* col, T identity, T* dev_result, gdf_size_type dev_result_size) {
  B200_REQUIRE(col != nullptr && dev_result != nullptr, GDF_DATASET_EMPTY);
  B200_REQUIRE(dev_result_size >= 1, GDF_COLUMN_SIZE_MISMATCH);
  DeviceScratch* s = nullptr;
  B200_CUDA_TRY(get_scratch(&s));
  const size_t n = col->size;
  const T* data = static_cast<const T*>(col->data);
  constexpr int VEC = 16 / sizeof(T);
  int blocks = sm_count() * kCtasPerSm;
  if (blocks > kMaxPartials) blocks = kMaxPartials;
  size_t need = (n / VEC + kThreads - 1) / kThreads;
  if (need < 1) need = 1;
  if ((size_t)blocks > need) blocks = (int)need;
  T* partials = static_cast<T*>(s->partials);
  B200_TIMED("reduce");
  if (aligned16(data))
    reduce_kernel<T, Op, SQUARE, true><<<blocks, kThreads>>>(data, col->valid, n, identity, partials, s->ticket, dev_result);
  else
    reduce_kernel<T, Op, SQUARE, false><<<blocks, kThreads>>>(data, col->valid, n, identity, partials, s->ticket, dev_result);
  B200_CHECK_LAST();
  return GDF_SUCCESS;
}

}  // namespace
}  // namespace b200

using namespace b200;

extern "C" unsigned int gdf_reduce_optimal_output_size() { return 128; }  // ref reductions.cu:8,224

#define B200_REDUCE(NAME, OP, SQ, SUFFIX, T, ID)                                                      \
  extern "C" gdf_error gdf_##NAME##_##SUFFIX(gdf_column* col, T* dev_result, gdf_size_type n) {       \
    return launch_reduce<T, OP, SQ>(col, (T)(ID), dev_result, n);                                     \
  }

#define B200_REDUCE_GENERIC_NUM(NAME)                                                                 \
  extern "C" gdf_error gdf_##NAME##_generic(gdf_column* col, void* r, gdf_size_type n) {              \
    switch (col->dtype) { /* ref reductions.cu:195-206 */                                             \
      case GDF_FLOAT64: return gdf_##NAME##_f64(col, (double*)r, n);                                  \
      case GDF_FLOAT32: return gdf_##NAME##_f32(col, (float*)r, n);                                   \
      case GDF_INT64: return gdf_##NAME##_i64(col, (int64_t*)r, n);                                   \
      case GDF_INT32: return gdf_##NAME##_i32(col, (int32_t*)r, n);                                   \
      case GDF_INT8: return gdf_##NAME##_i8(col, (int8_t*)r, n);                                      \
      default: return GDF_UNSUPPORTED_DTYPE;                                                          \
    }                                                                                                 \
  }

B200_REDUCE(sum, RSum, false, f64, double, 0)
B200_REDUCE(sum, RSum, false, f32, float, 0)
B200_REDUCE(sum, RSum, false, i64, int64_t, 0)
B200_REDUCE(sum, RSum, false, i32, int32_t, 0)
B200_REDUCE(sum, RSum, false, i8, int8_t, 0)
B200_REDUCE_GENERIC_NUM(sum)

B200_REDUCE(product, RProd, false, f64, double, 1)
B200_REDUCE(product, RProd, false, f32, float, 1)
B200_REDUCE(product, RProd, false, i64, int64_t, 1)
B200_REDUCE(product, RProd, false, i32, int32_t, 1)
B200_REDUCE(product, RProd, false, i8, int8_t, 1)
B200_REDUCE_GENERIC_NUM(product)

B200_REDUCE(sum_squared, RSum, true, f64, double, 0)
B200_REDUCE(sum_squared, RSum, true, f32, float, 0)
extern "C" gdf_error gdf_sum_squared_generic(gdf_column* col, void* r, gdf_size_type n) {
  switch (col->dtype) {  // ref reductions.cu:208-216
    case GDF_FLOAT64: return gdf_sum_squared_f64(col, (double*)r, n);
    case GDF_FLOAT32: return gdf_sum_squared_f32(col, (float*)r, n);
    default: return GDF_UNSUPPORTED_DTYPE;
  }
}

B200_REDUCE(min, RMin, false, f64, double, std::numeric_limits<double>::max())
B200_REDUCE(min, RMin, false, f32, float, std::numeric_limits<float>::max())
B200_REDUCE(min, RMin, false, i64, int64_t, std::numeric_limits<int64_t>::max())
B200_REDUCE(min, RMin, false, i32, int32_t, std::numeric_limits<int32_t>::max())
B200_REDUCE(min, RMin, false, i8, int8_t, std::numeric_limits<int8_t>::max())
B200_REDUCE_GENERIC_NUM(min)

B200_REDUCE(max, RMax, false, f64, double, std::numeric_limits<double>::lowest())
B200_REDUCE(max, RMax, false, f32, float, std::numeric_limits<float>::lowest())
B200_REDUCE(max, RMax, false, i64, int64_t, std::numeric_limits<int64_t>::lowest())
B200_REDUCE(max, RMax, false, i32, int32_t, std::numeric_limits<int32_t>::lowest())
B200_REDUCE(max, RMax, false, i8, int8_t, std::numeric_limits<int8_t>::lowest())
B200_REDUCE_GENERIC_NUM(max)

// ---- gdf_count_nonzero_mask (ref validops.cu:138-185): number of set bits among the first num_rows
// bits of a validity bitmap.  Bytes are read one by one (the buffer is only guaranteed ceil(n/8)
// bytes, ref include/gdf/utils.h:21-23) and the ragged last byte is masked. ----
namespace b200 {
namespace {
__global__ void __launch_bounds__(256) count_bits_kernel(const gdf_valid_type* __restrict__ mask, size_t rows,
                                                         unsigned long long* __restrict__ out) {
  const size_t nbytes = (rows + 7) / 8;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  unsigned local = 0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nbytes; i += stride) {
    const unsigned live = (i == nbytes - 1 && (rows & 7)) ? ((1u << (rows & 7)) - 1u) : 0xffu;
    local += __popc((unsigned)mask[i] & live);
  }
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) local += __shfl_xor_sync(0xffffffffu, local, s);
  if (lane_id() == 0 && local) atomicAdd(out, (unsigned long long)local);
}
}  // namespace
}  // namespace b200

extern "C" gdf_error gdf_count_nonzero_mask(gdf_valid_type const* masks, int num_rows, int* count) {
  if (masks == nullptr || count == nullptr) return GDF_DATASET_EMPTY;
  if (num_rows == 0) return GDF_SUCCESS;
  B200_REQUIRE(num_rows > 0, GDF_INVALID_API_CALL);
  Scratch d;
  B200_CUDA_TRY(d.alloc(sizeof(unsigned long long)));
  B200_CUDA_TRY(cudaMemsetAsync(d.ptr, 0, sizeof(unsigned long long), 0));
  const size_t nbytes = valid_bytes((size_t)num_rows);
  size_t want = (nbytes + 255) / 256;
  const size_t cap = (size_t)sm_count() * 8;
  count_bits_kernel<<<(int)(want < cap ? want : cap), 256>>>(masks, (size_t)num_rows, d.as<unsigned long long>());
  B200_CHECK_LAST();
  unsigned long long h = 0;
  B200_CUDA_TRY(cudaMemcpy(&h, d.ptr, sizeof(h), cudaMemcpyDeviceToHost));
  *count = (int)h;
  return GDF_SUCCESS;
}
