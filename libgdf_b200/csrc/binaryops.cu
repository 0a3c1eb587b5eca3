// Element-wise binary ops over equal-dtype columns: arithmetic, comparison -> int8, bitwise, and
// gdf_validity_and.  Contract followed (reference libgdf/src/binaryops.cu):
//   * empty input -> GDF_SUCCESS without touching anything (:38-41)
//   * size mismatch -> GDF_COLUMN_SIZE_MISMATCH, lhs/rhs dtype mismatch -> GDF_UNSUPPORTED_DTYPE (:43-45)
//   * arithmetic output dtype must equal lhs dtype (:83), logical output must be GDF_INT8 (:92)
//   * with a validity mask on either side only lanes valid on BOTH sides are written; the other
//     output lanes are left untouched and output->valid is not written (:21-25)
//   * integer floordiv goes through double: floor((double)l / (double)r) (:143-149)
//
// B200 design: one streaming kernel per (type, op).  Without masks every thread issues 2 x UNROLL
// independent 128-bit no-allocate loads before its first use (>= 128 B in flight per thread) and the
// grid is a multiple of the SM count running a grid-stride loop; with masks the validity bytes are
// read inside the same pass (never materialised).  Work is stream-ordered on the legacy default
// stream: no device-wide synchronisation per call (the reference's cudaDeviceSynchronize, :73, is
// what makes 1M-row adds latency-bound).
#include <cmath>

#include "common.cuh"

namespace b200 {
namespace {

constexpr int kThreads = 256;
constexpr int kUnroll = 4;

template <typename T> struct OpAdd { static __device__ T apply(T a, T b) { return a + b; } };
template <typename T> struct OpSub { static __device__ T apply(T a, T b) { return a - b; } };
template <typename T> struct OpMul { static __device__ T apply(T a, T b) { return a * b; } };
template <typename T> struct OpDiv { static __device__ T apply(T a, T b) { return a / b; } };
template <typename T> struct OpFloorDiv {  // integers: via double, as the reference does
  static __device__ T apply(T a, T b) { return (T)floor((double)a / (double)b); }
};
template <> struct OpFloorDiv<float> { static __device__ float apply(float a, float b) { return floorf(a / b); } };
template <> struct OpFloorDiv<double> { static __device__ double apply(double a, double b) { return floor(a / b); } };
template <typename T> struct OpGt { static __device__ int8_t apply(T a, T b) { return a > b; } };
template <typename T> struct OpGe { static __device__ int8_t apply(T a, T b) { return a >= b; } };
template <typename T> struct OpLt { static __device__ int8_t apply(T a, T b) { return a < b; } };
template <typename T> struct OpLe { static __device__ int8_t apply(T a, T b) { return a <= b; } };
template <typename T> struct OpEq { static __device__ int8_t apply(T a, T b) { return a == b; } };
template <typename T> struct OpNe { static __device__ int8_t apply(T a, T b) { return a != b; } };
template <typename T> struct OpAnd { static __device__ T apply(T a, T b) { return a & b; } };
template <typename T> struct OpOr { static __device__ T apply(T a, T b) { return a | b; } };
template <typename T> struct OpXor { static __device__ T apply(T a, T b) { return a ^ b; } };

template <typename T, int N> struct alignas(sizeof(T) * N) Pack { T v[N]; };

// Unmasked, 16-byte aligned path.  VEC input elements per 128-bit load.
template <typename T, typename Tout, typename Op>
__global__ void __launch_bounds__(kThreads) binary_vec_kernel(const T* __restrict__ lhs,
                                                              const T* __restrict__ rhs,
                                                              Tout* __restrict__ out, size_t n) {
  constexpr int VEC = 16 / sizeof(T);
  using InPack = Pack<T, VEC>;
  using OutPack = Pack<Tout, VEC>;
  const size_t nvec = n / VEC;
  const size_t stride = (size_t)gridDim.x * kThreads;
  size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x;
  const uint4* l4 = reinterpret_cast<const uint4*>(lhs);
  const uint4* r4 = reinterpret_cast<const uint4*>(rhs);
  OutPack* o = reinterpret_cast<OutPack*>(out);
  for (; i + (kUnroll - 1) * stride < nvec; i += kUnroll * stride) {
    uint4 a[kUnroll], b[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) a[u] = ldg_stream(l4 + i + u * stride);
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) b[u] = ldg_stream(r4 + i + u * stride);
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const InPack& pa = reinterpret_cast<const InPack&>(a[u]);
      const InPack& pb = reinterpret_cast<const InPack&>(b[u]);
      OutPack po;
#pragma unroll
      for (int e = 0; e < VEC; ++e) po.v[e] = Op::apply(pa.v[e], pb.v[e]);
      o[i + u * stride] = po;
    }
  }
  for (; i < nvec; i += stride) {
    uint4 a = ldg_stream(l4 + i), b = ldg_stream(r4 + i);
    const InPack& pa = reinterpret_cast<const InPack&>(a);
    const InPack& pb = reinterpret_cast<const InPack&>(b);
    OutPack po;
#pragma unroll
    for (int e = 0; e < VEC; ++e) po.v[e] = Op::apply(pa.v[e], pb.v[e]);
    o[i] = po;
  }
  // scalar tail (< VEC elements), one thread
  if (blockIdx.x == 0 && threadIdx.x == 0)
    for (size_t k = nvec * VEC; k < n; ++k) out[k] = Op::apply(lhs[k], rhs[k]);
}

// Masked and/or unaligned path: one thread per 8 rows = one validity byte per side, read in-scan.
template <typename T, typename Tout, typename Op>
__global__ void __launch_bounds__(kThreads) binary_masked_kernel(const T* __restrict__ lhs,
                                                                 const gdf_valid_type* __restrict__ lv,
                                                                 const T* __restrict__ rhs,
                                                                 const gdf_valid_type* __restrict__ rv,
                                                                 Tout* __restrict__ out, size_t n) {
  const size_t nbytes = (n + 7) / 8;
  const size_t stride = (size_t)gridDim.x * kThreads;
  for (size_t g = (size_t)blockIdx.x * kThreads + threadIdx.x; g < nbytes; g += stride) {
    unsigned m = 0xffu;
    if (lv) m &= lv[g];
    if (rv) m &= rv[g];
    const size_t base = g * 8;
    const int cnt = (n - base) < 8 ? (int)(n - base) : 8;
#pragma unroll
    for (int e = 0; e < 8; ++e)
      if (e < cnt && ((m >> e) & 1u)) out[base + e] = Op::apply(lhs[base + e], rhs[base + e]);
  }
}

template <typename T, typename Tout, typename Op>
gdf_error launch_binary(gdf_column* lhs, gdf_column* rhs, gdf_column* output) {
  if (lhs->size == 0 || rhs->size == 0) return GDF_SUCCESS;
  B200_REQUIRE(lhs->size == rhs->size, GDF_COLUMN_SIZE_MISMATCH);
  B200_REQUIRE(lhs->size == output->size, GDF_COLUMN_SIZE_MISMATCH);
  B200_REQUIRE(lhs->dtype == rhs->dtype, GDF_UNSUPPORTED_DTYPE);
  const size_t n = lhs->size;
  const T* l = static_cast<const T*>(lhs->data);
  const T* r = static_cast<const T*>(rhs->data);
  Tout* o = static_cast<Tout*>(output->data);
  gdf_nvtx_range_push("LIBGDF_BINARY_OP", GDF_YELLOW);
  const int max_blocks = sm_count() * 8;
  B200_TIMED("binary_op");
  constexpr int VEC = 16 / sizeof(T);
  const bool out_aligned = (reinterpret_cast<uintptr_t>(o) % (sizeof(Tout) * VEC)) == 0;
  if (!lhs->valid && !rhs->valid && aligned16(l) && aligned16(r) && out_aligned) {
    size_t work = (n / VEC + (size_t)kThreads * kUnroll - 1) / ((size_t)kThreads * kUnroll);
    int blocks = (int)(work < (size_t)max_blocks ? (work ? work : 1) : (size_t)max_blocks);
    binary_vec_kernel<T, Tout, Op><<<blocks, kThreads>>>(l, r, o, n);
  } else {
    size_t work = ((n + 7) / 8 + kThreads - 1) / kThreads;
    int blocks = (int)(work < (size_t)max_blocks ? work : (size_t)max_blocks);
    binary_masked_kernel<T, Tout, Op><<<blocks, kThreads>>>(l, lhs->valid, r, rhs->valid, o, n);
  }
  gdf_nvtx_range_pop();
  B200_CHECK_LAST();
  return GDF_SUCCESS;
}

template <typename T, template <typename> class Op>
gdf_error arith(gdf_column* l, gdf_column* r, gdf_column* o) {
  B200_REQUIRE(o->dtype == l->dtype, GDF_UNSUPPORTED_DTYPE);
  return launch_binary<T, T, Op<T>>(l, r, o);
}
template <typename T, template <typename> class Op>
gdf_error logical(gdf_column* l, gdf_column* r, gdf_column* o) {
  B200_REQUIRE(o->dtype == GDF_INT8, GDF_UNSUPPORTED_DTYPE);
  return launch_binary<T, int8_t, Op<T>>(l, r, o);
}

}  // namespace
}  // namespace b200

using namespace b200;

#define B200_ARITH(NAME, OP)                                                                           \
  extern "C" gdf_error gdf_##NAME##_i32(gdf_column* l, gdf_column* r, gdf_column* o) { return arith<int32_t, OP>(l, r, o); } \
  extern "C" gdf_error gdf_##NAME##_i64(gdf_column* l, gdf_column* r, gdf_column* o) { return arith<int64_t, OP>(l, r, o); } \
  extern "C" gdf_error gdf_##NAME##_f32(gdf_column* l, gdf_column* r, gdf_column* o) { return arith<float, OP>(l, r, o); }   \
  extern "C" gdf_error gdf_##NAME##_f64(gdf_column* l, gdf_column* r, gdf_column* o) { return arith<double, OP>(l, r, o); }  \
  extern "C" gdf_error gdf_##NAME##_generic(gdf_column* l, gdf_column* r, gdf_column* o) {             \
    switch (l->dtype) { /* ref binaryops.cu:107-116 */                                                 \
      case GDF_INT32: return gdf_##NAME##_i32(l, r, o);                                                \
      case GDF_INT64: return gdf_##NAME##_i64(l, r, o);                                                \
      case GDF_FLOAT32: return gdf_##NAME##_f32(l, r, o);                                              \
      case GDF_FLOAT64: return gdf_##NAME##_f64(l, r, o);                                              \
      default: return GDF_UNSUPPORTED_DTYPE;                                                           \
    }                                                                                                  \
  }

B200_ARITH(add, OpAdd)
B200_ARITH(sub, OpSub)
B200_ARITH(mul, OpMul)
B200_ARITH(floordiv, OpFloorDiv)

extern "C" gdf_error gdf_div_f32(gdf_column* l, gdf_column* r, gdf_column* o) { return arith<float, OpDiv>(l, r, o); }
extern "C" gdf_error gdf_div_f64(gdf_column* l, gdf_column* r, gdf_column* o) { return arith<double, OpDiv>(l, r, o); }
extern "C" gdf_error gdf_div_generic(gdf_column* l, gdf_column* r, gdf_column* o) {
  switch (l->dtype) {  // ref binaryops.cu:97-105
    case GDF_FLOAT32: return gdf_div_f32(l, r, o);
    case GDF_FLOAT64: return gdf_div_f64(l, r, o);
    default: return GDF_UNSUPPORTED_DTYPE;
  }
}

#define B200_LOGICAL(NAME, OP)                                                                          \
  extern "C" gdf_error gdf_##NAME##_i8(gdf_column* l, gdf_column* r, gdf_column* o) { return logical<int8_t, OP>(l, r, o); }   \
  extern "C" gdf_error gdf_##NAME##_i32(gdf_column* l, gdf_column* r, gdf_column* o) { return logical<int32_t, OP>(l, r, o); } \
  extern "C" gdf_error gdf_##NAME##_i64(gdf_column* l, gdf_column* r, gdf_column* o) { return logical<int64_t, OP>(l, r, o); } \
  extern "C" gdf_error gdf_##NAME##_f32(gdf_column* l, gdf_column* r, gdf_column* o) { return logical<float, OP>(l, r, o); }   \
  extern "C" gdf_error gdf_##NAME##_f64(gdf_column* l, gdf_column* r, gdf_column* o) { return logical<double, OP>(l, r, o); }  \
  extern "C" gdf_error gdf_##NAME##_generic(gdf_column* l, gdf_column* r, gdf_column* o) {              \
    switch (l->dtype) { /* ref binaryops.cu:254-267 */                                                  \
      case GDF_INT8: return gdf_##NAME##_i8(l, r, o);                                                   \
      case GDF_INT32: case GDF_DATE32: return gdf_##NAME##_i32(l, r, o);                                \
      case GDF_INT64: case GDF_DATE64: case GDF_TIMESTAMP: return gdf_##NAME##_i64(l, r, o);            \
      case GDF_FLOAT32: return gdf_##NAME##_f32(l, r, o);                                               \
      case GDF_FLOAT64: return gdf_##NAME##_f64(l, r, o);                                               \
      default: return GDF_UNSUPPORTED_DTYPE;                                                            \
    }                                                                                                   \
  }

B200_LOGICAL(gt, OpGt)
B200_LOGICAL(ge, OpGe)
B200_LOGICAL(lt, OpLt)
B200_LOGICAL(le, OpLe)
B200_LOGICAL(eq, OpEq)
B200_LOGICAL(ne, OpNe)

#define B200_BITWISE(NAME, OP)                                                                          \
  extern "C" gdf_error gdf_bitwise_##NAME##_i8(gdf_column* l, gdf_column* r, gdf_column* o) { return arith<int8_t, OP>(l, r, o); }   \
  extern "C" gdf_error gdf_bitwise_##NAME##_i32(gdf_column* l, gdf_column* r, gdf_column* o) { return arith<int32_t, OP>(l, r, o); } \
  extern "C" gdf_error gdf_bitwise_##NAME##_i64(gdf_column* l, gdf_column* r, gdf_column* o) { return arith<int64_t, OP>(l, r, o); } \
  extern "C" gdf_error gdf_bitwise_##NAME##_generic(gdf_column* l, gdf_column* r, gdf_column* o) {      \
    switch (l->dtype) { /* ref binaryops.cu:445-453 */                                                  \
      case GDF_INT8: return gdf_bitwise_##NAME##_i8(l, r, o);                                           \
      case GDF_INT32: return gdf_bitwise_##NAME##_i32(l, r, o);                                         \
      case GDF_INT64: return gdf_bitwise_##NAME##_i64(l, r, o);                                         \
      default: return GDF_UNSUPPORTED_DTYPE;                                                            \
    }                                                                                                   \
  }

B200_BITWISE(and, OpAnd)
B200_BITWISE(or, OpOr)
B200_BITWISE(xor, OpXor)

// output->valid = lhs->valid & rhs->valid over ceil(size/8) bytes (ref binaryops.cu:509-526).
extern "C" gdf_error gdf_validity_and(gdf_column* lhs, gdf_column* rhs, gdf_column* output) {
  B200_REQUIRE(lhs->valid && rhs->valid && output->valid, GDF_VALIDITY_MISSING);
  auto as_bytes = [](const gdf_column& c) {
    gdf_column v;
    v.data = c.valid;
    v.valid = nullptr;
    v.size = valid_bytes(c.size);
    v.dtype = GDF_INT8;
    v.null_count = 0;
    return v;
  };
  gdf_column x = as_bytes(*lhs), y = as_bytes(*rhs), z = as_bytes(*output);
  gdf_bitwise_and_i8(&x, &y, &z);  // return value deliberately ignored, as in the reference (:524)
  return GDF_SUCCESS;
}
