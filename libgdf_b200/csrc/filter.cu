// Predicate filter + stream compaction:
//   gdf_filter                      multi-column "row == value tuple" -> ascending row indices
//   gpu_comparison_static_{i8..f64} column OP scalar  -> int8 stencil
//   gpu_comparison                  column OP column  -> int8 stencil
//   gpu_apply_stencil               compaction of a column by (stencil byte != 0 && stencil bit)
//
// Reference behaviour followed (file:line in /root/reference/libgdf/src):
//   gdf_filter          sqls_ops.cu:1401-1424, sqls_rtti_comp.hpp:200-213,343-370
//                       - rejects a validity mask on cols[0] (GDF_VALIDITY_UNSUPPORTED)
//                       - fills the caller's d_cols / d_types device scratch (soa_col_info, :27-41)
//                       - a row is kept when NO column differs (`!=`, so NaN rows are dropped)
//                       - indices are size_t and ascending (copy_if is stable)
//   comparison (static) filterops.cu:162-255   size mismatch and non-int8 output both report
//                       GDF_COLUMN_SIZE_MISMATCH (:163-165); mixed lhs/scalar types compare under
//                       the usual C++ arithmetic conversions; output->valid = all ones when
//                       null_count == 0 else a copy of lhs->valid, output->null_count follows (:139-148)
//   comparison (col)    filterops.cu:260-662   same, output->valid = lhs->valid & rhs->valid
//   apply_stencil       streamcompactionops.cu:208-339   see DESIGN.md "quirks" for the two
//                       documented divergences (LESS_THAN operators, rebuilt output mask).
#include "select.cuh"
#include "select_chunked.cuh"
#include "select_stream.cuh"

namespace b200 {
namespace {

// ------------------------------------------------------------------------------------------
// gdf_filter policies
// ------------------------------------------------------------------------------------------
template <typename T>
struct FilterOne {  // one column of T: the C2 benchmark shape
  static constexpr int V = 16 / sizeof(T);
  static constexpr int K = sizeof(T) == 8 ? 8 : 32 / V;
  const T* data;
  const void* const* d_vals;
  size_t* out;
  bool vec_ok;

  __device__ uint32_t flags(size_t warp_base, size_t n) const {
    const T target = *static_cast<const T*>(d_vals[0]);
    const unsigned lane = lane_id();
    uint32_t f = 0;
    if (vec_ok && warp_base + (size_t)32 * V * K <= n) {
      uint4 raw[K];
#pragma unroll
      for (int k = 0; k < K; ++k) raw[k] = ldg_stream(data + warp_base + (size_t)k * 32 * V + (size_t)lane * V);
#pragma unroll
      for (int k = 0; k < K; ++k) {
        const T* e = reinterpret_cast<const T*>(&raw[k]);
#pragma unroll
        for (int j = 0; j < V; ++j) f |= (uint32_t)(!(e[j] != target)) << (k * V + j);
      }
    } else {
#pragma unroll
      for (int k = 0; k < K; ++k)
#pragma unroll
        for (int j = 0; j < V; ++j) {
          const size_t row = warp_base + (size_t)k * 32 * V + (size_t)lane * V + j;
          if (row < n) f |= (uint32_t)(!(data[row] != target)) << (k * V + j);
        }
    }
    return f;
  }
  __device__ void emit(size_t row, size_t pos) const { out[pos] = row; }
};

template <typename T>
static __device__ __forceinline__ bool differs(const void* col, size_t row, const void* val) {
  return static_cast<const T*>(col)[row] != *static_cast<const T*>(val);
}

struct FilterMany {  // any number of columns, runtime dtypes, read from the caller's d_cols/d_types
  static constexpr int V = 1;
  static constexpr int K = 16;
  const void* const* d_cols;
  const int* d_types;
  const void* const* d_vals;
  int ncols;
  size_t* out;

  __device__ uint32_t flags(size_t warp_base, size_t n) const {
    const unsigned lane = lane_id();
    uint32_t f = 0;
#pragma unroll 4
    for (int k = 0; k < K; ++k) {
      const size_t row = warp_base + (size_t)k * 32 + lane;
      if (row >= n) continue;
      bool keep = true;
      for (int c = 0; c < ncols && keep; ++c) {
        const void* col = d_cols[c];
        const void* val = d_vals[c];
        switch (d_types[c]) {
          case GDF_INT8: keep = !differs<int8_t>(col, row, val); break;
          case GDF_INT16: keep = !differs<int16_t>(col, row, val); break;
          case GDF_INT32: keep = !differs<int32_t>(col, row, val); break;
          case GDF_INT64: keep = !differs<int64_t>(col, row, val); break;
          case GDF_FLOAT32: keep = !differs<float>(col, row, val); break;
          case GDF_FLOAT64: keep = !differs<double>(col, row, val); break;
          default: break;
        }
      }
      f |= (uint32_t)keep << k;
    }
    return f;
  }
  __device__ void emit(size_t row, size_t pos) const { out[pos] = row; }
};

// ------------------------------------------------------------------------------------------
// gpu_apply_stencil policy: stencil is an int8 column with a validity mask that the reference reads
// MSB-first inside each byte (bit 7-(i%8), streamcompactionops.cu:89-107 - its n_bytes member is
// never initialised, so the "last byte" branch is dead in practice).
// ------------------------------------------------------------------------------------------
template <typename T>
struct StencilPolicy {
  static constexpr int V = 16;
  static constexpr int K = 2;
  const int8_t* stencil;
  const gdf_valid_type* svalid;
  const T* data;
  T* out;
  bool vec_ok;        // the stencil bytes can be read with 128-bit loads
  bool data_vec_ok;   // so can the data column

  __device__ uint32_t flags(size_t warp_base, size_t n) const {
    const unsigned lane = lane_id();
    uint32_t f = 0;
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const size_t row0 = warp_base + (size_t)k * 32 * V + (size_t)lane * V;
      if (vec_ok && row0 + V <= n) {
        const uint4 raw = ldg_stream(stencil + row0);
        const int8_t* s = reinterpret_cast<const int8_t*>(&raw);
        const unsigned m = (unsigned)svalid[row0 >> 3] | ((unsigned)svalid[(row0 >> 3) + 1] << 8);
#pragma unroll
        for (int j = 0; j < V; ++j) {
          const unsigned bit = (m >> ((j & 8) + 7 - (j & 7))) & 1u;
          f |= (uint32_t)(bit && s[j] != 0) << (k * V + j);
        }
      } else {
#pragma unroll
        for (int j = 0; j < V; ++j) {
          const size_t row = row0 + j;
          if (row < n) {
            const unsigned bit = ((unsigned)svalid[row >> 3] >> (7 - (row & 7))) & 1u;
            f |= (uint32_t)(bit && stencil[row] != 0) << (k * V + j);
          }
        }
      }
    }
    return f;
  }
  __device__ void emit(size_t row, size_t pos) const { out[pos] = data[row]; }
  // A step is 32 lanes x V = 512 consecutive rows.  At C2's 10 % selectivity a per-row gather touches 81 % of the
  // column's 128-byte lines anyway (ncu: 7.7 GB of DRAM reads for 0.8 GB of selected values) and does so with one
  // divergent request per row; so when a step has >= kDenseMin selected rows the warp streams the step's rows with
  // coalesced 128-bit loads instead - instruction i covers the 32 consecutive 16-byte pieces i * 32 + lane - and every
  // lane fetches the flag bits and output position of the piece's OWNER lane with two shuffles.
  static constexpr uint32_t kDenseMin = 16;
  __device__ bool emit_dense(size_t step_row0, size_t n, uint32_t my_bits, size_t my_pos, uint32_t step_total) const {
    constexpr int E = 16 / (int)sizeof(T);   // rows per 16-byte piece
    constexpr int I = V / E;                 // pieces per lane = load instructions per step
    if (!data_vec_ok || step_total < kDenseMin || step_row0 + (size_t)32 * V > n) return false;  // warp-uniform
    const unsigned lane = lane_id();
    const uint4* src = reinterpret_cast<const uint4*>(data + step_row0);
    constexpr int B = I < 4 ? I : 4;         // loads in flight per lane (8 x 16 bytes spilled at 3 CTAs per SM)
#pragma unroll
    for (int i0 = 0; i0 < I; i0 += B) {
      uint4 raw[B];
#pragma unroll
      for (int i = 0; i < B; ++i) raw[i] = ldg_stream(src + (i0 + i) * 32 + lane);
#pragma unroll
      for (int i = 0; i < B; ++i) {
        const unsigned piece = (unsigned)(i0 + i) * 32u + lane;
        const unsigned owner = piece / I, sub = piece % I;
        const uint32_t obits = __shfl_sync(0xffffffffu, my_bits, owner);
        const unsigned long long opos = __shfl_sync(0xffffffffu, (unsigned long long)my_pos, owner);
        const uint32_t mine = (obits >> (sub * E)) & ((1u << E) - 1u);
        size_t p = (size_t)opos + __popc(obits & ((1u << (sub * E)) - 1u));
        const T* e = reinterpret_cast<const T*>(&raw[i]);
#pragma unroll
        for (int j = 0; j < E; ++j)
          if ((mine >> j) & 1u) out[p++] = e[j];
      }
    }
    return true;
  }
  // few selected rows: all gathers of the lane's step in flight before the first store (select_chunked.cuh: emit_step)
  __device__ void emit_step(size_t row0, uint32_t bits, size_t pos) const {
    T v[V];
#pragma unroll
    for (int j = 0; j < V; ++j)
      if ((bits >> j) & 1u) v[j] = data[row0 + j];
#pragma unroll
    for (int j = 0; j < V; ++j)
      if ((bits >> j) & 1u) out[pos++] = v[j];
  }
};

// gdf_filter, one 16-byte aligned column: streaming kernel (select_stream.cuh)
template <typename T>
struct EqualsDeviceScalar {  // row kept when !(value != *d_vals[0])  (ref sqls_rtti_comp.hpp:200-213)
  const void* const* d_vals;
  T target;
  __device__ void prepare() { target = *static_cast<const T*>(d_vals[0]); }
  __device__ bool operator()(T v) const { return !(v != target); }
};
struct EmitRowIndex {
  size_t* out;
  __device__ void operator()(size_t row, size_t pos) const { out[pos] = row; }
};

gdf_error read_count(const unsigned long long* d_count, size_t* h_count) {
  unsigned long long* box = static_cast<unsigned long long*>(pinned_mailbox());
  B200_REQUIRE(box != nullptr, GDF_CUDA_ERROR);
  B200_CUDA_TRY(cudaMemcpyAsync(box, d_count, sizeof(unsigned long long), cudaMemcpyDeviceToHost, 0));
  B200_CUDA_TRY(cudaStreamSynchronize(0));
  *h_count = (size_t)*box;
  return GDF_SUCCESS;
}

template <typename T>
gdf_error run_filter_stream(const T* data, size_t n, const void* const* d_vals, size_t* out, size_t* h_count) {
  using G = select_stream::Geom<T>;
  const size_t chunks = (n + G::kChunkRows - 1) / G::kChunkRows;
  B200_REQUIRE(chunks < (1ull << 31), GDF_COLUMN_SIZE_TOO_BIG);
  Scratch desc;  // [chunks] look-back descriptors | selected count | chunk ticket
  const size_t bytes = chunks * sizeof(uint64_t) + 2 * sizeof(unsigned long long);
  B200_CUDA_TRY(desc.alloc(bytes));
  B200_CUDA_TRY(cudaMemsetAsync(desc.ptr, 0, bytes, 0));
  uint64_t* d = desc.as<uint64_t>();
  unsigned long long* d_count = reinterpret_cast<unsigned long long*>(d + chunks);
  unsigned* d_ticket = reinterpret_cast<unsigned*>(d_count + 1);
  // Two dealings of the same kernel (select_stream.cuh).  STATIC (CTA b: chunks b, b + grid, ...) is the faster one
  // - 1.54 ms against 1.83-1.89 ms at C2 for every placement of the ticket atomic that was tried (profiles/
  // r02_notes.md) - but its look-back only makes progress if the whole grid is co-resident.  A COOPERATIVE launch
  // makes that a guarantee of the driver instead of an assumption (the grid is not started until every CTA fits, even
  // when another stream's persistent kernel holds SMs).  When the device refuses a cooperative launch, the TICKET
  // dealing runs: chunks are handed out in increasing order to CTAs that are already running, so it needs no
  // co-residency at all.
  auto kern_static = select_stream::select_stream_kernel<T, EqualsDeviceScalar<T>, EmitRowIndex, true>;
  auto kern_ticket = select_stream::select_stream_kernel<T, EqualsDeviceScalar<T>, EmitRowIndex, false>;
  const int smem = (int)select_stream::smem_bytes<T>();
  B200_CUDA_TRY(cudaFuncSetAttribute(kern_static, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  B200_CUDA_TRY(cudaFuncSetAttribute(kern_ticket, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  int per_sm = 0;  // persistent CTAs: one resident wave
  B200_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern_static, select_stream::kThreads, smem));
  B200_REQUIRE(per_sm >= 1, GDF_CUDA_ERROR);
  const size_t resident = (size_t)sm_count() * (size_t)per_sm;
  const unsigned blocks = (unsigned)(chunks < resident ? chunks : resident);
  EqualsDeviceScalar<T> pred{d_vals, T()};
  EmitRowIndex emit{out};
  bool launched = false;
  {
    B200_TIMED("select");
    if (cooperative_launch_ok() && select_dealing_mode() == 0) {
      void* args[] = {(void*)&data, (void*)&n, (void*)&pred, (void*)&emit, (void*)&d, (void*)&d_count, (void*)&d_ticket};
      const cudaError_t ce = cudaLaunchCooperativeKernel((const void*)kern_static, dim3(blocks), dim3(select_stream::kThreads), args,
                                                         (size_t)smem, 0);
      if (ce == cudaSuccess) launched = true;
      else (void)cudaGetLastError();  // refused (e.g. MPS / partitioned device): the ticket dealing needs no guarantee
    }
    if (!launched) kern_ticket<<<blocks, select_stream::kThreads, smem>>>(data, n, pred, emit, d, d_count, d_ticket);
  }
  B200_CHECK_LAST();
  return read_count(d_count, h_count);
}

// Launch one select pass (select_chunked.cuh); returns the number of selected rows through *h_count (host).
template <typename Policy>
gdf_error run_select(const Policy& pol, size_t n, size_t* h_count) {
  if (n == 0) {
    *h_count = 0;
    return GDF_SUCCESS;
  }
  using G = select_chunked::Geom<Policy>;
  const size_t chunks = (n + G::kChunkRows - 1) / G::kChunkRows;
  B200_REQUIRE(chunks < (1ull << 31), GDF_COLUMN_SIZE_TOO_BIG);
  Scratch desc;  // [chunks] look-back descriptors | selected count | chunk ticket
  const size_t bytes = chunks * sizeof(uint64_t) + 2 * sizeof(unsigned long long);
  B200_CUDA_TRY(desc.alloc(bytes));
  B200_CUDA_TRY(cudaMemsetAsync(desc.ptr, 0, bytes, 0));
  uint64_t* d = desc.as<uint64_t>();
  unsigned long long* d_count = reinterpret_cast<unsigned long long*>(d + chunks);
  unsigned* d_ticket = reinterpret_cast<unsigned*>(d_count + 1);
  auto kern = select_chunked::select_chunked_kernel<Policy>;
  int per_sm = 0;
  B200_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, select_chunked::kThreads, 0));
  const size_t resident = (size_t)sm_count() * (size_t)(per_sm > 0 ? per_sm : 1);
  const unsigned blocks = (unsigned)(chunks < resident ? chunks : resident);
  {
    B200_TIMED("select");
    kern<<<blocks, select_chunked::kThreads>>>(pol, n, d, d_count, d_ticket);
  }
  B200_CHECK_LAST();
  return read_count(d_count, h_count);
}

// ------------------------------------------------------------------------------------------
// comparison -> int8 stencil
// ------------------------------------------------------------------------------------------
template <typename A, typename B>
static __device__ __forceinline__ int8_t compare(A a, B b, int op) {
  switch (op) {
    case GDF_EQUALS: return a == b;
    case GDF_NOT_EQUALS: return a != b;
    case GDF_LESS_THAN: return a < b;
    case GDF_LESS_THAN_OR_EQUALS: return a <= b;
    case GDF_GREATER_THAN: return a > b;
    default: return a >= b;  // GDF_GREATER_THAN_OR_EQUALS
  }
}

template <typename T, int N> struct alignas(sizeof(T) * N) Pack { T v[N]; };

// column OP scalar, 128-bit loads on the column, V result bytes stored per step.
template <typename A, typename B>
__global__ void __launch_bounds__(256) compare_static_kernel(const A* __restrict__ lhs, B value,
                                                             int8_t* __restrict__ out, size_t n, int op,
                                                             bool vec_ok) {
  constexpr int V = 16 / sizeof(A);
  constexpr int U = 4;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const size_t gtid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (vec_ok) {
    const size_t nvec = n / V;
    const uint4* l4 = reinterpret_cast<const uint4*>(lhs);
    Pack<int8_t, V>* o = reinterpret_cast<Pack<int8_t, V>*>(out);
    size_t i = gtid;
    for (; i + (U - 1) * stride < nvec; i += U * stride) {
      uint4 raw[U];
#pragma unroll
      for (int u = 0; u < U; ++u) raw[u] = ldg_stream(l4 + i + u * stride);
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const A* e = reinterpret_cast<const A*>(&raw[u]);
        Pack<int8_t, V> r;
#pragma unroll
        for (int j = 0; j < V; ++j) r.v[j] = compare(e[j], value, op);
        o[i + u * stride] = r;
      }
    }
    for (; i < nvec; i += stride) {
      const uint4 raw = ldg_stream(l4 + i);
      const A* e = reinterpret_cast<const A*>(&raw);
      Pack<int8_t, V> r;
#pragma unroll
      for (int j = 0; j < V; ++j) r.v[j] = compare(e[j], value, op);
      o[i] = r;
    }
    if (gtid == 0)
      for (size_t k = nvec * V; k < n; ++k) out[k] = compare(lhs[k], value, op);
  } else {
    for (size_t i = gtid; i < n; i += stride) out[i] = compare(lhs[i], value, op);
  }
}

template <typename A, typename B>
__global__ void __launch_bounds__(256) compare_columns_kernel(const A* __restrict__ lhs,
                                                              const B* __restrict__ rhs,
                                                              int8_t* __restrict__ out, size_t n, int op) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    out[i] = compare(lhs[i], rhs[i], op);
}

// out = a & b over nbytes, counting zero bits among the first `rows` bits.
__global__ void and_masks_kernel(const gdf_valid_type* __restrict__ a, const gdf_valid_type* __restrict__ b,
                                 gdf_valid_type* __restrict__ out, size_t rows,
                                 unsigned long long* __restrict__ nulls) {
  const size_t nbytes = (rows + 7) / 8;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  unsigned local = 0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nbytes; i += stride) {
    unsigned v = a[i] & b[i];
    out[i] = (gdf_valid_type)v;
    const unsigned live = (i == nbytes - 1 && (rows & 7)) ? ((1u << (rows & 7)) - 1u) : 0xffu;
    local += __popc((~v) & live);
  }
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) local += __shfl_xor_sync(0xffffffffu, local, s);
  if (lane_id() == 0 && local) atomicAdd(nulls, (unsigned long long)local);
}

int stream_blocks(size_t items_per_thread_total) {
  const size_t want = (items_per_thread_total + 255) / 256;
  const size_t cap = (size_t)sm_count() * 8;
  return (int)(want < 1 ? 1 : (want < cap ? want : cap));
}

// Output validity of a comparison (ref filterops.cu:139-153).
gdf_error comparison_validity(gdf_column* lhs, gdf_column* rhs /*nullable*/, gdf_column* out) {
  const size_t nbytes = valid_bytes(out->size);
  const size_t lnull = lhs->null_count, rnull = rhs ? rhs->null_count : lhs->null_count;
  const gdf_valid_type* lv = lhs->valid;
  const gdf_valid_type* rv = rhs ? rhs->valid : lhs->valid;
  if (!out->valid) return GDF_SUCCESS;  // nothing to fill (the reference would fault here)
  if (lnull == 0 && rnull == 0) {
    B200_CUDA_TRY(cudaMemsetAsync(out->valid, 0xff, nbytes, 0));
    out->null_count = 0;
  } else if (lv == rv) {
    if (lv) B200_CUDA_TRY(cudaMemcpyAsync(out->valid, lv, nbytes, cudaMemcpyDeviceToDevice, 0));
    out->null_count = lnull;
  } else if (lv && rv) {
    Scratch cnt;
    B200_CUDA_TRY(cnt.alloc(sizeof(unsigned long long)));
    B200_CUDA_TRY(cudaMemsetAsync(cnt.ptr, 0, sizeof(unsigned long long), 0));
    and_masks_kernel<<<stream_blocks(nbytes), 256>>>(lv, rv, out->valid, out->size, cnt.as<unsigned long long>());
    B200_CHECK_LAST();
    unsigned long long* box = static_cast<unsigned long long*>(pinned_mailbox());
    B200_REQUIRE(box != nullptr, GDF_CUDA_ERROR);
    B200_CUDA_TRY(cudaMemcpyAsync(box, cnt.ptr, sizeof(unsigned long long), cudaMemcpyDeviceToHost, 0));
    B200_CUDA_TRY(cudaStreamSynchronize(0));
    out->null_count = (gdf_size_type)*box;
  } else {  // exactly one side carries a mask: the other side is all-valid
    const gdf_valid_type* only = lv ? lv : rv;
    B200_CUDA_TRY(cudaMemcpyAsync(out->valid, only, nbytes, cudaMemcpyDeviceToDevice, 0));
    out->null_count = lv ? lnull : rnull;
  }
  return GDF_SUCCESS;
}

template <typename A, typename B>
gdf_error launch_static(gdf_column* lhs, B value, gdf_column* out, int op) {
  const size_t n = lhs->size;
  if (n) {
    const A* l = static_cast<const A*>(lhs->data);
    int8_t* o = static_cast<int8_t*>(out->data);
    constexpr int V = 16 / sizeof(A);
    const bool vec_ok = aligned16(l) && (reinterpret_cast<uintptr_t>(o) % V == 0);
    B200_TIMED("compare_static");
    compare_static_kernel<A, B><<<stream_blocks(n / V / 4 + 1), 256>>>(l, value, o, n, op, vec_ok);
    B200_CHECK_LAST();
  }
  return comparison_validity(lhs, nullptr, out);
}

template <typename B>
gdf_error comparison_static(gdf_column* lhs, B value, gdf_column* out, gdf_comparison_operator op) {
  B200_REQUIRE(lhs->size == out->size, GDF_COLUMN_SIZE_MISMATCH);
  B200_REQUIRE(out->dtype == GDF_INT8, GDF_COLUMN_SIZE_MISMATCH);  // sic, ref filterops.cu:165
  switch (lhs->dtype) {
    case GDF_INT8: return launch_static<int8_t, B>(lhs, value, out, op);
    case GDF_INT16: return launch_static<int16_t, B>(lhs, value, out, op);
    case GDF_INT32: return launch_static<int32_t, B>(lhs, value, out, op);
    case GDF_INT64: return launch_static<int64_t, B>(lhs, value, out, op);
    case GDF_FLOAT32: return launch_static<float, B>(lhs, value, out, op);
    case GDF_FLOAT64: return launch_static<double, B>(lhs, value, out, op);
    default: return GDF_UNSUPPORTED_DTYPE;
  }
}

template <typename A, typename B>
gdf_error launch_columns(gdf_column* lhs, gdf_column* rhs, gdf_column* out, int op) {
  const size_t n = lhs->size;
  if (n) {
    compare_columns_kernel<A, B><<<stream_blocks(n / 4 + 1), 256>>>(
        static_cast<const A*>(lhs->data), static_cast<const B*>(rhs->data), static_cast<int8_t*>(out->data), n, op);
    B200_CHECK_LAST();
  }
  return comparison_validity(lhs, rhs, out);
}

template <typename A>
gdf_error columns_rhs(gdf_column* lhs, gdf_column* rhs, gdf_column* out, int op) {
  switch (rhs->dtype) {
    case GDF_INT8: return launch_columns<A, int8_t>(lhs, rhs, out, op);
    case GDF_INT16: return launch_columns<A, int16_t>(lhs, rhs, out, op);
    case GDF_INT32: return launch_columns<A, int32_t>(lhs, rhs, out, op);
    case GDF_INT64: return launch_columns<A, int64_t>(lhs, rhs, out, op);
    case GDF_FLOAT32: return launch_columns<A, float>(lhs, rhs, out, op);
    case GDF_FLOAT64: return launch_columns<A, double>(lhs, rhs, out, op);
    default: return GDF_UNSUPPORTED_DTYPE;
  }
}

template <typename T>
gdf_error stencil_typed(gdf_column* lhs, gdf_column* stencil, gdf_column* out, size_t* count) {
  StencilPolicy<T> pol;
  pol.stencil = static_cast<const int8_t*>(stencil->data);
  pol.svalid = stencil->valid;
  pol.data = static_cast<const T*>(lhs->data);
  pol.out = static_cast<T*>(out->data);
  pol.vec_ok = aligned16(stencil->data);
  pol.data_vec_ok = aligned16(lhs->data);
  return run_select(pol, (size_t)lhs->size, count);
}

__global__ void fill_compacted_mask_kernel(gdf_valid_type* out, size_t nbytes, unsigned last_byte) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nbytes; i += stride)
    out[i] = (i == nbytes - 1) ? (gdf_valid_type)last_byte : (gdf_valid_type)0xff;
}

}  // namespace
}  // namespace b200

using namespace b200;

extern "C" gdf_error gdf_filter(size_t nrows, gdf_column* cols, size_t ncols, void** d_cols, int* d_types,
                                void** d_vals, size_t* d_indx, size_t* new_sz) {
  B200_REQUIRE(cols != nullptr && new_sz != nullptr, GDF_DATASET_EMPTY);
  B200_REQUIRE(!cols->valid, GDF_VALIDITY_UNSUPPORTED);
  B200_REQUIRE(ncols >= 1, GDF_DATASET_EMPTY);
  for (size_t c = 0; c < ncols; ++c) {
    const int t = cols[c].dtype;
    B200_REQUIRE(t >= GDF_INT8 && t <= GDF_FLOAT64, GDF_UNSUPPORTED_DTYPE);
  }
  // The caller's device scratch is part of the observable contract (ref sqls_ops.cu:27-41).
  {
    void* h_cols[64];
    int h_types[64];
    for (size_t base = 0; base < ncols; base += 64) {
      const size_t m = ncols - base < 64 ? ncols - base : 64;
      for (size_t c = 0; c < m; ++c) {
        h_cols[c] = cols[base + c].data;
        h_types[c] = cols[base + c].dtype;
      }
      B200_CUDA_TRY(cudaMemcpy(d_cols + base, h_cols, m * sizeof(void*), cudaMemcpyHostToDevice));
      B200_CUDA_TRY(cudaMemcpy(d_types + base, h_types, m * sizeof(int), cudaMemcpyHostToDevice));
    }
  }
  if (ncols == 1) {
#define B200_FILTER_ONE(T)                                   \
  {                                                          \
    if (nrows && aligned16(cols[0].data))                    \
      return run_filter_stream<T>(static_cast<const T*>(cols[0].data), nrows, d_vals, d_indx, new_sz); \
    FilterOne<T> pol;                                        \
    pol.data = static_cast<const T*>(cols[0].data);          \
    pol.d_vals = d_vals;                                     \
    pol.out = d_indx;                                        \
    pol.vec_ok = aligned16(cols[0].data);                    \
    return run_select(pol, nrows, new_sz);                   \
  }
    switch (cols[0].dtype) {
      case GDF_INT8: B200_FILTER_ONE(int8_t)
      case GDF_INT16: B200_FILTER_ONE(int16_t)
      case GDF_INT32: B200_FILTER_ONE(int32_t)
      case GDF_INT64: B200_FILTER_ONE(int64_t)
      case GDF_FLOAT32: B200_FILTER_ONE(float)
      case GDF_FLOAT64: B200_FILTER_ONE(double)
      default: return GDF_UNSUPPORTED_DTYPE;
    }
#undef B200_FILTER_ONE
  }
  FilterMany pol;
  pol.d_cols = d_cols;
  pol.d_types = d_types;
  pol.d_vals = d_vals;
  pol.ncols = (int)ncols;
  pol.out = d_indx;
  return run_select(pol, nrows, new_sz);
}

#define B200_CMP_STATIC(SUFFIX, T)                                                                   \
  extern "C" gdf_error gpu_comparison_static_##SUFFIX(gdf_column* lhs, T value, gdf_column* output,  \
                                                      gdf_comparison_operator operation) {           \
    return comparison_static<T>(lhs, value, output, operation);                                      \
  }
B200_CMP_STATIC(i8, int8_t)
B200_CMP_STATIC(i16, int16_t)
B200_CMP_STATIC(i32, int32_t)
B200_CMP_STATIC(i64, int64_t)
B200_CMP_STATIC(f32, float)
B200_CMP_STATIC(f64, double)

extern "C" gdf_error gpu_comparison(gdf_column* lhs, gdf_column* rhs, gdf_column* output,
                                    gdf_comparison_operator operation) {
  B200_REQUIRE(lhs->size == rhs->size, GDF_COLUMN_SIZE_MISMATCH);
  B200_REQUIRE(lhs->size == output->size, GDF_COLUMN_SIZE_MISMATCH);
  B200_REQUIRE(output->dtype == GDF_INT8, GDF_COLUMN_SIZE_MISMATCH);  // sic, ref filterops.cu:264
  switch (lhs->dtype) {
    case GDF_INT8: return columns_rhs<int8_t>(lhs, rhs, output, operation);
    case GDF_INT16: return columns_rhs<int16_t>(lhs, rhs, output, operation);
    case GDF_INT32: return columns_rhs<int32_t>(lhs, rhs, output, operation);
    case GDF_INT64: return columns_rhs<int64_t>(lhs, rhs, output, operation);
    case GDF_FLOAT32: return columns_rhs<float>(lhs, rhs, output, operation);
    case GDF_FLOAT64: return columns_rhs<double>(lhs, rhs, output, operation);
    default: return GDF_UNSUPPORTED_DTYPE;
  }
}

extern "C" gdf_error gpu_apply_stencil(gdf_column* lhs, gdf_column* stencil, gdf_column* output) {
  B200_REQUIRE(output->size == lhs->size, GDF_COLUMN_SIZE_MISMATCH);
  B200_REQUIRE(lhs->dtype == output->dtype, GDF_DTYPE_MISMATCH);
  B200_REQUIRE(!lhs->valid, GDF_VALIDITY_UNSUPPORTED);
  B200_REQUIRE(stencil->valid != nullptr, GDF_VALIDITY_MISSING);  // the reference dereferences it
  B200_REQUIRE(stencil->size == lhs->size, GDF_COLUMN_SIZE_MISMATCH);
  const size_t n = lhs->size;
  size_t count = 0;
  gdf_error err;
  switch (dtype_width(lhs->dtype)) {
    case 1: err = stencil_typed<int8_t>(lhs, stencil, output, &count); break;
    case 2: err = stencil_typed<int16_t>(lhs, stencil, output, &count); break;
    case 4: err = stencil_typed<int32_t>(lhs, stencil, output, &count); break;
    case 8: err = stencil_typed<int64_t>(lhs, stencil, output, &count); break;
    default: return GDF_UNSUPPORTED_DTYPE;
  }
  if (err != GDF_SUCCESS) return err;
  // Rebuilt output mask.  lhs carries no nulls, so every compacted row is valid.  The reference
  // writes ceil(n_in/8) bytes (streamcompactionops.cu:329-331): all ones, except that a ragged last
  // byte is packed MSB-first by bit_mask_pack_op (:185-198) - reproduced here so the bytes match
  // the reference whenever the stencil's own mask is all-valid (see DESIGN.md quirk iii).
  if (output->valid && n) {
    const size_t nbytes = valid_bytes(n);
    const unsigned r = (unsigned)(n & 7);
    const unsigned last = r ? ((0xffu << (8 - r)) & 0xffu) : 0xffu;
    fill_compacted_mask_kernel<<<stream_blocks(nbytes), 256>>>(output->valid, nbytes, last);
    B200_CHECK_LAST();
  }
  output->size = count;
  return GDF_SUCCESS;
}
