// Row ordering: gdf_order_by and the permutation sort behind the sort-based group-by and the hash group-by's
// flag_sort_result.
//
// Reference behaviour followed (file:line in /root/reference/libgdf/src):
//   gdf_order_by: rejects a mask on the first column, fills the caller's d_cols / d_types device scratch, writes the
//   row indices (size_t) in lexicographic ascending order of the columns          sqls_ops.cu:1373-1392,27-41
//   comparator: typed `<` column after column (INT8..FLOAT64, dates as integers)  sqls_rtti_comp.hpp:84-175,299-320
//   the reference sorts with thrust::sort (not stable): the order of equal rows is unspecified there; here it is
//   ascending row order.
//
// B200 design.  The reference sorts an index array with a comparator that dereferences every column through a
// runtime type switch - a comparison sort with random global reads per comparison.  Here the order is built by LSD
// radix passes: for each key column, last to first, the column's values are gathered through the current
// permutation into order-preserving unsigned keys and {key, row id} pairs go through one stable 8-bit counting pass
// per key byte (chunked histogram -> scan -> stable scatter, ranks inside a warp from match.any).  A pass whose
// digit is the same for every row is detected on the device from its histogram and skipped, so narrow value ranges
// in wide columns (ids in int64) cost only the passes that carry information.  No host synchronisation anywhere.
#include <cstring>
#include <vector>

#include "sort.cuh"

namespace b200 {
namespace {

constexpr int kThreads = 256;
constexpr int kItems = 8;
constexpr int kTile = kThreads * kItems;  // 2048 rows
constexpr int kWarps = kThreads / 32;
constexpr int kMaxPasses = 8;

struct SortState {              // device memory, one per column sort
  unsigned in_idx[kMaxPasses + 1];   // which of the two {key, id} buffers holds the input of pass p
  unsigned trivial[kMaxPasses];      // pass p moves nothing (every row has the same digit)
};

struct Buffers {
  unsigned long long* keys[2];
  uint32_t* ids[2];
};

// order-preserving map of a typed value of W bytes to an unsigned integer
template <int W>
static __device__ __forceinline__ unsigned long long sortable(unsigned long long bits, bool is_float) {
  constexpr unsigned long long sign = 1ull << (8 * W - 1);
  constexpr unsigned long long all = W == 8 ? ~0ull : ((1ull << (8 * (W & 7))) - 1ull);
  if (is_float) return (bits & sign) ? (~bits & all) : (bits | sign);
  return bits ^ sign;
}

__global__ void iota_kernel(uint32_t* perm, size_t n) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) perm[i] = (uint32_t)i;
}

template <int W>
__global__ void gather_keys_kernel(const void* __restrict__ col, bool is_float, const uint32_t* __restrict__ perm, size_t n,
                                   unsigned long long* __restrict__ keys, uint32_t* __restrict__ ids, SortState* st) {
  if (blockIdx.x == 0 && threadIdx.x == 0) st->in_idx[0] = 0;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const uint32_t r = perm[i];
    unsigned long long bits;
    if (W == 1) bits = static_cast<const uint8_t*>(col)[r];
    else if (W == 2) bits = static_cast<const uint16_t*>(col)[r];
    else if (W == 4) bits = static_cast<const uint32_t*>(col)[r];
    else bits = static_cast<const unsigned long long*>(col)[r];
    keys[i] = sortable<W>(bits, is_float);
    ids[i] = r;
  }
}

// block b owns the rows [b * chunk_rows, (b + 1) * chunk_rows): the same split in the histogram and the scatter
__global__ void __launch_bounds__(kThreads)
hist_kernel(Buffers buf, const SortState* __restrict__ st, int pass, size_t n, size_t chunk_rows, unsigned* __restrict__ ghist) {
  __shared__ unsigned hist[256];
  hist[threadIdx.x] = 0;
  __syncthreads();
  const unsigned long long* keys = buf.keys[st->in_idx[pass]];
  const size_t lo = (size_t)blockIdx.x * chunk_rows;
  const size_t hi = lo + chunk_rows < n ? lo + chunk_rows : n;
  const int shift = 8 * pass;
  for (size_t i = lo + threadIdx.x; i < hi; i += kThreads) atomicAdd(&hist[(unsigned)(keys[i] >> shift) & 255u], 1u);
  __syncthreads();
  ghist[(size_t)threadIdx.x * gridDim.x + blockIdx.x] = hist[threadIdx.x];
}

// exclusive scan of ghist in (digit, block) order, in place; decides whether the pass is trivial
__global__ void __launch_bounds__(1024)
scan_kernel(unsigned* __restrict__ ghist, unsigned nblocks, size_t n, SortState* __restrict__ st, int pass) {
  __shared__ unsigned warp_sums[32];
  __shared__ unsigned is_trivial;
  const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  const size_t total = (size_t)256 * nblocks;
  if (tid == 0) is_trivial = 0;
  __syncthreads();
  if (tid < 256) {
    size_t sum = 0;
    for (unsigned b = 0; b < nblocks; ++b) sum += ghist[(size_t)tid * nblocks + b];
    if (sum == n) is_trivial = 1;
  }
  __syncthreads();
  const size_t per = (total + 1023) / 1024;
  const size_t lo = (size_t)tid * per, hi = lo + per < total ? lo + per : total;
  unsigned sum = 0;
  for (size_t i = lo; i < hi; ++i) sum += ghist[i];
  unsigned inc = sum;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const unsigned o = __shfl_up_sync(0xffffffffu, inc, d);
    if (lane >= (unsigned)d) inc += o;
  }
  if (lane == 31) warp_sums[warp] = inc;
  __syncthreads();
  unsigned off = 0;
  for (unsigned w = 0; w < warp; ++w) off += warp_sums[w];
  unsigned run = off + inc - sum;
  for (size_t i = lo; i < hi; ++i) {
    const unsigned c = ghist[i];
    ghist[i] = run;
    run += c;
  }
  if (tid == 0) {
    st->trivial[pass] = is_trivial;
    st->in_idx[pass + 1] = st->in_idx[pass] ^ (is_trivial ? 0u : 1u);
  }
}

__global__ void __launch_bounds__(kThreads)
scatter_kernel(Buffers buf, const SortState* __restrict__ st, int pass, size_t n, size_t chunk_rows,
               const unsigned* __restrict__ gprefix) {
  if (st->trivial[pass]) return;
  __shared__ unsigned offsets[256];
  __shared__ unsigned warp_cnt[kWarps][256];
  const unsigned in = st->in_idx[pass];
  const unsigned long long* __restrict__ keys_in = buf.keys[in];
  const uint32_t* __restrict__ ids_in = buf.ids[in];
  unsigned long long* __restrict__ keys_out = buf.keys[in ^ 1u];
  uint32_t* __restrict__ ids_out = buf.ids[in ^ 1u];
  const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  const int shift = 8 * pass;
  offsets[tid] = gprefix[(size_t)tid * gridDim.x + blockIdx.x];
  const size_t lo = (size_t)blockIdx.x * chunk_rows;
  const size_t hi = lo + chunk_rows < n ? lo + chunk_rows : n;
  for (size_t tile = lo; tile < hi; tile += kTile) {
#pragma unroll
    for (int w = 0; w < kWarps; ++w) warp_cnt[w][tid] = 0;
    __syncthreads();
    // a warp owns 32 * kItems consecutive rows; item j of lane l is row seg + j * 32 + l: (warp, j, lane) order = row order
    const size_t seg = tile + (size_t)warp * (32 * kItems);
    unsigned long long key[kItems];
    uint32_t id[kItems];
    unsigned rank[kItems], digit[kItems];
#pragma unroll
    for (int j = 0; j < kItems; ++j) {
      const size_t i = seg + (size_t)j * 32 + lane;
      const bool have = i < hi;
      key[j] = have ? keys_in[i] : 0ull;
      id[j] = have ? ids_in[i] : 0u;
      digit[j] = have ? ((unsigned)(key[j] >> shift) & 255u) : 0xffffu;
    }
#pragma unroll
    for (int j = 0; j < kItems; ++j) {
      const unsigned m = __match_any_sync(0xffffffffu, digit[j]);
      unsigned base = 0;
      if (digit[j] != 0xffffu) base = warp_cnt[warp][digit[j]];
      rank[j] = base + __popc(m & ((1u << lane) - 1u));
      __syncwarp();
      if (digit[j] != 0xffffu && (m & ((1u << lane) - 1u)) == 0) warp_cnt[warp][digit[j]] = base + __popc(m);  // group leader
      __syncwarp();
    }
    __syncthreads();
    unsigned tile_cnt = 0;
    {  // thread d: exclusive prefix of digit d's counts over the warps
      unsigned run = 0;
#pragma unroll
      for (int w = 0; w < kWarps; ++w) {
        const unsigned c = warp_cnt[w][tid];
        warp_cnt[w][tid] = run;
        run += c;
      }
      tile_cnt = run;
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < kItems; ++j) {
      if (digit[j] == 0xffffu) continue;
      const size_t pos = (size_t)offsets[digit[j]] + warp_cnt[warp][digit[j]] + rank[j];
      keys_out[pos] = key[j];
      ids_out[pos] = id[j];
    }
    __syncthreads();
    offsets[tid] += tile_cnt;
  }
}

__global__ void collect_kernel(Buffers buf, const SortState* __restrict__ st, int passes, size_t n, uint32_t* __restrict__ perm) {
  const uint32_t* ids = buf.ids[st->in_idx[passes]];
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) perm[i] = ids[i];
}

template <int W>
__global__ void gather_rows_kernel(const void* __restrict__ in, void* __restrict__ out, size_t n, const uint32_t* __restrict__ perm) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const uint32_t r = perm[i];
    if (W == 1) static_cast<uint8_t*>(out)[i] = static_cast<const uint8_t*>(in)[r];
    else if (W == 2) static_cast<uint16_t*>(out)[i] = static_cast<const uint16_t*>(in)[r];
    else if (W == 4) static_cast<uint32_t*>(out)[i] = static_cast<const uint32_t*>(in)[r];
    else static_cast<unsigned long long*>(out)[i] = static_cast<const unsigned long long*>(in)[r];
  }
}

__global__ void widen_indices_kernel(const uint32_t* __restrict__ perm, size_t n, size_t* __restrict__ out) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = perm[i];
}

int grid_for(size_t items) {
  size_t want = (items + kThreads - 1) / kThreads;
  const size_t cap = (size_t)sm_count() * 8;
  return (int)(want < 1 ? 1 : (want < cap ? want : cap));
}

bool float_dtype(int t) { return t == GDF_FLOAT32 || t == GDF_FLOAT64; }

}  // namespace

gdf_error gather_rows(const void* in, void* out, int width, size_t n, const uint32_t* d_perm) {
  if (n == 0) return GDF_SUCCESS;
  const int g = grid_for(n);
  switch (width) {
    case 1: gather_rows_kernel<1><<<g, kThreads>>>(in, out, n, d_perm); break;
    case 2: gather_rows_kernel<2><<<g, kThreads>>>(in, out, n, d_perm); break;
    case 4: gather_rows_kernel<4><<<g, kThreads>>>(in, out, n, d_perm); break;
    case 8: gather_rows_kernel<8><<<g, kThreads>>>(in, out, n, d_perm); break;
    default: return GDF_UNSUPPORTED_DTYPE;
  }
  B200_CHECK_LAST();
  return GDF_SUCCESS;
}

gdf_error permute_in_place(void* data, int width, size_t n, const uint32_t* d_perm) {
  if (n == 0) return GDF_SUCCESS;
  Scratch tmp;
  B200_CUDA_TRY(tmp.alloc(n * (size_t)width));
  gdf_error e = gather_rows(data, tmp.ptr, width, n, d_perm);
  if (e != GDF_SUCCESS) return e;
  B200_CUDA_TRY(cudaMemcpyAsync(data, tmp.ptr, n * (size_t)width, cudaMemcpyDeviceToDevice, 0));
  return GDF_SUCCESS;
}

gdf_error sort_permutation(const gdf_column* const* cols, int ncols, size_t n, uint32_t* d_perm) {
  B200_REQUIRE(n < 0xffffffffull, GDF_COLUMN_SIZE_TOO_BIG);
  if (n == 0) return GDF_SUCCESS;
  for (int c = 0; c < ncols; ++c) B200_REQUIRE(dtype_width(cols[c]->dtype) != 0, GDF_UNSUPPORTED_DTYPE);
  B200_TIMED("sort_permutation");
  iota_kernel<<<grid_for(n), kThreads>>>(d_perm, n);
  B200_CHECK_LAST();
  // chunked split of the rows over at most 4 CTAs per SM, whole tiles per chunk
  const size_t tiles = (n + kTile - 1) / kTile;
  size_t nblocks = (size_t)sm_count() * 4;
  if (nblocks > tiles) nblocks = tiles;
  const size_t chunk_rows = ((tiles + nblocks - 1) / nblocks) * kTile;
  nblocks = (n + chunk_rows - 1) / chunk_rows;
  Scratch k0, k1, i0, i1, hist, state;
  B200_CUDA_TRY(k0.alloc(n * sizeof(unsigned long long)));
  B200_CUDA_TRY(k1.alloc(n * sizeof(unsigned long long)));
  B200_CUDA_TRY(i0.alloc(n * sizeof(uint32_t)));
  B200_CUDA_TRY(i1.alloc(n * sizeof(uint32_t)));
  B200_CUDA_TRY(hist.alloc(256 * nblocks * sizeof(unsigned)));
  B200_CUDA_TRY(state.alloc(sizeof(SortState)));
  Buffers buf{{k0.as<unsigned long long>(), k1.as<unsigned long long>()}, {i0.as<uint32_t>(), i1.as<uint32_t>()}};
  SortState* st = state.as<SortState>();
  for (int c = ncols - 1; c >= 0; --c) {
    const int w = dtype_width(cols[c]->dtype);
    const bool fl = float_dtype(cols[c]->dtype);
    const int g = grid_for(n);
    switch (w) {
      case 1: gather_keys_kernel<1><<<g, kThreads>>>(cols[c]->data, fl, d_perm, n, buf.keys[0], buf.ids[0], st); break;
      case 2: gather_keys_kernel<2><<<g, kThreads>>>(cols[c]->data, fl, d_perm, n, buf.keys[0], buf.ids[0], st); break;
      case 4: gather_keys_kernel<4><<<g, kThreads>>>(cols[c]->data, fl, d_perm, n, buf.keys[0], buf.ids[0], st); break;
      default: gather_keys_kernel<8><<<g, kThreads>>>(cols[c]->data, fl, d_perm, n, buf.keys[0], buf.ids[0], st); break;
    }
    for (int pass = 0; pass < w; ++pass) {
      hist_kernel<<<(unsigned)nblocks, kThreads>>>(buf, st, pass, n, chunk_rows, hist.as<unsigned>());
      scan_kernel<<<1, 1024>>>(hist.as<unsigned>(), (unsigned)nblocks, n, st, pass);
      scatter_kernel<<<(unsigned)nblocks, kThreads>>>(buf, st, pass, n, chunk_rows, hist.as<unsigned>());
    }
    collect_kernel<<<g, kThreads>>>(buf, st, w, n, d_perm);
    B200_CHECK_LAST();
  }
  return GDF_SUCCESS;
}

}  // namespace b200

using namespace b200;

// ref sqls_ops.cu:1373-1392
extern "C" gdf_error gdf_order_by(size_t nrows, gdf_column* cols, size_t ncols, void** d_cols, int* d_types, size_t* d_indx) {
  B200_REQUIRE(cols != nullptr && ncols > 0, GDF_DATASET_EMPTY);
  B200_REQUIRE(!cols->valid, GDF_VALIDITY_UNSUPPORTED);
  B200_REQUIRE(nrows < 0x7fffffffull, GDF_COLUMN_SIZE_TOO_BIG);
  std::vector<void*> h_cols(ncols);
  std::vector<int> h_types(ncols);
  std::vector<const gdf_column*> ptrs(ncols);
  for (size_t c = 0; c < ncols; ++c) {
    h_cols[c] = cols[c].data;
    h_types[c] = (int)cols[c].dtype;
    ptrs[c] = &cols[c];
    B200_REQUIRE(dtype_width(cols[c].dtype) != 0 && cols[c].dtype <= GDF_FLOAT64, GDF_UNSUPPORTED_DTYPE);  // sqls_rtti_comp.hpp:224-272
  }
  // the reference fills the caller's device scratch with the columns' data pointers and dtypes (soa_col_info)
  if (d_cols) B200_CUDA_TRY(cudaMemcpy(d_cols, h_cols.data(), ncols * sizeof(void*), cudaMemcpyHostToDevice));
  if (d_types) B200_CUDA_TRY(cudaMemcpy(d_types, h_types.data(), ncols * sizeof(int), cudaMemcpyHostToDevice));
  if (nrows == 0) return GDF_SUCCESS;
  B200_REQUIRE(d_indx != nullptr, GDF_DATASET_EMPTY);
  Scratch perm;
  B200_CUDA_TRY(perm.alloc(nrows * sizeof(uint32_t)));
  gdf_error e = sort_permutation(ptrs.data(), (int)ncols, nrows, perm.as<uint32_t>());
  if (e != GDF_SUCCESS) return e;
  widen_indices_kernel<<<grid_for(nrows), kThreads>>>(perm.as<uint32_t>(), nrows, d_indx);
  B200_CHECK_LAST();
  return GDF_SUCCESS;
}
