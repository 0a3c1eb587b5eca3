// gdf_hash (row hash -> int32 column) and gdf_hash_partition (reorder rows into hash partitions).
//
// Reference behaviour followed (libgdf/src/hashing.cu):
//   gdf_hash            :83-154   null/empty checks, output must be GDF_INT32, MURMUR3 or IDENTITY
//   gdf_hash_partition  :559-654  argument checks and their error codes, `int` partition offsets
//                                 returned on the HOST as the exclusive scan of partition sizes,
//                                 power-of-two partition counts use hash & (n-1), others the unsigned
//                                 32-bit hash % n (:196-237); row order inside a partition is
//                                 unspecified; a column's validity bits move with its rows only when
//                                 both the input and the output column carry a mask
//                                 (gdf_table.cuh:1102-1116)
//
// B200 design.  The reference materialises a partition id per row (4 B write + 4 B re-read), scans a
// (blocks x partitions) matrix, writes a scatter map and then runs one Thrust scatter per column
// with fully uncoalesced 8-byte stores.  Here:
//   pass 1  persistent grid, per-CTA shared-memory histogram, one global atomic per partition per CTA;
//   scan    one CTA;
//   pass 2  tiles of 2048 rows: the row hash is RECOMPUTED (cheaper than 8 B/row of id traffic),
//           ranks inside the tile come from shared-memory atomics, each (tile, partition) run reserves
//           its output range with ONE global atomic, and all columns are moved in the same kernel.
// Partition counts above kSmemPartitions fall back to per-row global cursors (still one kernel).
#include "table.cuh"

namespace b200 {
namespace {

constexpr int kThreads = 256;
constexpr int kRowsPerThread = 8;
constexpr int kTileRows = kThreads * kRowsPerThread;
constexpr int kSmemPartitions = 2048;

struct Partitioner {
  unsigned n;
  bool pow2;
  __device__ __forceinline__ unsigned operator()(uint32_t h) const { return pow2 ? (h & (n - 1)) : (h % n); }
};

template <bool IDENTITY>
__global__ void __launch_bounds__(kThreads) hash_rows_kernel(TableView tv, int32_t* __restrict__ out) {
  const size_t stride = (size_t)gridDim.x * kThreads;
  for (size_t r = (size_t)blockIdx.x * kThreads + threadIdx.x; r < tv.rows; r += stride)
    out[r] = (int32_t)row_hash<IDENTITY>(tv, r);
}

template <bool IDENTITY, bool SMEM>
__global__ void __launch_bounds__(kThreads) partition_hist_kernel(TableView keys, Partitioner part,
                                                                  unsigned* __restrict__ totals) {
  extern __shared__ unsigned hist[];
  if (SMEM) {
    for (unsigned p = threadIdx.x; p < part.n; p += kThreads) hist[p] = 0;
    __syncthreads();
  }
  const size_t stride = (size_t)gridDim.x * kThreads;
  for (size_t r = (size_t)blockIdx.x * kThreads + threadIdx.x; r < keys.rows; r += stride) {
    const unsigned p = part(row_hash<IDENTITY>(keys, r));
    atomicAdd(SMEM ? &hist[p] : &totals[p], 1u);
  }
  if (SMEM) {
    __syncthreads();
    for (unsigned p = threadIdx.x; p < part.n; p += kThreads)
      if (hist[p]) atomicAdd(&totals[p], hist[p]);
  }
}

// totals[P] -> offsets[P] (exclusive scan, int) and cursors[P] (same values, consumed by pass 2)
__global__ void __launch_bounds__(1024) partition_scan_kernel(const unsigned* __restrict__ totals, unsigned n,
                                                              int* __restrict__ offsets,
                                                              unsigned* __restrict__ cursors) {
  __shared__ unsigned chunk_sum[1024];
  const unsigned per = (n + 1023) / 1024;
  const unsigned lo = threadIdx.x * per, hi = min(n, lo + per);
  unsigned s = 0;
  for (unsigned i = lo; i < hi; ++i) s += totals[i];
  chunk_sum[threadIdx.x] = s;
  __syncthreads();
  for (unsigned d = 1; d < 1024; d <<= 1) {  // inclusive Hillis-Steele over the 1024 chunk sums
    unsigned v = threadIdx.x >= d ? chunk_sum[threadIdx.x - d] : 0;
    __syncthreads();
    chunk_sum[threadIdx.x] += v;
    __syncthreads();
  }
  unsigned run = threadIdx.x ? chunk_sum[threadIdx.x - 1] : 0;
  for (unsigned i = lo; i < hi; ++i) {
    offsets[i] = (int)run;
    cursors[i] = run;
    run += totals[i];
  }
}

struct ColumnMove {  // every input column and its destination
  const void* in[kMaxCols];
  void* out[kMaxCols];
  const gdf_valid_type* in_valid[kMaxCols];
  gdf_valid_type* out_valid[kMaxCols];
  unsigned char width[kMaxCols];
  int ncols;
};

static __device__ __forceinline__ void move_value(const ColumnMove& m, int c, size_t from, size_t to) {
  switch (m.width[c]) {
    case 1: static_cast<uint8_t*>(m.out[c])[to] = static_cast<const uint8_t*>(m.in[c])[from]; break;
    case 2: static_cast<uint16_t*>(m.out[c])[to] = static_cast<const uint16_t*>(m.in[c])[from]; break;
    case 4: static_cast<uint32_t*>(m.out[c])[to] = static_cast<const uint32_t*>(m.in[c])[from]; break;
    default: static_cast<uint64_t*>(m.out[c])[to] = static_cast<const uint64_t*>(m.in[c])[from]; break;
  }
  if (m.in_valid[c] && m.out_valid[c] && bit_valid(m.in_valid[c], from)) {
    // set bit `to` of the (pre-zeroed) output mask; 32-bit atomic on the enclosing aligned word
    uintptr_t byte_addr = reinterpret_cast<uintptr_t>(m.out_valid[c] + (to >> 3));
    unsigned* word = reinterpret_cast<unsigned*>(byte_addr & ~(uintptr_t)3);
    const unsigned shift = (unsigned)(byte_addr & 3) * 8 + (unsigned)(to & 7);
    atomicOr(word, 1u << shift);
  }
}

// Pass 2.  SMEM (P <= 2048): the tile's rows are first ORDERED by partition in shared memory (order[j] = row of the tile
// that lands at sorted position j, ppos[j] = its partition), then every column is moved with consecutive threads on
// consecutive sorted positions - which are consecutive output addresses inside a partition's run - so the stores of a
// warp are coalesced runs instead of 32 scattered elements, and the loads stay inside the tile's own 2048-row window of
// the input column (every fetched sector is used).  The first version moved rows in input order with one scattered
// store per row and column: 1.7 TB/s at 2.5e8 x (int64, int64) rows into 8 partitions (profiles/r02_notes.md).
template <bool IDENTITY, bool SMEM>
__global__ void __launch_bounds__(kThreads) partition_scatter_kernel(TableView keys, Partitioner part,
                                                                     ColumnMove mv,
                                                                     unsigned* __restrict__ cursors) {
  extern __shared__ unsigned sm[];
  unsigned* hist = sm;                 // rows of this tile per partition
  unsigned* base = sm + part.n;        // reserved global start of this tile's run per partition
  unsigned* lstart = sm + 2 * part.n;  // start of the partition's run inside the sorted tile
  unsigned short* order = reinterpret_cast<unsigned short*>(sm + 3 * part.n);  // [kTileRows]
  unsigned short* ppos = order + kTileRows;                                    // [kTileRows]
  __shared__ unsigned warp_sums[kThreads / 32];
  const size_t rows = keys.rows;
  const size_t tiles = (rows + kTileRows - 1) / kTileRows;
  const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  for (size_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const size_t tile_base = tile * kTileRows;
    unsigned pid[kRowsPerThread], rank[kRowsPerThread];
    if (SMEM) {
      for (unsigned p = threadIdx.x; p < part.n; p += kThreads) hist[p] = 0;
      __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < kRowsPerThread; ++i) {
      const size_t r = tile_base + (size_t)i * kThreads + threadIdx.x;
      if (r < rows) {
        pid[i] = part(row_hash<IDENTITY>(keys, r));
        rank[i] = SMEM ? atomicAdd(&hist[pid[i]], 1u) : atomicAdd(&cursors[pid[i]], 1u);
      }
    }
    if (!SMEM) {  // too many partitions for shared memory: rows move in input order, positions from global cursors
#pragma unroll 1
      for (int c = 0; c < mv.ncols; ++c) {
#pragma unroll
        for (int i = 0; i < kRowsPerThread; ++i) {
          const size_t r = tile_base + (size_t)i * kThreads + threadIdx.x;
          if (r < rows) move_value(mv, c, r, rank[i]);
        }
      }
      continue;
    }
    __syncthreads();
    {  // reserve the runs (one global atomic per non-empty partition) and scan the tile's counts
      const unsigned per = (part.n + kThreads - 1) / kThreads;
      const unsigned lo = threadIdx.x * per, hi = min(part.n, lo + per);
      unsigned sum = 0;
      for (unsigned p = lo; p < hi; ++p) {
        const unsigned h = hist[p];
        if (h) base[p] = atomicAdd(&cursors[p], h);
        sum += h;
      }
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
      for (unsigned p = lo; p < hi; ++p) {
        lstart[p] = run;
        run += hist[p];
      }
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < kRowsPerThread; ++i) {
      const size_t r = tile_base + (size_t)i * kThreads + threadIdx.x;
      if (r < rows) {
        const unsigned j = lstart[pid[i]] + rank[i];
        order[j] = (unsigned short)(i * kThreads + threadIdx.x);
        ppos[j] = (unsigned short)pid[i];
      }
    }
    __syncthreads();
    const unsigned live = (unsigned)(rows - tile_base < (size_t)kTileRows ? rows - tile_base : (size_t)kTileRows);
#pragma unroll 1
    for (int c = 0; c < mv.ncols; ++c) {
#pragma unroll 4
      for (unsigned j = threadIdx.x; j < live; j += kThreads) {
        const unsigned p = ppos[j];
        move_value(mv, c, tile_base + order[j], (size_t)base[p] + (j - lstart[p]));
      }
    }
    __syncthreads();
  }
}

template <bool IDENTITY>
gdf_error run_partition(const TableView& keys, const ColumnMove& mv, int num_partitions, int* h_offsets) {
  Partitioner part{(unsigned)num_partitions, (num_partitions & (num_partitions - 1)) == 0};
  const size_t P = (size_t)num_partitions;
  Scratch buf;
  B200_CUDA_TRY(buf.alloc(P * (2 * sizeof(unsigned) + sizeof(int))));
  unsigned* totals = buf.as<unsigned>();
  unsigned* cursors = totals + P;
  int* d_offsets = reinterpret_cast<int*>(cursors + P);
  B200_CUDA_TRY(cudaMemsetAsync(totals, 0, P * sizeof(unsigned), 0));
  const bool smem = num_partitions <= kSmemPartitions;
  const size_t rows = keys.rows;
  size_t want = (rows + kTileRows - 1) / kTileRows;
  const size_t cap = (size_t)sm_count() * 4;
  const int blocks = (int)(want < cap ? (want ? want : 1) : cap);
  B200_TIMED("hash_partition");
  if (smem)
    partition_hist_kernel<IDENTITY, true><<<blocks, kThreads, P * sizeof(unsigned)>>>(keys, part, totals);
  else
    partition_hist_kernel<IDENTITY, false><<<blocks, kThreads>>>(keys, part, totals);
  B200_CHECK_LAST();
  partition_scan_kernel<<<1, 1024>>>(totals, part.n, d_offsets, cursors);
  B200_CHECK_LAST();
  if (smem)
    partition_scatter_kernel<IDENTITY, true><<<blocks, kThreads, 3 * P * sizeof(unsigned) + 2 * kTileRows * sizeof(unsigned short)>>>(keys, part, mv, cursors);
  else
    partition_scatter_kernel<IDENTITY, false><<<blocks, kThreads>>>(keys, part, mv, cursors);
  B200_CHECK_LAST();
  // blocking copy: also the point where the call becomes synchronous, like the reference (:531)
  B200_CUDA_TRY(cudaMemcpy(h_offsets, d_offsets, P * sizeof(int), cudaMemcpyDeviceToHost));
  return GDF_SUCCESS;
}

}  // namespace
}  // namespace b200

using namespace b200;

extern "C" gdf_error gdf_hash(int num_cols, gdf_column** input, gdf_hash_func hash, gdf_column* output) {
  if (num_cols == 0 || input == nullptr || output == nullptr) return GDF_DATASET_EMPTY;
  if (output->dtype != GDF_INT32) return GDF_UNSUPPORTED_DTYPE;
  if (input[0] != nullptr && input[0]->size == 0) return GDF_SUCCESS;
  if (output->size == 0) return GDF_SUCCESS;
  if (output->data == nullptr) return GDF_DATASET_EMPTY;
  if (hash != GDF_HASH_MURMUR3 && hash != GDF_HASH_IDENTITY) return GDF_INVALID_HASH_FUNCTION;
  TableView tv;
  B200_REQUIRE(make_view(tv, input, num_cols), GDF_JOIN_TOO_MANY_COLUMNS);
  for (int c = 0; c < num_cols; ++c) B200_REQUIRE(hashable_dtype(input[c]->dtype), GDF_UNSUPPORTED_DTYPE);
  size_t want = (tv.rows + kThreads * 4 - 1) / (kThreads * 4);
  const size_t cap = (size_t)sm_count() * 8;
  const int blocks = (int)(want < cap ? (want ? want : 1) : cap);
  int32_t* out = static_cast<int32_t*>(output->data);
  if (hash == GDF_HASH_MURMUR3)
    hash_rows_kernel<false><<<blocks, kThreads>>>(tv, out);
  else
    hash_rows_kernel<true><<<blocks, kThreads>>>(tv, out);
  B200_CHECK_LAST();
  return GDF_SUCCESS;
}

extern "C" gdf_error gdf_hash_partition(int num_input_cols, gdf_column* input[], int columns_to_hash[],
                                        int num_cols_to_hash, int num_partitions,
                                        gdf_column* partitioned_output[], int partition_offsets[],
                                        gdf_hash_func hash) {
  if (num_input_cols == 0 || num_cols_to_hash == 0 || num_partitions == 0 || input == nullptr ||
      partitioned_output == nullptr || columns_to_hash == nullptr || partition_offsets == nullptr)
    return GDF_INVALID_API_CALL;
  B200_REQUIRE(num_input_cols > 0 && num_cols_to_hash > 0 && num_partitions > 0, GDF_INVALID_API_CALL);
  const size_t num_rows = input[0]->size;
  if (num_rows == 0) return GDF_SUCCESS;
  B200_REQUIRE(num_rows < 0x7fffffffull, GDF_COLUMN_SIZE_TOO_BIG);  // offsets are `int` in this ABI
  for (int i = 0; i < num_input_cols; ++i) {
    if (input[i]->data == nullptr || partitioned_output[i]->data == nullptr) return GDF_DATASET_EMPTY;
    if (input[i]->dtype != partitioned_output[i]->dtype) return GDF_PARTITION_DTYPE_MISMATCH;
    if (num_rows != input[i]->size || num_rows != partitioned_output[i]->size) return GDF_COLUMN_SIZE_MISMATCH;
  }
  if (hash != GDF_HASH_MURMUR3 && hash != GDF_HASH_IDENTITY) return GDF_INVALID_HASH_FUNCTION;
  B200_REQUIRE(num_input_cols <= kMaxCols && num_cols_to_hash <= kMaxCols, GDF_JOIN_TOO_MANY_COLUMNS);

  gdf_column* key_cols[kMaxCols];
  for (int i = 0; i < num_cols_to_hash; ++i) {
    B200_REQUIRE(columns_to_hash[i] >= 0 && columns_to_hash[i] < num_input_cols, GDF_INVALID_API_CALL);
    key_cols[i] = input[columns_to_hash[i]];
    B200_REQUIRE(hashable_dtype(key_cols[i]->dtype), GDF_UNSUPPORTED_DTYPE);
  }
  TableView keys;
  make_view(keys, key_cols, num_cols_to_hash);
  ColumnMove mv;
  mv.ncols = num_input_cols;
  for (int i = 0; i < num_input_cols; ++i) {
    const int w = dtype_width(input[i]->dtype);
    B200_REQUIRE(w != 0, GDF_UNSUPPORTED_DTYPE);
    mv.in[i] = input[i]->data;
    mv.out[i] = partitioned_output[i]->data;
    mv.in_valid[i] = input[i]->valid;
    mv.out_valid[i] = partitioned_output[i]->valid;
    mv.width[i] = (unsigned char)w;
    if (mv.in_valid[i] && mv.out_valid[i])
      B200_CUDA_TRY(cudaMemsetAsync(mv.out_valid[i], 0, valid_bytes(num_rows), 0));
  }
  gdf_nvtx_range_push("LIBGDF_HASH_PARTITION", GDF_PURPLE);
  gdf_error err = (hash == GDF_HASH_MURMUR3) ? run_partition<false>(keys, mv, num_partitions, partition_offsets)
                                             : run_partition<true>(keys, mv, num_partitions, partition_offsets);
  gdf_nvtx_range_pop();
  return err;
}
