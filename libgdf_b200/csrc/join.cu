// Hash joins: gdf_inner_join / gdf_left_join / gdf_full_join (GDF_HASH method).
//
// Reference behaviour followed (file:line in /root/reference/libgdf/src/join):
//   argument checks, empty-input short cuts, size limit, dtype/size checks   joining.cu:283-373,482-569
//   build on the right table; for INNER build on the smaller side and flip   joining.h:49-74
//   a row with a NULL in any key column is never inserted and never matches; LEFT emits (l,-1)
//                                                                            hash/join_kernels.cuh:59,316
//   FULL = LEFT + one (-1, r) pair per build row that never appears in the right output (this also
//   covers build rows with NULL keys)                                        hash/join_compute_api.h:147-186
//   outputs: two GDF_INT32 index columns of equal length allocated with rmmAlloc, valid=NULL,
//   null_count=0; no match at all -> {data=NULL,size=0,dtype=N_GDF_TYPES}    :353-354,438-441,547-548
//   trivial FULL join when one side is empty                                 joining.cu:228-264
//   optional result_cols = [left non-key.., key.., right non-key..] gathered with the index columns
//                                                                            joining.cu:376-479
//
// B200 design.  The reference probes the probe table two to three times (sampled size estimate,
// probe, retry when the estimate was low), synchronises the device five times and finally copies
// the result into right-sized buffers.  Here:
//   * the build kernel detects duplicate build keys while it inserts (the later of two equal keys
//     always walks over the earlier one), so when build keys are unique - the common PK/FK case -
//     the output bound is known (<= probe rows) and the probe runs exactly ONCE;
//     with duplicates an exact count pass replaces the estimate/retry loop;
//   * the probe kernel works on tiles: count matches, block-scan, ONE cursor atomic per 1024-row
//     tile (the reference issues one per warp flush), then coalesced writes in (slice, thread) order;
//   * table slots are one packed 64-bit word {hash32,row32} like the reference's, but EMPTY is the
//     all-ones word, which no real entry can equal because row ids are < 2^31 - the reference's
//     0xFFFFFFFF-hash sentinel collision cannot happen.
//   * single 4/8-byte integer key columns take the radix-partitioned path of join_part.cuh so that
//     every table probe is an L2 hit instead of a DRAM row miss.
#include <cstring>
#include <vector>

#include "select.cuh"
#include "table.cuh"

namespace b200 {

// radix-partitioned single-key path (join_part.cu)
gdf_error partitioned_join(int kind, const gdf_column* probe_key, const gdf_column* build_key, bool flip,
                           gdf_column* out_l, gdf_column* out_r, bool* handled, const int32_t* probe_payload = nullptr,
                           const int32_t* build_payload = nullptr, const gdf_column* probe_key2 = nullptr,
                           const gdf_column* build_key2 = nullptr);
gdf_error partition_pairs(const gdf_column* key, int32_t id_base, unsigned num_partitions, void* out_keys,
                          int32_t* out_ids, unsigned long long* h_offsets);
gdf_error partition_count(const gdf_column* key, unsigned num_partitions, unsigned long long* h_counts);
gdf_error partition_scatter_peer(const gdf_column* key, int32_t id_base, unsigned num_partitions, void* const* dst_keys,
                                 int32_t* const* dst_ids, const unsigned long long* dst_offsets);
gdf_error xjoin_count(const gdf_column* key, unsigned ranks, unsigned nlocal, unsigned long long* h_counts, unsigned* hi_or);
gdf_error xjoin_scatter(const gdf_column* key, int32_t id_base, unsigned ranks, unsigned nlocal, void* const* dst_pairs,
                        const unsigned long long* offsets, const unsigned long long* counts, const int* d_status, int ctas_per_sm);
gdf_error xjoin_build(const void* build_pairs, const unsigned long long* build_counts, unsigned nlocal, bool side, void** handle);
gdf_error xjoin_probe(void* handle, const void* probe_pairs, const unsigned long long* probe_counts, gdf_column* out_l,
                      gdf_column* out_r);
gdf_error xjoin_count_dev(const gdf_column* key, unsigned ranks, unsigned nlocal, unsigned long long* d_counts);
gdf_error xjoin_plan_dev(const unsigned long long* d_all, unsigned ranks, unsigned nlocal, unsigned rank, unsigned long long cap_build,
                         unsigned long long cap_probe, unsigned long long* d_off_build, unsigned long long* d_off_probe,
                         int* d_status);
gdf_error xjoin_local(const void* probe_pairs, const unsigned long long* probe_counts, const void* build_pairs,
                      const unsigned long long* build_counts, unsigned nlocal, gdf_column* out_l, gdf_column* out_r);

namespace {

constexpr int kThreads = 256;
constexpr int kRowsPerThread = 4;
constexpr int kTileRows = kThreads * kRowsPerThread;
constexpr unsigned long long kEmpty = ~0ull;

enum JoinKind { JOIN_INNER = 0, JOIN_LEFT = 1, JOIN_FULL = 2 };

__global__ void __launch_bounds__(kThreads) build_kernel(TableView build, unsigned long long* __restrict__ slots,
                                                         unsigned mask, int* __restrict__ has_dup) {
  const size_t stride = (size_t)gridDim.x * kThreads;
  for (size_t r = (size_t)blockIdx.x * kThreads + threadIdx.x; r < build.rows; r += stride) {
    if (!row_valid(build, r)) continue;
    const uint32_t h = row_hash<false>(build, r);
    const unsigned long long packed = ((unsigned long long)h << 32) | (unsigned long long)(uint32_t)r;
    unsigned s = h & mask;
    while (true) {
      unsigned long long cur = slots[s];
      if (cur == kEmpty) {
        const unsigned long long prev = atomicCAS(&slots[s], kEmpty, packed);
        if (prev == kEmpty) break;
        cur = prev;
      }
      if ((uint32_t)(cur >> 32) == h && *(volatile int*)has_dup == 0 &&
          rows_equal(build, r, build, (size_t)(uint32_t)cur))
        *has_dup = 1;
      s = (s + 1) & mask;
    }
  }
}

// block-wide exclusive scan of one unsigned per thread; returns the exclusive prefix, *total = sum
static __device__ __forceinline__ unsigned block_exclusive_scan(unsigned v, unsigned* warp_sums, unsigned* total) {
  const unsigned lane = lane_id(), warp = threadIdx.x >> 5;
  unsigned inc = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const unsigned o = __shfl_up_sync(0xffffffffu, inc, d);
    if (lane >= (unsigned)d) inc += o;
  }
  if (lane == 31) warp_sums[warp] = inc;
  __syncthreads();
  unsigned warp_off = 0, tot = 0;
#pragma unroll
  for (int w = 0; w < kThreads / 32; ++w) {
    const unsigned ws = warp_sums[w];
    if ((unsigned)w < warp) warp_off += ws;
    tot += ws;
  }
  __syncthreads();
  *total = tot;
  return warp_off + inc - v;
}

// Walk the chain of `h` calling f(build_row) for each equal row.
template <typename F>
static __device__ __forceinline__ void for_each_match(const TableView& probe, size_t pr, const TableView& build,
                                                      const unsigned long long* __restrict__ slots, unsigned mask,
                                                      uint32_t h, F f) {
  unsigned s = h & mask;
  while (true) {
    const unsigned long long cur = slots[s];
    if (cur == kEmpty) return;
    if ((uint32_t)(cur >> 32) == h) {
      const uint32_t br = (uint32_t)cur;
      if (rows_equal(probe, pr, build, (size_t)br)) f((int32_t)br);
    }
    s = (s + 1) & mask;
  }
}

// KIND: JOIN_INNER or JOIN_LEFT (FULL runs as LEFT).  WRITE=false only counts.
template <int KIND, bool WRITE>
__global__ void __launch_bounds__(kThreads)
probe_kernel(TableView probe, TableView build, const unsigned long long* __restrict__ slots, unsigned mask,
             int32_t* __restrict__ out_probe, int32_t* __restrict__ out_build,
             unsigned long long* __restrict__ cursor) {
  __shared__ unsigned warp_sums[kThreads / 32];
  __shared__ unsigned long long tile_base_out;
  const size_t tile_base = (size_t)blockIdx.x * kTileRows;
  unsigned cnt[kRowsPerThread];
  int32_t first[kRowsPerThread];
  uint32_t hashes[kRowsPerThread];
  bool valid[kRowsPerThread];
#pragma unroll
  for (int i = 0; i < kRowsPerThread; ++i) {
    const size_t r = tile_base + (size_t)i * kThreads + threadIdx.x;
    cnt[i] = 0;
    first[i] = -1;
    valid[i] = false;
    hashes[i] = 0;
    if (r < probe.rows) {
      if (row_valid(probe, r)) {
        valid[i] = true;
        hashes[i] = row_hash<false>(probe, r);
        unsigned c = 0;
        int32_t f0 = -1;
        for_each_match(probe, r, build, slots, mask, hashes[i], [&](int32_t br) {
          if (c == 0) f0 = br;
          ++c;
        });
        cnt[i] = c;
        first[i] = f0;
      }
      if (KIND == JOIN_LEFT && cnt[i] == 0) cnt[i] = 1;  // (l, -1)
    }
  }
  // order of output inside the tile: slice i major, thread minor -> coalesced stores
  unsigned excl[kRowsPerThread], slice_total[kRowsPerThread], tile_total = 0;
#pragma unroll
  for (int i = 0; i < kRowsPerThread; ++i) {
    excl[i] = block_exclusive_scan(cnt[i], warp_sums, &slice_total[i]);
    tile_total += slice_total[i];
  }
  if (threadIdx.x == 0) tile_base_out = tile_total ? atomicAdd(cursor, (unsigned long long)tile_total) : 0ull;
  if (!WRITE) return;
  __syncthreads();
  size_t pos0 = (size_t)tile_base_out;
#pragma unroll
  for (int i = 0; i < kRowsPerThread; ++i) {
    const size_t r = tile_base + (size_t)i * kThreads + threadIdx.x;
    size_t pos = pos0 + excl[i];
    if (cnt[i] == 1) {
      out_probe[pos] = (int32_t)r;
      out_build[pos] = first[i];  // -1 for an unmatched LEFT row
    } else if (cnt[i] > 1) {
      for_each_match(probe, r, build, slots, mask, hashes[i], [&](int32_t br) {
        out_probe[pos] = (int32_t)r;
        out_build[pos] = br;
        ++pos;
      });
    }
    pos0 += slice_total[i];
  }
}

__global__ void fill_kernel(int32_t* p, size_t n, int32_t value, bool iota) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) p[i] = iota ? (int32_t)i : value;
}

__global__ void mark_kernel(const int32_t* __restrict__ idx, size_t n, unsigned char* __restrict__ marks, size_t limit) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int32_t v = idx[i];
    if (v >= 0 && (size_t)v < limit) marks[v] = 1;
  }
}

// FULL join tail: every unmarked build row r appends (-1, r) after the first `found` pairs.
struct UnmatchedPolicy {
  static constexpr int V = 16;
  static constexpr int K = 2;
  const unsigned char* marks;
  int32_t* out_none;   // receives -1
  int32_t* out_rows;   // receives the build row
  size_t found;
  __device__ uint32_t flags(size_t warp_base, size_t n) const {
    uint32_t f = 0;
    const unsigned lane = lane_id();
#pragma unroll
    for (int k = 0; k < K; ++k)
#pragma unroll
      for (int j = 0; j < V; ++j) {
        const size_t r = warp_base + (size_t)k * 32 * V + (size_t)lane * V + j;
        if (r < n && marks[r] == 0) f |= 1u << (k * V + j);
      }
    return f;
  }
  __device__ void emit(size_t row, size_t pos) const {
    out_none[found + pos] = -1;
    out_rows[found + pos] = (int32_t)row;
  }
};

int grid_for(size_t items, int per_block = kThreads) {
  size_t want = (items + per_block - 1) / per_block;
  const size_t cap = (size_t)sm_count() * 8;
  return (int)(want < 1 ? 1 : (want < cap ? want : cap));
}

unsigned pow2_at_least(size_t x) {
  unsigned p = 1;
  while ((size_t)p < x && p < (1u << 31)) p <<= 1;
  return p;
}

gdf_error read_u64(const unsigned long long* d, unsigned long long* h) {
  unsigned long long* box = static_cast<unsigned long long*>(pinned_mailbox());
  B200_REQUIRE(box != nullptr, GDF_CUDA_ERROR);
  B200_CUDA_TRY(cudaMemcpyAsync(box, d, sizeof(unsigned long long), cudaMemcpyDeviceToHost, 0));
  B200_CUDA_TRY(cudaStreamSynchronize(0));
  *h = *box;
  return GDF_SUCCESS;
}

gdf_error alloc_index_pair(size_t capacity, int32_t** a, int32_t** b) {
  *a = *b = nullptr;
  B200_RMM_TRY(output_alloc((void**)a, (capacity ? capacity : 1) * sizeof(int32_t)));
  if (output_alloc((void**)b, (capacity ? capacity : 1) * sizeof(int32_t)) != RMM_SUCCESS) {
    rmmFree(*a, 0);
    *a = nullptr;
    return GDF_MEMORYMANAGER_ERROR;
  }
  return GDF_SUCCESS;
}

// Appends the FULL-join tail to (out_probe,out_build) which must have room for found + build_rows.
gdf_error append_unmatched(int32_t* out_probe, int32_t* out_build, size_t found, size_t build_rows, size_t* total) {
  Scratch marks;
  B200_CUDA_TRY(marks.alloc(build_rows));
  B200_CUDA_TRY(cudaMemsetAsync(marks.ptr, 0, build_rows, 0));
  if (found) {
    mark_kernel<<<grid_for(found), kThreads>>>(out_build, found, marks.as<unsigned char>(), build_rows);
    B200_CHECK_LAST();
  }
  UnmatchedPolicy pol{marks.as<unsigned char>(), out_probe, out_build, found};
  const size_t tiles = select_tiles<UnmatchedPolicy>(build_rows);
  Scratch desc;
  B200_CUDA_TRY(desc.alloc(tiles * sizeof(uint64_t) + sizeof(unsigned long long)));
  B200_CUDA_TRY(cudaMemsetAsync(desc.ptr, 0, tiles * sizeof(uint64_t) + sizeof(unsigned long long), 0));
  uint64_t* d = desc.as<uint64_t>();
  unsigned long long* d_count = reinterpret_cast<unsigned long long*>(d + tiles);
  select_kernel<UnmatchedPolicy><<<(unsigned)tiles, select_detail::kThreads>>>(pol, build_rows, d, d_count);
  B200_CHECK_LAST();
  unsigned long long extra = 0;
  gdf_error e = read_u64(d_count, &extra);
  if (e != GDF_SUCCESS) return e;
  *total = found + (size_t)extra;
  return GDF_SUCCESS;
}

void view_indices(gdf_column* c, int32_t* data, size_t n) {
  if (n == 0 && data == nullptr) gdf_column_view(c, nullptr, nullptr, 0, N_GDF_TYPES);
  else gdf_column_view(c, data, nullptr, n, GDF_INT32);
}

// probe = left table (after an INNER flip: the caller's right table), build = the other one.
gdf_error generic_hash_join(int kind, const TableView& probe, const TableView& build, bool flip,
                            gdf_column* out_l, gdf_column* out_r) {
  const size_t P = probe.rows, B = build.rows;
  const unsigned slots = pow2_at_least(B ? 2 * B : 1);
  Scratch table, misc;
  B200_CUDA_TRY(table.alloc((size_t)slots * sizeof(unsigned long long)));
  B200_CUDA_TRY(cudaMemsetAsync(table.ptr, 0xff, (size_t)slots * sizeof(unsigned long long), 0));
  B200_CUDA_TRY(misc.alloc(4 * sizeof(unsigned long long)));
  B200_CUDA_TRY(cudaMemsetAsync(misc.ptr, 0, 4 * sizeof(unsigned long long), 0));
  unsigned long long* cursor = misc.as<unsigned long long>();
  int* has_dup = reinterpret_cast<int*>(cursor + 1);
  if (B) {
    B200_TIMED("join_generic_build");
    build_kernel<<<grid_for(B), kThreads>>>(build, table.as<unsigned long long>(), slots - 1, has_dup);
    B200_CHECK_LAST();
  }
  int h_dup = 0;
  B200_CUDA_TRY(cudaMemcpy(&h_dup, has_dup, sizeof(int), cudaMemcpyDeviceToHost));

  const bool left_like = kind != JOIN_INNER;
  const unsigned tiles = (unsigned)((P + kTileRows - 1) / kTileRows);
  size_t bound = P;  // unique build keys: at most one pair per probe row
  if (h_dup) {       // exact count instead of the reference's estimate/retry loop
    if (left_like)
      probe_kernel<JOIN_LEFT, false><<<tiles, kThreads>>>(probe, build, table.as<unsigned long long>(), slots - 1,
                                                         nullptr, nullptr, cursor);
    else
      probe_kernel<JOIN_INNER, false><<<tiles, kThreads>>>(probe, build, table.as<unsigned long long>(), slots - 1,
                                                          nullptr, nullptr, cursor);
    B200_CHECK_LAST();
    unsigned long long exact = 0;
    gdf_error e = read_u64(cursor, &exact);
    if (e != GDF_SUCCESS) return e;
    B200_REQUIRE(exact < 0x7fffffffull * 4ull, GDF_COLUMN_SIZE_TOO_BIG);
    bound = (size_t)exact;
    B200_CUDA_TRY(cudaMemsetAsync(cursor, 0, sizeof(unsigned long long), 0));
  }
  const size_t capacity = bound + (kind == JOIN_FULL ? B : 0);
  int32_t *o_probe = nullptr, *o_build = nullptr;
  if (capacity == 0) {
    view_indices(out_l, nullptr, 0);
    view_indices(out_r, nullptr, 0);
    return GDF_SUCCESS;
  }
  gdf_error e = alloc_index_pair(capacity, &o_probe, &o_build);
  if (e != GDF_SUCCESS) return e;
  B200_TIMED("join_generic_probe");
  if (left_like)
    probe_kernel<JOIN_LEFT, true><<<tiles, kThreads>>>(probe, build, table.as<unsigned long long>(), slots - 1,
                                                      o_probe, o_build, cursor);
  else
    probe_kernel<JOIN_INNER, true><<<tiles, kThreads>>>(probe, build, table.as<unsigned long long>(), slots - 1,
                                                       o_probe, o_build, cursor);
  cudaError_t ce = cudaPeekAtLastError();
  unsigned long long found = 0;
  if (ce == cudaSuccess) e = read_u64(cursor, &found);
  size_t total = (size_t)found;
  if (ce == cudaSuccess && e == GDF_SUCCESS && kind == JOIN_FULL) e = append_unmatched(o_probe, o_build, total, B, &total);
  if (ce != cudaSuccess || e != GDF_SUCCESS) {
    rmmFree(o_probe, 0);
    rmmFree(o_build, 0);
    return ce != cudaSuccess ? GDF_CUDA_ERROR : e;
  }
  if (total == 0) {
    rmmFree(o_probe, 0);
    rmmFree(o_build, 0);
    view_indices(out_l, nullptr, 0);
    view_indices(out_r, nullptr, 0);
    return GDF_SUCCESS;
  }
  view_indices(flip ? out_r : out_l, o_probe, total);
  view_indices(flip ? out_l : out_r, o_build, total);
  return GDF_SUCCESS;
}

gdf_error trivial_full_join(size_t left_size, size_t right_size, gdf_column* out_l, gdf_column* out_r) {
  if (left_size == 0 && right_size == 0) return GDF_DATASET_EMPTY;
  const size_t n = left_size ? left_size : right_size;
  int32_t *l = nullptr, *r = nullptr;
  gdf_error e = alloc_index_pair(n, &l, &r);
  if (e != GDF_SUCCESS) return e;
  fill_kernel<<<grid_for(n), kThreads>>>(l, n, -1, left_size != 0);
  fill_kernel<<<grid_for(n), kThreads>>>(r, n, -1, left_size == 0);
  gdf_column_view(out_l, l, nullptr, n, GDF_INT32);
  gdf_column_view(out_r, r, nullptr, n, GDF_INT32);
  B200_CHECK_LAST();
  return GDF_SUCCESS;
}

gdf_error join_call(int kind, int num_cols, gdf_column** leftcol, gdf_column** rightcol, gdf_column* out_l,
                    gdf_column* out_r, gdf_context* ctx) {
  if (num_cols == 0 || leftcol == nullptr || rightcol == nullptr) return GDF_DATASET_EMPTY;
  if (ctx == nullptr) return GDF_INVALID_API_CALL;
  const size_t L = leftcol[0]->size, R = rightcol[0]->size;
  if (L >= 0x7fffffffull || R >= 0x7fffffffull) return GDF_COLUMN_SIZE_TOO_BIG;
  if (L == 0 && R == 0) return GDF_SUCCESS;
  if (kind == JOIN_LEFT && L == 0) return GDF_SUCCESS;
  if (kind == JOIN_INNER && (L == 0 || R == 0)) return GDF_SUCCESS;
  if (kind == JOIN_FULL && (L == 0 || R == 0)) return trivial_full_join(L, R, out_l, out_r);
  for (int i = 0; i < num_cols; ++i) {
    if (R > 0 && rightcol[i]->data == nullptr) return GDF_DATASET_EMPTY;
    if (L > 0 && leftcol[i]->data == nullptr) return GDF_DATASET_EMPTY;
    if (rightcol[i]->dtype != leftcol[i]->dtype) return GDF_JOIN_DTYPE_MISMATCH;
    if (L != leftcol[i]->size) return GDF_COLUMN_SIZE_MISMATCH;
    if (R != rightcol[i]->size) return GDF_COLUMN_SIZE_MISMATCH;
  }
  if (ctx->flag_method == GDF_SORT) return num_cols == 1 ? GDF_UNSUPPORTED_METHOD : GDF_JOIN_TOO_MANY_COLUMNS;
  if (ctx->flag_method != GDF_HASH) return GDF_UNSUPPORTED_METHOD;
  B200_REQUIRE(num_cols <= kMaxCols, GDF_JOIN_TOO_MANY_COLUMNS);
  for (int i = 0; i < num_cols; ++i) B200_REQUIRE(hashable_dtype(leftcol[i]->dtype), GDF_UNSUPPORTED_DTYPE);

  gdf_nvtx_range_push("LIBGDF_JOIN", GDF_CYAN);
  // build on the right table; INNER builds on the smaller side (ref joining.h:59-67)
  const bool flip = (kind == JOIN_INNER) && (R > L);
  gdf_column** probe_cols = flip ? rightcol : leftcol;
  gdf_column** build_cols = flip ? leftcol : rightcol;
  gdf_error e = GDF_SUCCESS;
  bool handled = false;
  if (num_cols == 1) e = partitioned_join(kind, probe_cols[0], build_cols[0], flip, out_l, out_r, &handled);
  else if (num_cols == 2)  // (4/8-byte integer, 4-byte integer) composite keys, e.g. C5's (int64,int32)
    e = partitioned_join(kind, probe_cols[0], build_cols[0], flip, out_l, out_r, &handled, nullptr, nullptr, probe_cols[1],
                         build_cols[1]);
  if (e == GDF_SUCCESS && !handled) {
    TableView probe, build;
    make_view(probe, probe_cols, num_cols);
    make_view(build, build_cols, num_cols);
    e = generic_hash_join(kind, probe, build, flip, out_l, out_r);
  }
  gdf_nvtx_range_pop();
  return e;
}

// ---------------------------------------------------------------------------------------------
// result_cols materialisation (ref joining.cu:376-479, gdf_table.cuh:168-214,1215-1297)
// ---------------------------------------------------------------------------------------------
struct GatherCols {
  const void* in[kMaxCols];
  const gdf_valid_type* in_valid[kMaxCols];
  void* out[kMaxCols];
  gdf_valid_type* out_valid[kMaxCols];
  unsigned char width[kMaxCols];
  int ncols;
};

// One thread per output row, four rows in flight per thread: index reads and value stores of a warp are coalesced, the
// four random value loads of a thread are all issued before the first store, and the validity bits of a warp's 32 rows
// come from one ballot - lanes 0..3 each own one whole output byte, so no atomics.  (Round 1: one thread per 8 rows,
// every store waiting for its own load, stores of a warp 64 bytes apart: 3.8 ms per 2e8-row int64 column.)
static __device__ __forceinline__ uint64_t gather_load(const void* col, int width, size_t i) {
  switch (width) {
    case 1: return static_cast<const uint8_t*>(col)[i];
    case 2: return static_cast<const uint16_t*>(col)[i];
    case 4: return static_cast<const uint32_t*>(col)[i];
    default: return static_cast<const uint64_t*>(col)[i];
  }
}
static __device__ __forceinline__ void gather_store(void* col, int width, size_t i, uint64_t v) {
  switch (width) {
    case 1: static_cast<uint8_t*>(col)[i] = (uint8_t)v; break;
    case 2: static_cast<uint16_t*>(col)[i] = (uint16_t)v; break;
    case 4: static_cast<uint32_t*>(col)[i] = (uint32_t)v; break;
    default: static_cast<uint64_t*>(col)[i] = v; break;
  }
}

__global__ void __launch_bounds__(kThreads) gather_kernel(GatherCols g, const int32_t* __restrict__ idx, size_t n,
                                                          size_t in_rows, bool merge_valid) {
  constexpr int U = 4;
  const size_t nbytes = (n + 7) / 8;
  const unsigned lane = threadIdx.x & 31u;
  const size_t warps = (size_t)gridDim.x * (kThreads / 32);
  const size_t groups = (n + 31) / 32;  // a group = 32 consecutive rows = 4 validity bytes
  for (size_t g0 = (size_t)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5); g0 < groups; g0 += warps * U) {
    int32_t src[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const size_t grp = g0 + (size_t)u * warps, i = grp * 32 + lane;
      src[u] = (grp < groups && i < n) ? idx[i] : -1;
      if (src[u] >= 0 && (size_t)src[u] >= in_rows) src[u] = -1;  // range-checked gather
    }
#pragma unroll 1
    for (int c = 0; c < g.ncols; ++c) {
      const int width = g.width[c];
      uint64_t val[U];
      bool ok[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        val[u] = 0;
        ok[u] = false;
        if (src[u] >= 0) {
          val[u] = gather_load(g.in[c], width, (size_t)src[u]);
          ok[u] = bit_valid(g.in_valid[c], (size_t)src[u]);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const size_t grp = g0 + (size_t)u * warps;
        // rows without a partner keep whatever the buffer held, their bit stays 0
        if (src[u] >= 0) gather_store(g.out[c], width, grp * 32 + lane, val[u]);
        const unsigned bits = __ballot_sync(0xffffffffu, ok[u]);
        const size_t byte = grp * 4 + lane;
        if (g.out_valid[c] && lane < 4 && grp < groups && byte < nbytes) {
          const gdf_valid_type v = (gdf_valid_type)((bits >> (8 * lane)) & 0xffu);
          g.out_valid[c][byte] = merge_valid ? (gdf_valid_type)(g.out_valid[c][byte] | v) : v;
        }
      }
    }
  }
}

gdf_error gather_columns(gdf_column* const* in_cols, gdf_column* const* out_cols, int ncols,
                         const gdf_column* indices, bool merge_valid) {
  const size_t n = indices->size;
  if (n == 0 || ncols == 0) return GDF_SUCCESS;
  B200_TIMED("join_gather");
  for (int base = 0; base < ncols; base += kMaxCols) {
    GatherCols g;
    g.ncols = ncols - base < kMaxCols ? ncols - base : kMaxCols;
    for (int c = 0; c < g.ncols; ++c) {
      const gdf_column* ic = in_cols[base + c];
      gdf_column* oc = out_cols[base + c];
      const int w = dtype_width(ic->dtype);
      B200_REQUIRE(w != 0, GDF_UNSUPPORTED_DTYPE);
      g.in[c] = ic->data;
      g.in_valid[c] = ic->valid;
      g.out[c] = oc->data;
      g.out_valid[c] = oc->valid;
      g.width[c] = (unsigned char)w;
    }
    gather_kernel<<<grid_for((n + 3) / 4), kThreads>>>(g, static_cast<const int32_t*>(indices->data), n,
                                                      in_cols[base]->size, merge_valid);
    B200_CHECK_LAST();
  }
  return GDF_SUCCESS;
}

gdf_error alloc_result_col(gdf_column* out, size_t n, gdf_dtype dtype) {
  gdf_column_view(out, nullptr, nullptr, n, dtype);
  const int w = dtype_width(dtype);
  B200_REQUIRE(w != 0, GDF_UNSUPPORTED_DTYPE);
  B200_RMM_TRY(output_alloc(&out->data, (n ? n : 1) * (size_t)w));
  if (output_alloc((void**)&out->valid, valid_bytes(n) ? valid_bytes(n) : 1) != RMM_SUCCESS) {
    rmmFree(out->data, 0);
    out->data = nullptr;
    return GDF_MEMORYMANAGER_ERROR;
  }
  B200_CUDA_TRY(cudaMemsetAsync(out->valid, 0, valid_bytes(n), 0));
  return GDF_SUCCESS;
}

gdf_error construct_join_output(int kind, gdf_column** left_cols, int num_left_cols, int left_join_cols[],
                                gdf_column** right_cols, int num_right_cols, int right_join_cols[],
                                int num_cols_to_join, int result_num_cols, gdf_column** result_cols,
                                gdf_column* left_indices, gdf_column* right_indices) {
  gdf_nvtx_range_push("LIBGDF_JOIN_OUTPUT", GDF_CYAN);
  auto is_join_col = [&](const int* set, int idx) {
    for (int i = 0; i < num_cols_to_join; ++i)
      if (set[i] == idx) return true;
    return false;
  };
  std::vector<gdf_column*> lnon, rnon;
  for (int i = 0; i < num_left_cols; ++i)
    if (!is_join_col(left_join_cols, i)) lnon.push_back(left_cols[i]);
  for (int i = 0; i < num_right_cols; ++i)
    if (!is_join_col(right_join_cols, i)) rnon.push_back(right_cols[i]);
  const int nl = (int)lnon.size(), nr = (int)rnon.size();
  // layout [left non-key.., key.., right non-key..] (ref joining.cu:412-439): the caller's counts must describe
  // exactly that - duplicate join indices or a wrong result_num_cols would otherwise index past the lists
  if (result_num_cols != nl + num_cols_to_join + nr) {
    gdf_nvtx_range_pop();
    return GDF_INVALID_API_CALL;
  }
  const size_t n = left_indices->size;
  const int left_end = nl;
  const int right_begin = nl + num_cols_to_join;
  for (int i = 0; i < result_num_cols; ++i) {
    if (result_cols[i] == nullptr) {
      gdf_nvtx_range_pop();
      return GDF_DATASET_EMPTY;
    }
    result_cols[i]->data = nullptr;
    result_cols[i]->valid = nullptr;
  }
  gdf_error e = GDF_SUCCESS;
  for (int i = 0; i < left_end && e == GDF_SUCCESS; ++i) e = alloc_result_col(result_cols[i], n, lnon[i]->dtype);
  for (int i = right_begin; i < result_num_cols && e == GDF_SUCCESS; ++i)
    e = alloc_result_col(result_cols[i], n, rnon[i - right_begin]->dtype);
  gdf_column* ljoin[kMaxCols];
  gdf_column* rjoin[kMaxCols];
  for (int j = 0; j < num_cols_to_join && e == GDF_SUCCESS; ++j) {
    ljoin[j] = left_cols[left_join_cols[j]];
    rjoin[j] = right_cols[right_join_cols[j]];
    e = alloc_result_col(result_cols[left_end + j], n, ljoin[j]->dtype);
  }
  if (e == GDF_SUCCESS && nl) e = gather_columns(lnon.data(), result_cols, nl, left_indices, false);
  if (e == GDF_SUCCESS && nr) e = gather_columns(rnon.data(), result_cols + right_begin, nr, right_indices, false);
  if (e == GDF_SUCCESS && num_cols_to_join) {
    // key columns: FULL first takes the right side's keys, then the left side overwrites where it has a row
    if (kind == JOIN_FULL) e = gather_columns(rjoin, result_cols + left_end, num_cols_to_join, right_indices, false);
    if (e == GDF_SUCCESS)
      e = gather_columns(ljoin, result_cols + left_end, num_cols_to_join, left_indices, kind == JOIN_FULL);
  }
  if (e != GDF_SUCCESS) {  // hand back whatever was allocated before the failure
    for (int i = 0; i < result_num_cols; ++i) {
      rmmFree(result_cols[i]->data, 0);
      rmmFree(result_cols[i]->valid, 0);
      result_cols[i]->data = nullptr;
      result_cols[i]->valid = nullptr;
    }
  }
  gdf_nvtx_range_pop();
  return e;
}

gdf_error join_call_compute_df(int kind, gdf_column** left_cols, int num_left_cols, int left_join_cols[],
                               gdf_column** right_cols, int num_right_cols, int right_join_cols[],
                               int num_cols_to_join, int result_num_cols, gdf_column** result_cols,
                               gdf_column* left_indices, gdf_column* right_indices, gdf_context* ctx) {
  if (left_cols == nullptr || right_cols == nullptr) return GDF_DATASET_EMPTY;
  if (num_cols_to_join == 0) return GDF_SUCCESS;
  if (left_join_cols == nullptr || right_join_cols == nullptr) return GDF_DATASET_EMPTY;
  const bool compute_df = result_cols != nullptr;
  if ((left_indices == nullptr || right_indices == nullptr) && !compute_df) return GDF_DATASET_EMPTY;
  if (ctx == nullptr) return GDF_INVALID_API_CALL;
  B200_REQUIRE(num_cols_to_join <= kMaxCols, GDF_JOIN_TOO_MANY_COLUMNS);
  gdf_column tmp_l, tmp_r;
  gdf_column_view(&tmp_l, nullptr, nullptr, 0, N_GDF_TYPES);
  gdf_column_view(&tmp_r, nullptr, nullptr, 0, N_GDF_TYPES);
  gdf_column* lo = left_indices ? left_indices : &tmp_l;
  gdf_column* ro = right_indices ? right_indices : &tmp_r;
  gdf_column* lj[kMaxCols];
  gdf_column* rj[kMaxCols];
  for (int i = 0; i < num_cols_to_join; ++i) {
    B200_REQUIRE(left_join_cols[i] >= 0 && left_join_cols[i] < num_left_cols, GDF_INVALID_API_CALL);
    B200_REQUIRE(right_join_cols[i] >= 0 && right_join_cols[i] < num_right_cols, GDF_INVALID_API_CALL);
    lj[i] = left_cols[left_join_cols[i]];
    rj[i] = right_cols[right_join_cols[i]];
  }
  gdf_error e = join_call(kind, num_cols_to_join, lj, rj, lo, ro, ctx);
  if (e == GDF_SUCCESS && compute_df)
    e = construct_join_output(kind, left_cols, num_left_cols, left_join_cols, right_cols, num_right_cols,
                              right_join_cols, num_cols_to_join, result_num_cols, result_cols, lo, ro);
  if (!left_indices && tmp_l.data) rmmFree(tmp_l.data, 0);
  if (!right_indices && tmp_r.data) rmmFree(tmp_r.data, 0);
  return e;
}

}  // namespace
}  // namespace b200

using namespace b200;

#define B200_JOIN(NAME, KIND)                                                                                  \
  extern "C" gdf_error gdf_##NAME##_join(gdf_column** left_cols, int num_left_cols, int left_join_cols[],      \
                                         gdf_column** right_cols, int num_right_cols, int right_join_cols[],   \
                                         int num_cols_to_join, int result_num_cols, gdf_column** result_cols,  \
                                         gdf_column* left_indices, gdf_column* right_indices,                  \
                                         gdf_context* join_context) {                                          \
    return join_call_compute_df(KIND, left_cols, num_left_cols, left_join_cols, right_cols, num_right_cols,    \
                                right_join_cols, num_cols_to_join, result_num_cols, result_cols, left_indices, \
                                right_indices, join_context);                                                  \
  }
B200_JOIN(inner, JOIN_INNER)
B200_JOIN(left, JOIN_LEFT)
B200_JOIN(full, JOIN_FULL)

// ---- multi-GPU layer entry points (include/gdf_b200_ext.h); the reference has no counterpart ----
extern "C" gdf_error gdfx_remap_indices(gdf_column* indices, const int32_t* payload, size_t payload_rows);

extern "C" gdf_error gdfx_partition_pairs(gdf_column* key, int32_t id_base, int num_partitions, void* out_keys,
                                          int32_t* out_ids, unsigned long long* partition_offsets) {
  B200_REQUIRE(key != nullptr && out_keys != nullptr && out_ids != nullptr && partition_offsets != nullptr,
               GDF_DATASET_EMPTY);
  B200_REQUIRE(key->valid == nullptr, GDF_VALIDITY_UNSUPPORTED);
  B200_REQUIRE(num_partitions >= 1, GDF_INVALID_API_CALL);
  B200_REQUIRE(key->size < 0x7fffffffull, GDF_COLUMN_SIZE_TOO_BIG);
  if (key->size == 0) {
    for (int p = 0; p <= num_partitions; ++p) partition_offsets[p] = 0;
    return GDF_SUCCESS;
  }
  return partition_pairs(key, id_base, (unsigned)num_partitions, out_keys, out_ids, partition_offsets);
}

// Join of exchanged {key, global row id} pairs: like gdf_inner_join / gdf_left_join on ONE key column, but the
// output columns hold the rows' ids (payload) instead of their positions.  kind: 0 inner, 1 left.
extern "C" gdf_error gdfx_join_pairs(int kind, gdf_column* left_key, const int32_t* left_ids, gdf_column* right_key,
                                     const int32_t* right_ids, gdf_column* out_l, gdf_column* out_r) {
  B200_REQUIRE(left_key && right_key && out_l && out_r && left_ids && right_ids, GDF_DATASET_EMPTY);
  B200_REQUIRE(kind == JOIN_INNER || kind == JOIN_LEFT, GDF_UNSUPPORTED_JOIN_TYPE);
  B200_REQUIRE(left_key->dtype == right_key->dtype, GDF_JOIN_DTYPE_MISMATCH);
  B200_REQUIRE(!left_key->valid && !right_key->valid, GDF_VALIDITY_UNSUPPORTED);
  const size_t L = left_key->size, R = right_key->size;
  B200_REQUIRE(L < 0x7fffffffull && R < 0x7fffffffull, GDF_COLUMN_SIZE_TOO_BIG);
  gdf_column_view(out_l, nullptr, nullptr, 0, N_GDF_TYPES);
  gdf_column_view(out_r, nullptr, nullptr, 0, N_GDF_TYPES);
  if (L == 0 || (kind == JOIN_INNER && R == 0)) return GDF_SUCCESS;
  const bool flip = kind == JOIN_INNER && R > L;
  gdf_column* probe = flip ? right_key : left_key;
  gdf_column* build = flip ? left_key : right_key;
  bool handled = false;
  gdf_error e = partitioned_join(kind, probe, build, flip, out_l, out_r, &handled, flip ? right_ids : left_ids,
                                 flip ? left_ids : right_ids);
  if (e != GDF_SUCCESS || handled) return e;
  // key types / inputs the partitioned path does not take: generic join on positions, then map to ids
  gdf_column* pc[1] = {probe};
  gdf_column* bc[1] = {build};
  TableView pv, bv;
  make_view(pv, pc, 1);
  make_view(bv, bc, 1);
  e = generic_hash_join(kind, pv, bv, flip, out_l, out_r);
  if (e != GDF_SUCCESS) return e;
  e = gdfx_remap_indices(out_l, left_ids, L);
  if (e == GDF_SUCCESS) e = gdfx_remap_indices(out_r, right_ids, R);
  return e;
}

// ---- fused partition + exchange over peer memory ----
extern "C" gdf_error gdfx_partition_count(gdf_column* key, int num_partitions, unsigned long long* counts) {
  B200_REQUIRE(key != nullptr && counts != nullptr, GDF_DATASET_EMPTY);
  B200_REQUIRE(key->valid == nullptr, GDF_VALIDITY_UNSUPPORTED);
  B200_REQUIRE(num_partitions >= 1, GDF_INVALID_API_CALL);
  if (key->size == 0) {
    for (int p = 0; p < num_partitions; ++p) counts[p] = 0;
    return GDF_SUCCESS;
  }
  return partition_count(key, (unsigned)num_partitions, counts);
}

extern "C" gdf_error gdfx_partition_scatter_peer(gdf_column* key, int32_t id_base, int num_partitions,
                                                 void* const* dst_keys, int32_t* const* dst_ids,
                                                 const unsigned long long* dst_offsets) {
  B200_REQUIRE(key && dst_keys && dst_ids && dst_offsets, GDF_DATASET_EMPTY);
  B200_REQUIRE(key->valid == nullptr, GDF_VALIDITY_UNSUPPORTED);
  B200_REQUIRE(key->size < 0x7fffffffull, GDF_COLUMN_SIZE_TOO_BIG);
  if (key->size == 0) return GDF_SUCCESS;
  return partition_scatter_peer(key, id_base, (unsigned)num_partitions, dst_keys, dst_ids, dst_offsets);
}

// ---- fused exchange, one partition pass per side (include/gdf_b200_ext.h) ----
extern "C" gdf_error gdfx_xjoin_count(gdf_column* key, int ranks, int nlocal, unsigned long long* counts, unsigned* hi_or) {
  B200_REQUIRE(key != nullptr && counts != nullptr && hi_or != nullptr, GDF_DATASET_EMPTY);
  B200_REQUIRE(key->valid == nullptr, GDF_VALIDITY_UNSUPPORTED);
  B200_REQUIRE(ranks >= 1 && nlocal >= 1, GDF_INVALID_API_CALL);
  *hi_or = 0;
  if (key->size == 0) {
    for (int p = 0; p < ranks * nlocal; ++p) counts[p] = 0;
    return GDF_SUCCESS;
  }
  return xjoin_count(key, (unsigned)ranks, (unsigned)nlocal, counts, hi_or);
}
extern "C" gdf_error gdfx_xjoin_scatter(gdf_column* key, int32_t id_base, int ranks, int nlocal, void* const* dst_pairs,
                                        const unsigned long long* dst_offsets, const unsigned long long* counts) {
  B200_REQUIRE(key && dst_pairs && dst_offsets && counts, GDF_DATASET_EMPTY);
  B200_REQUIRE(key->valid == nullptr, GDF_VALIDITY_UNSUPPORTED);
  B200_REQUIRE(key->size < 0x7fffffffull, GDF_COLUMN_SIZE_TOO_BIG);
  B200_REQUIRE(ranks >= 1 && nlocal >= 1, GDF_INVALID_API_CALL);
  if (key->size == 0) return GDF_SUCCESS;
  return xjoin_scatter(key, id_base, (unsigned)ranks, (unsigned)nlocal, dst_pairs, dst_offsets, counts, nullptr, 0);
}
// the local join in two stages: tables are filled on a private stream while the probe side is still being exchanged
extern "C" gdf_error gdfx_xjoin_build(const void* build_pairs, const unsigned long long* build_counts, int nlocal, int overlap,
                                      void** handle) {
  B200_REQUIRE(build_pairs && build_counts && handle, GDF_DATASET_EMPTY);
  B200_REQUIRE(nlocal >= 1, GDF_INVALID_API_CALL);
  return xjoin_build(build_pairs, build_counts, (unsigned)nlocal, overlap != 0, handle);
}
extern "C" gdf_error gdfx_xjoin_probe(void* handle, const void* probe_pairs, const unsigned long long* probe_counts,
                                      gdf_column* out_l, gdf_column* out_r) {
  B200_REQUIRE(handle && probe_pairs && probe_counts && out_l && out_r, GDF_DATASET_EMPTY);
  return xjoin_probe(handle, probe_pairs, probe_counts, out_l, out_r);
}
// asynchronous variants: no host synchronisation anywhere (include/gdf_b200_ext.h)
extern "C" gdf_error gdfx_xjoin_count_dev(gdf_column* key, int ranks, int nlocal, unsigned long long* d_counts) {
  B200_REQUIRE(key != nullptr && d_counts != nullptr, GDF_DATASET_EMPTY);
  B200_REQUIRE(key->valid == nullptr, GDF_VALIDITY_UNSUPPORTED);
  B200_REQUIRE(ranks >= 1 && nlocal >= 1, GDF_INVALID_API_CALL);
  return xjoin_count_dev(key, (unsigned)ranks, (unsigned)nlocal, d_counts);
}
extern "C" gdf_error gdfx_xjoin_plan_dev(const unsigned long long* d_all, int ranks, int nlocal, int rank, unsigned long long cap_build,
                                         unsigned long long cap_probe, unsigned long long* d_off_build,
                                         unsigned long long* d_off_probe, int* d_status) {
  B200_REQUIRE(d_all && d_off_build && d_off_probe && d_status, GDF_DATASET_EMPTY);
  B200_REQUIRE(ranks >= 1 && nlocal >= 1 && rank >= 0, GDF_INVALID_API_CALL);
  return xjoin_plan_dev(d_all, (unsigned)ranks, (unsigned)nlocal, (unsigned)rank, cap_build, cap_probe, d_off_build, d_off_probe,
                        d_status);
}
extern "C" gdf_error gdfx_xjoin_scatter_dev(gdf_column* key, int32_t id_base, int ranks, int nlocal, void* const* dst_pairs,
                                            const unsigned long long* d_offsets, const unsigned long long* d_counts,
                                            const int* d_status, int ctas_per_sm) {
  B200_REQUIRE(key && dst_pairs && d_offsets && d_counts && d_status, GDF_DATASET_EMPTY);
  B200_REQUIRE(key->valid == nullptr, GDF_VALIDITY_UNSUPPORTED);
  B200_REQUIRE(key->size < 0x7fffffffull, GDF_COLUMN_SIZE_TOO_BIG);
  B200_REQUIRE(ranks >= 1 && nlocal >= 1, GDF_INVALID_API_CALL);
  if (key->size == 0) return GDF_SUCCESS;
  return xjoin_scatter(key, id_base, (unsigned)ranks, (unsigned)nlocal, dst_pairs, d_offsets, d_counts, d_status, ctas_per_sm);
}
extern "C" gdf_error gdfx_xjoin_local(const void* probe_pairs, const unsigned long long* probe_counts, const void* build_pairs,
                                      const unsigned long long* build_counts, int nlocal, gdf_column* out_l, gdf_column* out_r) {
  B200_REQUIRE(probe_pairs && probe_counts && build_pairs && build_counts && out_l && out_r, GDF_DATASET_EMPTY);
  B200_REQUIRE(nlocal >= 1, GDF_INVALID_API_CALL);
  return xjoin_local(probe_pairs, probe_counts, build_pairs, build_counts, (unsigned)nlocal, out_l, out_r);
}

// Receive buffers that other ranks (processes) can map: plain cudaMalloc + legacy CUDA IPC handle (64 bytes).
extern "C" gdf_error gdfx_peer_alloc(void** ptr, size_t bytes, char* handle64) {
  B200_REQUIRE(ptr != nullptr && handle64 != nullptr, GDF_DATASET_EMPTY);
  B200_CUDA_TRY(cudaMalloc(ptr, bytes ? bytes : 1));
  cudaIpcMemHandle_t h;
  if (cudaIpcGetMemHandle(&h, *ptr) != cudaSuccess) {
    cudaFree(*ptr);
    *ptr = nullptr;
    return GDF_CUDA_ERROR;
  }
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
  memcpy(handle64, &h, 64);
  return GDF_SUCCESS;
}
extern "C" gdf_error gdfx_peer_open(const char* handle64, void** ptr) {
  B200_REQUIRE(ptr != nullptr && handle64 != nullptr, GDF_DATASET_EMPTY);
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  B200_CUDA_TRY(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return GDF_SUCCESS;
}
extern "C" gdf_error gdfx_peer_close(void* ptr) {
  B200_CUDA_TRY(cudaIpcCloseMemHandle(ptr));
  return GDF_SUCCESS;
}
extern "C" gdf_error gdfx_peer_free(void* ptr) {
  B200_CUDA_TRY(cudaFree(ptr));
  return GDF_SUCCESS;
}
