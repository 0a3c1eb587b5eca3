// Hash group-by with one aggregation column: gdf_group_by_{sum,min,max,count,avg}, GDF_HASH method.
//
// Reference behaviour followed (file:line in /root/reference/libgdf/src):
//   argument / mask / empty-input rules      sqls_ops.cu:1085-1128, groupby/groupby.cuh:218-239
//   aggregation runs in the INPUT column's C type (int8 sums wrap in int8); COUNT runs in the OUTPUT
//   column's type; AVG = SUM (input type) / COUNT cast to the output type, integer division for
//   integer outputs                          groupby/groupby.cuh:88-190,308-386
//   group keys are copied from the first-seen row of each group into out_col_values[i]->data, the
//   aggregate into out_col_agg->data, and every output size is set to the group count; output
//   order is unspecified unless flag_sort_result (AVG always sorts)
//                                            groupby/hash/groupby_compute_api.h:143-225
//
// B200 design.  The reference sizes its table at 2*N slots whatever the number of groups (32 GB for
// 1e9 rows), initialises it, hammers it with CAS loops and then streams all of it again to extract.
// Here the table is sized for L2, not for N:
//   * level 1 uses at most 2^22 slots (<= 64 MB, L2-resident on B200's 126 MB L2); rows whose key
//     cannot be placed within a bounded probe sequence are SPILLED as 4-byte row ids;
//   * level 2 (only if something spilled) regroups the spilled rows in a table of 2x their count,
//     which cannot fail.  A key lives in exactly one level (slots are never freed, so a key that was
//     placed is always found again within the probe bound), so the two result sets are disjoint;
//   * values are folded with native L2 atomics (red.add / atom.min / atom.max), never CAS loops for
//     integers; the single-key fast path (groupby_fast.cuh) puts a per-CTA shared-memory cache in front of
//     the table so a Zipf-hot key costs a shared-memory atomic, and keeps the table's keys in 32-byte
//     buckets so a cold key costs one L2 sector read and one red;
//   * extraction compacts only the small table, with one cursor atomic per warp.
#include <cstdlib>
#include <limits>
#include <type_traits>
#include <vector>

#include "sort.cuh"
#include "table.cuh"

namespace b200 {
namespace {

constexpr int kThreads = 256;
constexpr unsigned kLevel1Slots = 1u << 22;
constexpr unsigned kProbeLimitL1 = 64;

enum Op { OP_SUM = 0, OP_MIN = 1, OP_MAX = 2, OP_COUNT = 3, OP_AVG = 4 };

// ---- accumulator types: int8/int16/int32 fold in int32, int64 in int64 (wrap == truncate later) ----
template <typename IT> struct AccOf { using type = IT; };
template <> struct AccOf<int8_t> { using type = int32_t; };
template <> struct AccOf<int16_t> { using type = int32_t; };

__device__ __forceinline__ void fold(int32_t* p, int32_t v, int op) {
  if (op == OP_MIN) atomicMin(p, v);
  else if (op == OP_MAX) atomicMax(p, v);
  else atomicAdd(p, v);
}
__device__ __forceinline__ void fold(int64_t* p, int64_t v, int op) {
  if (op == OP_MIN) atomicMin(reinterpret_cast<long long*>(p), (long long)v);
  else if (op == OP_MAX) atomicMax(reinterpret_cast<long long*>(p), (long long)v);
  else atomicAdd(reinterpret_cast<unsigned long long*>(p), (unsigned long long)v);
}
__device__ __forceinline__ void fold(float* p, float v, int op) {
  if (op == OP_SUM || op == OP_COUNT || op == OP_AVG) { atomicAdd(p, v); return; }
  int* ip = reinterpret_cast<int*>(p);
  int old = *ip;
  while (true) {  // same selection rule as the reference functors (aggregation_operations.cuh:35-53)
    const float cur = __int_as_float(old);
    const bool take = (op == OP_MIN) ? (v < cur) : (v > cur);
    if (!take) break;
    const int prev = atomicCAS(ip, old, __float_as_int(v));
    if (prev == old) break;
    old = prev;
  }
}
__device__ __forceinline__ void fold(double* p, double v, int op) {
  if (op == OP_SUM || op == OP_COUNT || op == OP_AVG) { atomicAdd(p, v); return; }
  unsigned long long* ip = reinterpret_cast<unsigned long long*>(p);
  unsigned long long old = *ip;
  while (true) {
    const double cur = __longlong_as_double((long long)old);
    const bool take = (op == OP_MIN) ? (v < cur) : (v > cur);
    if (!take) break;
    const unsigned long long prev = atomicCAS(ip, old, (unsigned long long)__double_as_longlong(v));
    if (prev == old) break;
    old = prev;
  }
}

struct RowSource {  // implicit [0, n) or an explicit list of row ids (level 2)
  const int32_t* idx;
  size_t n;
  __device__ __forceinline__ size_t row(size_t i) const { return idx ? (size_t)idx[i] : i; }
};

struct Spill {
  int32_t* rows;           // capacity >= number of rows of this level
  unsigned long long* count;
};

// ------------------------------------------------------------------------------------------
// generic path: any key columns; slot key = id of the first row of the group
// ------------------------------------------------------------------------------------------
template <typename A>
__global__ void init_generic_kernel(int32_t* slot_row, A* acc, unsigned long long* cnt, size_t slots, A identity) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < slots; i += stride) {
    slot_row[i] = -1;
    acc[i] = identity;
    if (cnt) cnt[i] = 0;
  }
}

template <typename IT, typename A>
__global__ void __launch_bounds__(kThreads)
build_generic_kernel(TableView keys, const IT* __restrict__ values, RowSource src, int op,
                     int32_t* __restrict__ slot_row, A* __restrict__ acc, unsigned long long* __restrict__ cnt,
                     unsigned mask, unsigned probe_limit, Spill spill) {
  const size_t stride = (size_t)gridDim.x * kThreads;
  for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < src.n; i += stride) {
    const size_t r = src.row(i);
    unsigned s = row_hash<false>(keys, r) & mask;
    bool placed = false;
    for (unsigned probe = 0; probe < probe_limit; ++probe, s = (s + 1) & mask) {
      int32_t k = slot_row[s];
      if (k == -1) {
        const int32_t prev = atomicCAS(&slot_row[s], -1, (int32_t)r);
        k = (prev == -1) ? (int32_t)r : prev;
      }
      if (k == (int32_t)r || rows_equal(keys, r, keys, (size_t)k)) {
        if (op == OP_COUNT) fold(&acc[s], (A)1, OP_SUM);
        else fold(&acc[s], (A)values[r], op);
        if (cnt) atomicAdd(&cnt[s], 1ull);
        placed = true;
        break;
      }
    }
    if (!placed) {
      const unsigned long long at = atomicAdd(spill.count, 1ull);
      spill.rows[at] = (int32_t)r;
    }
  }
}

struct KeyOut {  // destination of the group keys: one output column per key column
  void* out[kMaxCols];
};

// AVG finalisation (ref groupby.cuh:308-328): sum (typed as the input column, narrow ints already
// wrapped) divided by static_cast<avg_type>(count), result converted to avg_type.
template <typename ST>
__device__ __forceinline__ void store_avg(void* out, int out_dtype, size_t at, ST sum, unsigned long long count) {
  switch (out_dtype) {
    case GDF_INT8: static_cast<int8_t*>(out)[at] = (int8_t)(sum / static_cast<int8_t>(count)); break;
    case GDF_INT16: static_cast<int16_t*>(out)[at] = (int16_t)(sum / static_cast<int16_t>(count)); break;
    case GDF_INT32: static_cast<int32_t*>(out)[at] = (int32_t)(sum / static_cast<int32_t>(count)); break;
    case GDF_INT64: static_cast<int64_t*>(out)[at] = (int64_t)(sum / static_cast<int64_t>(count)); break;
    case GDF_FLOAT32: static_cast<float*>(out)[at] = (float)(sum / static_cast<float>(count)); break;
    default: static_cast<double*>(out)[at] = (double)(sum / static_cast<double>(count)); break;
  }
}

// Write an accumulator into the aggregate output, typed as OT (the column's C type).
template <typename OT, typename A>
__device__ __forceinline__ void store_acc(void* out, size_t at, A v) { static_cast<OT*>(out)[at] = (OT)v; }

static __device__ __forceinline__ size_t claim_output(bool have, unsigned long long* cursor) {
  // one cursor atomic per warp
  const unsigned m = __ballot_sync(0xffffffffu, have);
  if (m == 0) return 0;
  unsigned long long base = 0;
  const int leader = __ffs(m) - 1;
  if ((int)lane_id() == leader) base = atomicAdd(cursor, (unsigned long long)__popc(m));
  base = __shfl_sync(0xffffffffu, base, leader);
  return (size_t)(base + __popc(m & lanemask_lt()));
}

static __device__ __forceinline__ void copy_key_row(const TableView& keys, const KeyOut& ko, size_t from, size_t to) {
#pragma unroll 1
  for (int c = 0; c < keys.ncols; ++c) {
    switch (dtype_width(keys.dtype[c])) {
      case 1: static_cast<uint8_t*>(ko.out[c])[to] = static_cast<const uint8_t*>(keys.data[c])[from]; break;
      case 2: static_cast<uint16_t*>(ko.out[c])[to] = static_cast<const uint16_t*>(keys.data[c])[from]; break;
      case 4: static_cast<uint32_t*>(ko.out[c])[to] = static_cast<const uint32_t*>(keys.data[c])[from]; break;
      default: static_cast<uint64_t*>(ko.out[c])[to] = static_cast<const uint64_t*>(keys.data[c])[from]; break;
    }
  }
}

template <typename IT, typename OT, typename A>
__global__ void __launch_bounds__(kThreads)
extract_generic_kernel(TableView keys, KeyOut ko, const int32_t* __restrict__ slot_row, const A* __restrict__ acc,
                       const unsigned long long* __restrict__ cnt, size_t slots, int op, void* out_agg,
                       int out_dtype, unsigned long long* cursor) {
  const size_t stride = (size_t)gridDim.x * kThreads;
  const size_t rounds = (slots + stride - 1) / stride;
  for (size_t it = 0; it < rounds; ++it) {
    const size_t s = it * stride + (size_t)blockIdx.x * kThreads + threadIdx.x;
    const int32_t k = s < slots ? slot_row[s] : -1;
    const bool have = k != -1;
    const size_t at = claim_output(have, cursor);
    if (have) {
      copy_key_row(keys, ko, (size_t)k, at);
      if (op == OP_AVG) store_avg<IT>(out_agg, out_dtype, at, (IT)acc[s], cnt[s]);
      else store_acc<OT, A>(out_agg, at, acc[s]);
    }
  }
}

// ------------------------------------------------------------------------------------------
// fast path: ONE integer key column of 4 or 8 bytes, integer aggregate.  Slot = {key bits, acc}.
// The all-ones key pattern doubles as the EMPTY marker; a real key with that value is folded into
// a dedicated side slot (index `slots`).
// ------------------------------------------------------------------------------------------
constexpr unsigned long long kEmptyKey = ~0ull;

// keys[i] = EMPTY, acc[i] = identity, cnt[i] = 0 for the slots + the side slot
__global__ void init_fast_kernel(unsigned long long* keys, int64_t* acc, unsigned long long* cnt, size_t slots_plus_side,
                                 int64_t identity) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < slots_plus_side; i += stride) {
    keys[i] = kEmptyKey;
    acc[i] = identity;
    if (cnt) cnt[i] = 0;
  }
}

// 64-bit wrapping add in SHARED memory built from 32-bit atomics: on B200 a 32-bit shared atomic
// runs ~8x faster than a 64-bit one (profiles/r02_microbench.txt: ~2900 vs ~400 Gop/s), and for
// small non-negative addends the high word is touched only on a carry.
static __device__ __forceinline__ void smem_add64(unsigned long long* acc, int64_t v) {
  unsigned* w = reinterpret_cast<unsigned*>(acc);
  const unsigned lo = (unsigned)(unsigned long long)v, hi = (unsigned)((unsigned long long)v >> 32);
  const unsigned old = atomicAdd(&w[0], lo);
  const unsigned up = hi + (unsigned)((old + lo) < old);
  if (up) atomicAdd(&w[1], up);
}

#include "groupby_fast.cuh"

template <typename KT, typename IT, typename OT>
__global__ void __launch_bounds__(kThreads)
extract_fast_kernel(const unsigned long long* __restrict__ keys, const int64_t* __restrict__ acc,
                    const unsigned long long* __restrict__ cnt, size_t slots, const int* __restrict__ side_used, int op,
                    KT* __restrict__ out_keys, void* out_agg, int out_dtype, unsigned long long* cursor) {
  const size_t total = slots + 1;  // + side slot
  const size_t stride = (size_t)gridDim.x * kThreads;
  const size_t rounds = (total + stride - 1) / stride;
  for (size_t it = 0; it < rounds; ++it) {
    const size_t s = it * stride + (size_t)blockIdx.x * kThreads + threadIdx.x;
    bool have = false;
    unsigned long long key = kEmptyKey;
    if (s < slots) {
      key = keys[s];
      have = key != kEmptyKey;
    } else if (s == slots && *side_used) {
      have = true;  // the side slot holds the real key that equals the EMPTY pattern
    }
    const size_t at = claim_output(have, cursor);
    if (have) {
      out_keys[at] = (KT)key;
      if (op == OP_AVG) store_avg<IT>(out_agg, out_dtype, at, (IT)acc[s], cnt[s]);
      else store_acc<OT, int64_t>(out_agg, at, acc[s]);
    }
  }
}

int grid_for(size_t items) {
  size_t want = (items + kThreads - 1) / kThreads;
  const size_t cap = (size_t)sm_count() * 8;
  return (int)(want < 1 ? 1 : (want < cap ? want : cap));
}

unsigned pow2_at_least(size_t x) {
  unsigned p = 1;
  while ((size_t)p < x && p < (1u << 31)) p <<= 1;
  return p;
}

gdf_error read_count(const unsigned long long* d, unsigned long long* h) {
  unsigned long long* box = static_cast<unsigned long long*>(pinned_mailbox());
  B200_REQUIRE(box != nullptr, GDF_CUDA_ERROR);
  B200_CUDA_TRY(cudaMemcpyAsync(box, d, sizeof(unsigned long long), cudaMemcpyDeviceToHost, 0));
  B200_CUDA_TRY(cudaStreamSynchronize(0));
  *h = *box;
  return GDF_SUCCESS;
}

// ---- generic driver: level 1 (bounded table) + level 2 (spilled rows) ----
template <typename IT, typename OT>
gdf_error groupby_generic(const TableView& keys, const KeyOut& ko, const void* values, int op, void* out_agg,
                          int out_dtype, size_t* out_groups) {
  using A = typename std::conditional<std::is_same<IT, void>::value, OT, IT>::type;  // COUNT: IT=void
  using ACC = typename AccOf<A>::type;
  const size_t n = keys.rows;
  Scratch counters;  // [0] output cursor, [1] spill count
  B200_CUDA_TRY(counters.alloc(2 * sizeof(unsigned long long)));
  B200_CUDA_TRY(cudaMemsetAsync(counters.ptr, 0, 2 * sizeof(unsigned long long), 0));
  unsigned long long* cursor = counters.as<unsigned long long>();
  unsigned long long* spill_count = cursor + 1;

  Scratch spill_a, spill_b;
  RowSource src{nullptr, n};
  for (int level = 0; level < 2; ++level) {
    unsigned slots = pow2_at_least(2 * src.n);
    bool can_spill = false;
    if (level == 0 && slots > kLevel1Slots) {
      slots = kLevel1Slots;
      can_spill = true;
    }
    Scratch t_rows, t_acc, t_cnt;
    B200_CUDA_TRY(t_rows.alloc((size_t)slots * sizeof(int32_t)));
    B200_CUDA_TRY(t_acc.alloc((size_t)slots * sizeof(ACC)));
    if (op == OP_AVG) B200_CUDA_TRY(t_cnt.alloc((size_t)slots * sizeof(unsigned long long)));
    Spill spill{nullptr, spill_count};
    Scratch& my_spill = level == 0 ? spill_a : spill_b;
    if (can_spill) {
      B200_CUDA_TRY(my_spill.alloc(src.n * sizeof(int32_t)));
      spill.rows = my_spill.as<int32_t>();
    }
    ACC identity;
    if (op == OP_MIN) identity = (ACC)std::numeric_limits<A>::max();
    else if (op == OP_MAX) identity = (ACC)std::numeric_limits<A>::lowest();
    else identity = (ACC)0;
    B200_TIMED("groupby_generic_level");
    init_generic_kernel<ACC><<<grid_for(slots), kThreads>>>(t_rows.as<int32_t>(), t_acc.as<ACC>(),
                                                           t_cnt.as<unsigned long long>(), slots, identity);
    B200_CHECK_LAST();
    using VT = typename std::conditional<std::is_same<IT, void>::value, int8_t, IT>::type;
    build_generic_kernel<VT, ACC><<<grid_for(src.n), kThreads>>>(
        keys, static_cast<const VT*>(values), src, op, t_rows.as<int32_t>(), t_acc.as<ACC>(),
        t_cnt.as<unsigned long long>(), slots - 1, can_spill ? kProbeLimitL1 : slots, spill);
    B200_CHECK_LAST();
    extract_generic_kernel<VT, OT, ACC><<<grid_for(slots), kThreads>>>(
        keys, ko, t_rows.as<int32_t>(), t_acc.as<ACC>(), t_cnt.as<unsigned long long>(), slots, op, out_agg,
        out_dtype, cursor);
    B200_CHECK_LAST();
    if (!can_spill) break;
    unsigned long long spilled = 0;
    gdf_error e = read_count(spill_count, &spilled);
    if (e != GDF_SUCCESS) return e;
    if (spilled == 0) break;
    src = RowSource{spill.rows, (size_t)spilled};
  }
  unsigned long long groups = 0;
  gdf_error e = read_count(cursor, &groups);
  if (e != GDF_SUCCESS) return e;
  *out_groups = (size_t)groups;
  return GDF_SUCCESS;
}

// ---- fast driver: one bounded, L2-resident table.  If the keys do not fit (more groups than the
// table can hold within the probe bound) the call reports `overflowed` and the caller reruns the
// input through the generic two-level path. ----
template <typename KT, typename IT, typename OT>
gdf_error groupby_fast(const KT* key_col, size_t n, const void* values, int op, KT* out_keys, void* out_agg,
                       int out_dtype, size_t* out_groups, bool* overflowed) {
  using VT = typename std::conditional<std::is_same<IT, void>::value, int32_t, IT>::type;
  *overflowed = false;
  Scratch misc;
  B200_CUDA_TRY(misc.alloc(4 * sizeof(unsigned long long)));
  B200_CUDA_TRY(cudaMemsetAsync(misc.ptr, 0, 4 * sizeof(unsigned long long), 0));
  unsigned long long* cursor = misc.as<unsigned long long>();
  int* flags = reinterpret_cast<int*>(cursor + 1);
  int64_t identity = 0;
  if (op == OP_MIN) identity = (int64_t)std::numeric_limits<VT>::max();
  if (op == OP_MAX) identity = (int64_t)std::numeric_limits<VT>::lowest();
  unsigned slots = pow2_at_least(2 * n);
  if (slots < kBucketKeys) slots = kBucketKeys;
  const unsigned level1 = (unsigned)lab_knob("B200_GB_SLOTS_LOG2", 0) ? 1u << lab_knob("B200_GB_SLOTS_LOG2", 0) : kLevel1Slots;
  const bool bounded = slots > level1;
  if (bounded) slots = level1;
  Scratch keys, acc, cnt;
  B200_CUDA_TRY(keys.alloc(((size_t)slots + kBucketKeys) * sizeof(unsigned long long)));
  B200_CUDA_TRY(acc.alloc(((size_t)slots + 1) * sizeof(int64_t)));
  if (op == OP_AVG) B200_CUDA_TRY(cnt.alloc(((size_t)slots + 1) * sizeof(unsigned long long)));
  {
    B200_TIMED("groupby_init_table");
    init_fast_kernel<<<grid_for((size_t)slots + 1), kThreads>>>(keys.as<unsigned long long>(), acc.as<int64_t>(),
                                                               cnt.as<unsigned long long>(), (size_t)slots + 1, identity);
  }
  B200_CHECK_LAST();
  const int fold_op = (op == OP_MIN) ? OP_MIN : (op == OP_MAX ? OP_MAX : OP_SUM);
  const bool count_rows = op == OP_COUNT, with_cnt = op == OP_AVG;
  void (*kern)(const KT*, const VT*, size_t, GlobalTable, int*, unsigned long long*, unsigned) = nullptr;
  if (count_rows) kern = build_fast_kernel_v5<KT, VT, OP_SUM, true, false>;
  else if (with_cnt) kern = build_fast_kernel_v5<KT, VT, OP_SUM, false, true>;
  else if (fold_op == OP_MIN) kern = build_fast_kernel_v5<KT, VT, OP_MIN, false, false>;
  else if (fold_op == OP_MAX) kern = build_fast_kernel_v5<KT, VT, OP_MAX, false, false>;
  else kern = build_fast_kernel_v5<KT, VT, OP_SUM, false, false>;
  const int smem = (int)(sizeof(FastCache4) + (with_cnt ? kCacheSlots4 * sizeof(unsigned) : 0));
  B200_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  size_t want = (n + kGrabRows4 - 1) / kGrabRows4;                 // one warp-grab each
  want = (want + kFastThreads4 / 32 - 1) / (kFastThreads4 / 32);    // CTAs needed
  const size_t resident = (size_t)sm_count();                      // one 1024-thread CTA per SM
  const int blocks = (int)(want < resident ? (want ? want : 1) : resident);
  // bounded table: a key that finds no room within kProbeLimitL1 slots reports overflow (generic path reruns it)
  GlobalTable tab{keys.as<unsigned long long>(), acc.as<int64_t>(), cnt.as<unsigned long long>(), slots - 1,
                  bounded ? kProbeLimitL1 / kBucketKeys : slots / kBucketKeys};
  {
    B200_TIMED("groupby_build_fast");
    kern<<<blocks, kFastThreads4, smem>>>(key_col, static_cast<const VT*>(values), n, tab, flags, cursor + 3,
                                          (unsigned)lab_knob("B200_LAB_GB", 0));
  }
  B200_CHECK_LAST();
  if (bounded) {
    int h_flags[2] = {0, 0};
    B200_CUDA_TRY(cudaMemcpy(h_flags, flags, sizeof(h_flags), cudaMemcpyDeviceToHost));
    if (h_flags[1]) {
      *overflowed = true;
      return GDF_SUCCESS;
    }
  }
  {
    B200_TIMED("groupby_extract");
    extract_fast_kernel<KT, VT, OT><<<grid_for((size_t)slots + 1), kThreads>>>(
        keys.as<unsigned long long>(), acc.as<int64_t>(), cnt.as<unsigned long long>(), slots, flags, op, out_keys, out_agg,
        out_dtype, cursor);
  }
  B200_CHECK_LAST();
  unsigned long long groups = 0;
  gdf_error e = read_count(cursor, &groups);
  if (e != GDF_SUCCESS) return e;
  *out_groups = (size_t)groups;
  return GDF_SUCCESS;
}

bool integer_storage(int dtype) {
  switch (dtype) {
    case GDF_INT8: case GDF_INT16: case GDF_INT32: case GDF_INT64:
    case GDF_DATE32: case GDF_DATE64: case GDF_TIMESTAMP: return true;
    default: return false;
  }
}

// dtype -> C type dispatch for the generic driver
template <typename IT>
gdf_error generic_by_out(const TableView& keys, const KeyOut& ko, const void* values, int op, gdf_column* out_agg,
                         size_t* groups) {
  // SUM/MIN/MAX write the accumulator as the INPUT column's type whatever out->dtype says
  // (ref groupby.cuh:55-58); AVG's store switch handles the output type itself.
  return groupby_generic<IT, IT>(keys, ko, values, op, out_agg->data, out_agg->dtype, groups);
}

gdf_error dispatch_generic(const TableView& keys, const KeyOut& ko, gdf_column* col_agg, int op,
                           gdf_column* out_agg, size_t* groups) {
  const int t = (op == OP_COUNT) ? out_agg->dtype : col_agg->dtype;
  const void* v = col_agg->data;
  if (op == OP_COUNT) {
    switch (t) {
      case GDF_INT8: return groupby_generic<void, int8_t>(keys, ko, v, op, out_agg->data, t, groups);
      case GDF_INT16: return groupby_generic<void, int16_t>(keys, ko, v, op, out_agg->data, t, groups);
      case GDF_INT32: case GDF_DATE32: return groupby_generic<void, int32_t>(keys, ko, v, op, out_agg->data, t, groups);
      case GDF_INT64: case GDF_DATE64: case GDF_TIMESTAMP:
        return groupby_generic<void, int64_t>(keys, ko, v, op, out_agg->data, t, groups);
      case GDF_FLOAT32: return groupby_generic<void, float>(keys, ko, v, op, out_agg->data, t, groups);
      case GDF_FLOAT64: return groupby_generic<void, double>(keys, ko, v, op, out_agg->data, t, groups);
      default: return GDF_UNSUPPORTED_DTYPE;
    }
  }
  switch (t) {
    case GDF_INT8: return generic_by_out<int8_t>(keys, ko, v, op, out_agg, groups);
    case GDF_INT16: return generic_by_out<int16_t>(keys, ko, v, op, out_agg, groups);
    case GDF_INT32: case GDF_DATE32: return generic_by_out<int32_t>(keys, ko, v, op, out_agg, groups);
    case GDF_INT64: case GDF_DATE64: case GDF_TIMESTAMP: return generic_by_out<int64_t>(keys, ko, v, op, out_agg, groups);
    case GDF_FLOAT32: return generic_by_out<float>(keys, ko, v, op, out_agg, groups);
    case GDF_FLOAT64: return generic_by_out<double>(keys, ko, v, op, out_agg, groups);
    default: return GDF_UNSUPPORTED_DTYPE;
  }
}

template <typename KT>
gdf_error dispatch_fast_key(gdf_column* key, gdf_column* col_agg, int op, gdf_column* out_key, gdf_column* out_agg,
                            size_t* groups, bool* overflowed, bool* handled) {
  *handled = true;
  const KT* k = static_cast<const KT*>(key->data);
  KT* ok = static_cast<KT*>(out_key->data);
  const size_t n = key->size;
  const void* v = col_agg->data;
  void* oa = out_agg->data;
  const int od = out_agg->dtype;
  if (op == OP_COUNT) {
    if (dtype_width(od) == 4 && integer_storage(od)) return groupby_fast<KT, void, int32_t>(k, n, v, op, ok, oa, od, groups, overflowed);
    if (dtype_width(od) == 8 && integer_storage(od)) return groupby_fast<KT, void, int64_t>(k, n, v, op, ok, oa, od, groups, overflowed);
  } else {
    const int id = col_agg->dtype;
    if (dtype_width(id) == 4 && integer_storage(id)) return groupby_fast<KT, int32_t, int32_t>(k, n, v, op, ok, oa, od, groups, overflowed);
    if (dtype_width(id) == 8 && integer_storage(id)) return groupby_fast<KT, int64_t, int64_t>(k, n, v, op, ok, oa, od, groups, overflowed);
  }
  *handled = false;
  return GDF_SUCCESS;
}

// Reorders the `groups` output rows (key columns + aggregate) into lexicographic key order.
gdf_error sort_groups(int ncols, gdf_column** cols, gdf_column** out_vals, gdf_column* out_agg, int agg_dtype, size_t groups) {
  B200_TIMED("groupby_sort_result");
  std::vector<gdf_column> views((size_t)ncols);
  std::vector<const gdf_column*> ptrs((size_t)ncols);
  for (int c = 0; c < ncols; ++c) {  // the keys were copied bit for bit, so they order as the INPUT column's type
    gdf_column_view(&views[c], out_vals[c]->data, nullptr, groups, cols[c]->dtype);
    ptrs[c] = &views[c];
  }
  Scratch perm;
  B200_CUDA_TRY(perm.alloc(groups * sizeof(uint32_t)));
  gdf_error e = sort_permutation(ptrs.data(), ncols, groups, perm.as<uint32_t>());
  for (int c = 0; c < ncols && e == GDF_SUCCESS; ++c)
    e = permute_in_place(out_vals[c]->data, dtype_width(cols[c]->dtype), groups, perm.as<uint32_t>());
  if (e == GDF_SUCCESS) e = permute_in_place(out_agg->data, dtype_width(agg_dtype), groups, perm.as<uint32_t>());
  return e;
}

// order-preserving unsigned image of a typed value (the same map the radix sort uses, sort.cu)
static __device__ __forceinline__ unsigned long long ordered_bits(int dtype, unsigned long long bits) {
  const int w = dtype_width(dtype);
  const unsigned long long sign = 1ull << (8 * w - 1);
  const unsigned long long all = w == 8 ? ~0ull : ((1ull << (8 * (w & 7))) - 1ull);
  if (dtype == GDF_FLOAT32 || dtype == GDF_FLOAT64) return (bits & sign) ? (~bits & all) : (bits | sign);
  return bits ^ sign;
}

// Sort method: index of one row of every group (the smallest row id), groups given in sorted key order.
// Every input row finds its group by binary search over the sorted group keys and lowers that group's entry.
__global__ void __launch_bounds__(kThreads)
group_rows_kernel(TableView in, TableView sorted_groups, unsigned long long* __restrict__ rep) {
  const size_t stride = (size_t)gridDim.x * kThreads;
  for (size_t r = (size_t)blockIdx.x * kThreads + threadIdx.x; r < in.rows; r += stride) {
    size_t lo = 0, hi = sorted_groups.rows;  // first group whose key is >= row r's key
    while (lo < hi) {
      const size_t mid = (lo + hi) >> 1;
      bool less = false;  // group[mid] < row r ?
#pragma unroll 1
      for (int c = 0; c < in.ncols; ++c) {
        const unsigned long long a = ordered_bits(in.dtype[c], load_bits(sorted_groups, c, mid));
        const unsigned long long b = ordered_bits(in.dtype[c], load_bits(in, c, r));
        if (a != b) {
          less = a < b;
          break;
        }
      }
      if (less) lo = mid + 1;
      else hi = mid;
    }
    if (lo < sorted_groups.rows) atomicMin(&rep[lo], (unsigned long long)r);
  }
}

__global__ void fill_u64_kernel(unsigned long long* p, size_t n, unsigned long long v) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) p[i] = v;
}

template <typename T>
__global__ void store_scalar_kernel(T* p, double v) { p[0] = (T)v; }

gdf_error group_by_hash(int ncols, gdf_column** cols, gdf_column* col_agg, gdf_column** out_vals,
                        gdf_column* out_agg, int op, bool sort_result) {
  if (ncols == 0 || cols == nullptr || col_agg == nullptr) return GDF_DATASET_EMPTY;
  if (out_vals == nullptr || out_agg == nullptr) return GDF_DATASET_EMPTY;
  if (cols[0]->size == 0 || col_agg->size == 0) return GDF_SUCCESS;
  B200_REQUIRE(ncols <= kMaxCols, GDF_JOIN_TOO_MANY_COLUMNS);
  const size_t n = cols[0]->size;
  B200_REQUIRE(n < 0x7fffffffull, GDF_COLUMN_SIZE_TOO_BIG);  // row ids are int32 inside the tables
  for (int c = 0; c < ncols; ++c) {
    B200_REQUIRE(hashable_dtype(cols[c]->dtype), GDF_UNSUPPORTED_DTYPE);
    B200_REQUIRE(cols[c]->size == n, GDF_COLUMN_SIZE_MISMATCH);
    B200_REQUIRE(out_vals[c] != nullptr && out_vals[c]->data != nullptr, GDF_DATASET_EMPTY);
  }
  B200_REQUIRE(col_agg->size == n, GDF_COLUMN_SIZE_MISMATCH);
  B200_REQUIRE(out_agg->data != nullptr, GDF_DATASET_EMPTY);
  if (op == OP_AVG) B200_REQUIRE(dtype_width(out_agg->dtype) != 0 && out_agg->dtype <= GDF_FLOAT64, GDF_UNSUPPORTED_DTYPE);

  size_t groups = 0;
  bool done = false;
  if (ncols == 1 && integer_storage(cols[0]->dtype) && dtype_width(cols[0]->dtype) >= 4) {
    bool handled = false, overflowed = false;
    gdf_error e = dtype_width(cols[0]->dtype) == 8
                      ? dispatch_fast_key<int64_t>(cols[0], col_agg, op, out_vals[0], out_agg, &groups, &overflowed, &handled)
                      : dispatch_fast_key<int32_t>(cols[0], col_agg, op, out_vals[0], out_agg, &groups, &overflowed, &handled);
    if (e != GDF_SUCCESS) return e;
    done = handled && !overflowed;
  }
  if (!done) {
    TableView keys;
    make_view(keys, cols, ncols);
    KeyOut ko;
    for (int c = 0; c < ncols; ++c) ko.out[c] = out_vals[c]->data;
    gdf_error e = dispatch_generic(keys, ko, col_agg, op, out_agg, &groups);
    if (e != GDF_SUCCESS) return e;
  }
  if ((sort_result || op == OP_AVG) && groups > 1) {
    // flag_sort_result: groups in lexicographic key order (ref groupby_compute_api.h:211-222); AVG always comes back
    // sorted in the reference because its SUM and COUNT passes are matched up through sorted outputs (groupby.cuh:346-386)
    gdf_error e = sort_groups(ncols, cols, out_vals, out_agg, op == OP_COUNT || op == OP_AVG ? out_agg->dtype : col_agg->dtype,
                              groups);
    if (e != GDF_SUCCESS) return e;
  }
  for (int c = 0; c < ncols; ++c) out_vals[c]->size = groups;
  out_agg->size = groups;
  return GDF_SUCCESS;
}

// GDF_SORT method (ref sqls_ops.cu:1134-1289, sqls_rtti_comp.hpp:372-640).  Observable contract of the reference:
// groups come back in lexicographic key order; out_col_agg holds one aggregate per group; out_col_indices (if given)
// holds, as size_t, the index of one row of each group (the reference returns whichever row its unstable sort put
// first); out_col_values (if given) the group keys; flag_distinct with COUNT returns the NUMBER of groups in
// out_col_agg[0] and size 1.  B200 design: sorting 1e9 rows only to reduce neighbouring runs moves every row through
// ~8 radix passes; the same contract is met by the hash aggregation above followed by a sort of the (few) groups, and
// the representative rows come from one binary-search pass over the input.  Floating-point sums differ from a
// sort-and-reduce order within the tolerance stated in the tests (the reference's own order is not deterministic).
gdf_error group_by_sort_method(int ncols, gdf_column** cols, gdf_column* col_agg, gdf_column* out_col_indices,
                               gdf_column** out_col_values, gdf_column* out_col_agg, gdf_context* ctxt, int op) {
  const size_t n = cols[0]->size;
  B200_REQUIRE(ncols <= kMaxCols, GDF_JOIN_TOO_MANY_COLUMNS);
  // group keys are needed internally even when the caller does not ask for them
  std::vector<Scratch> tmp((size_t)ncols);
  std::vector<gdf_column> tmp_cols((size_t)ncols);
  std::vector<gdf_column*> key_out((size_t)ncols);
  for (int c = 0; c < ncols; ++c) {
    B200_REQUIRE(hashable_dtype(cols[c]->dtype), GDF_UNSUPPORTED_DTYPE);
    if (out_col_values && out_col_values[c] && out_col_values[c]->data) {
      key_out[c] = out_col_values[c];
    } else {
      B200_CUDA_TRY(tmp[c].alloc(n * (size_t)dtype_width(cols[c]->dtype)));
      gdf_column_view(&tmp_cols[c], tmp[c].ptr, nullptr, n, cols[c]->dtype);
      key_out[c] = &tmp_cols[c];
    }
  }
  gdf_error e = group_by_hash(ncols, cols, col_agg, key_out.data(), out_col_agg, op, true);
  if (e != GDF_SUCCESS) return e;
  const size_t groups = out_col_agg->size;
  if (out_col_values)
    for (int c = 0; c < ncols; ++c)
      if (out_col_values[c]) {
        out_col_values[c]->dtype = cols[c]->dtype;  // multi_gather_host, sqls_ops.cu:145-165
        out_col_values[c]->size = groups;
      }
  if (out_col_indices && out_col_indices->data && groups) {
    unsigned long long* rep = static_cast<unsigned long long*>(out_col_indices->data);
    fill_u64_kernel<<<grid_for(groups), kThreads>>>(rep, groups, ~0ull);
    TableView in, sorted;
    make_view(in, cols, ncols);
    make_view(sorted, key_out.data(), ncols);
    sorted.rows = groups;
    for (int c = 0; c < ncols; ++c) sorted.dtype[c] = (unsigned char)cols[c]->dtype;
    group_rows_kernel<<<grid_for(n), kThreads>>>(in, sorted, rep);
    B200_CHECK_LAST();
  }
  if (out_col_indices) out_col_indices->size = groups;
  if (op == OP_COUNT && ctxt->flag_distinct) {  // COUNT DISTINCT: one row holding the number of groups
    switch (out_col_agg->dtype) {
      case GDF_INT8: store_scalar_kernel<<<1, 1>>>(static_cast<int8_t*>(out_col_agg->data), (double)groups); break;
      case GDF_INT16: store_scalar_kernel<<<1, 1>>>(static_cast<int16_t*>(out_col_agg->data), (double)groups); break;
      case GDF_INT32: case GDF_DATE32: store_scalar_kernel<<<1, 1>>>(static_cast<int32_t*>(out_col_agg->data), (double)groups); break;
      case GDF_FLOAT32: store_scalar_kernel<<<1, 1>>>(static_cast<float*>(out_col_agg->data), (double)groups); break;
      case GDF_FLOAT64: store_scalar_kernel<<<1, 1>>>(static_cast<double*>(out_col_agg->data), (double)groups); break;
      default: store_scalar_kernel<<<1, 1>>>(static_cast<int64_t*>(out_col_agg->data), (double)groups); break;
    }
    B200_CHECK_LAST();
    out_col_agg->size = 1;
    if (out_col_indices) out_col_indices->size = 1;
    if (out_col_values)
      for (int c = 0; c < ncols; ++c)
        if (out_col_values[c]) out_col_values[c]->size = 1;
  }
  return GDF_SUCCESS;
}

gdf_error group_by_single(int ncols, gdf_column** cols, gdf_column* col_agg, gdf_column* out_col_indices,
                          gdf_column** out_col_values, gdf_column* out_col_agg, gdf_context* ctxt, int op) {
  if (ncols == 0 || cols == nullptr || col_agg == nullptr || out_col_agg == nullptr || ctxt == nullptr)
    return GDF_DATASET_EMPTY;
  for (int i = 0; i < ncols; ++i) B200_REQUIRE(!cols[i]->valid, GDF_VALIDITY_UNSUPPORTED);
  B200_REQUIRE(!col_agg->valid, GDF_VALIDITY_UNSUPPORTED);
  if (cols[0]->size == 0 || col_agg->size == 0) {
    out_col_agg->size = 0;
    if (out_col_indices) out_col_indices->size = 0;
    if (out_col_values)
      for (int c = 0; c < ncols; ++c)
        if (out_col_values[c]) out_col_values[c]->size = 0;
    return GDF_SUCCESS;
  }
  if (ctxt->flag_method != GDF_HASH && ctxt->flag_method != GDF_SORT) return GDF_UNSUPPORTED_METHOD;
  gdf_nvtx_range_push("LIBGDF_GROUPBY", GDF_GREEN);
  gdf_error e = GDF_SUCCESS;
  if (ctxt->flag_method == GDF_HASH) {
    e = group_by_hash(ncols, cols, col_agg, out_col_values, out_col_agg, op, ctxt->flag_sort_result == 1);
  } else {
    e = group_by_sort_method(ncols, cols, col_agg, out_col_indices, out_col_values, out_col_agg, ctxt, op);
  }
  gdf_nvtx_range_pop();
  return e;
}

}  // namespace
}  // namespace b200

using namespace b200;

#define B200_GROUPBY(NAME, OP)                                                                                  \
  extern "C" gdf_error gdf_group_by_##NAME(int ncols, gdf_column** cols, gdf_column* col_agg,                   \
                                           gdf_column* out_col_indices, gdf_column** out_col_values,            \
                                           gdf_column* out_col_agg, gdf_context* ctxt) {                        \
    return group_by_single(ncols, cols, col_agg, out_col_indices, out_col_values, out_col_agg, ctxt, OP);       \
  }
B200_GROUPBY(sum, OP_SUM)
B200_GROUPBY(min, OP_MIN)
B200_GROUPBY(max, OP_MAX)
B200_GROUPBY(avg, OP_AVG)

extern "C" gdf_error gdf_group_by_count(int ncols, gdf_column** cols, gdf_column* col_agg,
                                        gdf_column* out_col_indices, gdf_column** out_col_values,
                                        gdf_column* out_col_agg, gdf_context* ctxt) {
  if (ctxt && ctxt->flag_distinct && ctxt->flag_method != GDF_SORT) return GDF_UNSUPPORTED_METHOD;  // ref sqls_ops.cu:1357-1359 (hash path)
  return group_by_single(ncols, cols, col_agg, out_col_indices, out_col_values, out_col_agg, ctxt, OP_COUNT);
}
