// Single-pass, order-preserving stream compaction ("select") for sm_100a.
//
// Used by gdf_filter (payload = row index) and gpu_apply_stencil (payload = column value).  The
// reference runs thrust::copy_if, i.e. a multi-kernel select with a temporary allocation, and reads
// its comparands from global memory per row (ref src/sqls_rtti_comp.hpp:343-370,
// src/streamcompactionops.cu:250-282).  Here the whole operation is ONE kernel that touches every
// input byte exactly once:
//
//   * a tile is kThreads x R rows; a warp owns 32*R consecutive rows and reads them "warp-striped"
//     (lane l, step k -> rows k*32*V + l*V .. +V-1) so every load instruction of a warp covers one
//     contiguous 512-byte span: perfectly coalesced 128-bit no-allocate loads, all K loads of a thread
//     issued before the first use;
//   * the thread keeps only the R predicate bits, not the values;
//   * tile totals are chained with a decoupled look-back over 64-bit {status,value} descriptors
//     (one relaxed 64-bit word per tile, so status and value can never be observed torn);
//   * output ranks inside a warp come from ballots in (step, lane, element) order, which is exactly
//     ascending row order, so the result is stable like copy_if.
//
// Algorithmic traffic: rows*width read + selected*payload written; descriptors add 8 B per tile.
#pragma once
#include "common.cuh"

namespace b200 {
namespace select_detail {

constexpr int kThreads = 256;
constexpr uint64_t kAgg = 1ull << 62;
constexpr uint64_t kPrefix = 2ull << 62;
constexpr uint64_t kValMask = (1ull << 62) - 1;

static __device__ __forceinline__ uint64_t ld_desc(const uint64_t* p) {
  uint64_t v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
static __device__ __forceinline__ void st_desc(uint64_t* p, uint64_t v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// Runs in warp 0 of a tile.  Publishes this tile's total, returns the exclusive prefix of all
// earlier tiles (broadcast to every lane of warp 0).
static __device__ __forceinline__ uint64_t lookback(uint64_t* desc, unsigned tile, uint64_t total) {
  const unsigned lane = lane_id();
  if (tile == 0) {
    if (lane == 0) st_desc(desc, kPrefix | total);
    return 0;
  }
  if (lane == 0) st_desc(desc + tile, kAgg | total);
  uint64_t excl = 0;
  long long idx = (long long)tile - 1 - lane;  // lane 0 looks at the nearest predecessor
  while (true) {
    uint64_t d = kPrefix;  // tiles before 0 behave like a published prefix of 0
    if (idx >= 0) {
      do { d = ld_desc(desc + idx); } while ((d >> 62) == 0);
    }
    const unsigned has_prefix = __ballot_sync(0xffffffffu, (d >> 62) == 2);
    // take every lane up to and including the nearest lane holding a full prefix
    const unsigned first = has_prefix ? (__ffs(has_prefix) - 1) : 31;
    uint64_t v = (lane <= first) ? (d & kValMask) : 0;
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
    excl += v;
    if (has_prefix) break;
    idx -= 32;
  }
  if (lane == 0) st_desc(desc + tile, kPrefix | (excl + total));
  return excl;
}

}  // namespace select_detail

// Policy concept:
//   static constexpr int V;   rows per 128-bit load step      (16 / element width)
//   static constexpr int K;   load steps per thread           (R = V*K <= 32 rows per thread)
//   __device__ uint32_t flags(size_t warp_base, size_t n) const;   predicate bits, bit (k*V+j)
//   __device__ void emit(size_t row, size_t pos) const;            write payload of `row` at `pos`
template <typename Policy>
__global__ void __launch_bounds__(select_detail::kThreads)
select_kernel(Policy pol, size_t n, uint64_t* __restrict__ desc, unsigned long long* __restrict__ count_out) {
  using namespace select_detail;
  constexpr int V = Policy::V, K = Policy::K, R = V * K;
  constexpr int kWarps = kThreads / 32;
  __shared__ uint32_t warp_tot[kWarps];
  __shared__ uint64_t tile_excl;
  const unsigned tile = blockIdx.x;
  const unsigned warp = threadIdx.x >> 5, lane = lane_id();
  const size_t warp_base = (size_t)tile * (kThreads * R) + (size_t)warp * (32 * R);

  const uint32_t f = pol.flags(warp_base, n);

  uint32_t wsum = __popc(f);
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) wsum += __shfl_xor_sync(0xffffffffu, wsum, s);
  if (lane == 0) warp_tot[warp] = wsum;
  __syncthreads();
  if (warp == 0) {
    uint32_t t = 0;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) t += warp_tot[w];
    const uint64_t excl = lookback(desc, tile, t);
    if (lane == 0) {
      tile_excl = excl;
      if (tile == gridDim.x - 1) *count_out = excl + t;
    }
  }
  __syncthreads();
  size_t pos = tile_excl;
  for (unsigned w = 0; w < warp; ++w) pos += warp_tot[w];
  if (wsum == 0) return;  // warp-uniform

  const unsigned lt = lanemask_lt();
#pragma unroll
  for (int k = 0; k < K; ++k) {
    // rows of one step are ordered (lane, j): rank = selected rows of lower lanes (all j) + my lower j
    const uint32_t my_step = (f >> (k * V)) & ((V == 32) ? 0xffffffffu : ((1u << V) - 1u));
    uint32_t below = 0, step_total = 0;
#pragma unroll
    for (int j = 0; j < V; ++j) {
      const unsigned b = __ballot_sync(0xffffffffu, (my_step >> j) & 1u);
      below += __popc(b & lt);
      step_total += __popc(b);
    }
    if (my_step) {
      size_t p = pos + below;
      const size_t row0 = warp_base + (size_t)k * (32 * V) + (size_t)lane * V;
#pragma unroll
      for (int j = 0; j < V; ++j)
        if ((my_step >> j) & 1u) pol.emit(row0 + j, p++);
    }
    pos += step_total;
  }
}

// Host helper: number of tiles for n rows under Policy.
template <typename Policy>
inline size_t select_tiles(size_t n) {
  const size_t tile_rows = (size_t)select_detail::kThreads * Policy::V * Policy::K;
  return (n + tile_rows - 1) / tile_rows;
}

}  // namespace b200
