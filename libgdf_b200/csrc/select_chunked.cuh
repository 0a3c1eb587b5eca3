// Chunked, order-preserving stream compaction for ANY select policy (multi-column / unaligned gdf_filter,
// gpu_apply_stencil, the FULL join's unmatched-row pass).
//
// Round 1 left these callers on the first-generation kernel (select.cuh: one tile per CTA, one look-back per tile):
// 0.42 of the measured HBM peak, because a tile cannot be written before the running total of ALL earlier tiles is
// known and that total moves through L2 one look-back hop (~1 us) at a time - 32 tiles of 32 KB per hop = 1 TB/s.
// select_stream.cuh fixed that for the one-aligned-column gdf_filter by chaining CHUNKS of 16 tiles; this kernel
// applies the same protocol to a policy:
//   * persistent CTAs take chunks of 16 tiles through an atomic TICKET (so a chunk's predecessors are always held
//     by CTAs that are already running: forward progress does not depend on the whole grid being co-resident);
//   * pass 1 evaluates the policy's predicate bits for the 16 tiles with no barrier in between (all loads of the
//     chunk can be in flight together) and keeps only the bits;
//   * ONE descriptor / ONE look-back per chunk (256 descriptors per hop, select_stream::lookback_wide);
//   * pass 2 ranks the bits with ballots in (step, lane, element) order = ascending row order and calls emit.
// Policy concept: the one of select.cuh (V, K, flags(warp_base, n), emit(row, pos)).
#pragma once
#include "select.cuh"
#include "select_stream.cuh"

namespace b200 {
namespace select_chunked {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kChunkTiles = 16;

template <typename Policy>
struct Geom {
  static constexpr int R = Policy::V * Policy::K;              // rows per thread per tile
  static constexpr size_t kTileRows = (size_t)kThreads * R;
  static constexpr size_t kChunkRows = kTileRows * kChunkTiles;
};

template <typename Policy>
__global__ void __launch_bounds__(kThreads)
select_chunked_kernel(Policy pol, size_t n, uint64_t* __restrict__ desc, unsigned long long* __restrict__ count_out,
                      unsigned* __restrict__ ticket) {
  constexpr int V = Policy::V, K = Policy::K, R = V * K;
  using G = Geom<Policy>;
  __shared__ uint32_t warp_tot[kChunkTiles][kWarps];
  __shared__ uint64_t chunk_excl;
  __shared__ unsigned chunk_id;
  const unsigned tid = threadIdx.x, warp = tid >> 5, lane = tid & 31u;
  const size_t chunks = (n + G::kChunkRows - 1) / G::kChunkRows;
  const unsigned lt = lanemask_lt();
  while (true) {
    if (tid == 0) chunk_id = atomicAdd(ticket, 1u);
    __syncthreads();
    const size_t chunk = chunk_id;
    if (chunk >= chunks) break;
    uint32_t f[kChunkTiles];
#pragma unroll
    for (int k = 0; k < kChunkTiles; ++k) {
      const size_t warp_base = (chunk * kChunkTiles + k) * G::kTileRows + (size_t)warp * (32 * R);
      f[k] = warp_base < n ? pol.flags(warp_base, n) : 0u;
      uint32_t wsum = __popc(f[k]);
#pragma unroll
      for (int s = 16; s > 0; s >>= 1) wsum += __shfl_xor_sync(0xffffffffu, wsum, s);
      if (lane == 0) warp_tot[k][warp] = wsum;
    }
    __syncthreads();
    if (warp == 0) {
      uint32_t part = 0;
#pragma unroll
      for (int j = 0; j < kChunkTiles * kWarps / 32; ++j) part += (&warp_tot[0][0])[j * 32 + lane];
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) part += __shfl_xor_sync(0xffffffffu, part, d);
      const uint64_t before = select_stream::lookback_wide(desc, (unsigned)chunk, part);
      if (lane == 0) {
        chunk_excl = before;
        if (chunk == chunks - 1) *count_out = before + part;
      }
    }
    __syncthreads();
    size_t run = (size_t)chunk_excl;
#pragma unroll
    for (int k = 0; k < kChunkTiles; ++k) {
      size_t pos = run;
      uint32_t tile_total = 0, mine = 0;
#pragma unroll
      for (int w = 0; w < kWarps; ++w) {
        const uint32_t wt = warp_tot[k][w];
        if ((unsigned)w < warp) pos += wt;
        if ((unsigned)w == warp) mine = wt;
        tile_total += wt;
      }
      run += tile_total;
      if (mine == 0) continue;  // warp-uniform
      const size_t warp_base = (chunk * kChunkTiles + k) * G::kTileRows + (size_t)warp * (32 * R);
#pragma unroll
      for (int s = 0; s < K; ++s) {
        // rows of one step are ordered (lane, j): rank = selected rows of lower lanes (all j) + my lower j
        const uint32_t my_step = (f[k] >> (s * V)) & ((V == 32) ? 0xffffffffu : ((1u << V) - 1u));
        uint32_t below = 0, step_total = 0;
#pragma unroll
        for (int j = 0; j < V; ++j) {
          const unsigned b = __ballot_sync(0xffffffffu, (my_step >> j) & 1u);
          below += __popc(b & lt);
          step_total += __popc(b);
        }
        if (my_step) {
          size_t p = pos + below;
          const size_t row0 = warp_base + (size_t)s * (32 * V) + (size_t)lane * V;
#pragma unroll
          for (int j = 0; j < V; ++j)
            if ((my_step >> j) & 1u) pol.emit(row0 + j, p++);
        }
        pos += step_total;
      }
    }
    __syncthreads();  // warp_tot / chunk_excl / chunk_id are rewritten by the next chunk
  }
}

}  // namespace select_chunked
}  // namespace b200
