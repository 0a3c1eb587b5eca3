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
//   * pass 2 ranks the bits (warp scan of per-lane counts) in (step, lane, element) order = ascending row order and
//     calls the policy's emit_dense / emit_step / emit.
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

// Writes the payloads of the selected rows (bit j of `bits` = row0 + j) of one thread's step to pos, pos + 1, ...
// A policy whose payload has to be LOADED first (gpu_apply_stencil: out[pos] = data[row]) provides
// emit_step(row0, bits, pos) and issues all its loads before the first store: with per-row emit() every store waits for
// its own load, and a warp paid V = 16 dependent DRAM round trips per step (ncu, profiles/r02_notes.md: 81 % of the
// stall samples on long_scoreboard, 3.5 ms at C2's size for 8.5 GB of traffic).
template <typename Policy>
static __device__ __forceinline__ auto emit_step(const Policy& pol, size_t row0, uint32_t bits, size_t pos)
    -> decltype(pol.emit_step(row0, bits, pos), void()) {
  pol.emit_step(row0, bits, pos);
}
template <typename Policy, typename... Rest>
static __device__ __forceinline__ void emit_step(const Policy& pol, size_t row0, uint32_t bits, size_t pos, Rest...) {
#pragma unroll
  for (int j = 0; j < Policy::V; ++j)
    if ((bits >> j) & 1u) pol.emit(row0 + j, pos++);
}

// Warp-wide alternative for a step with many selected rows: a policy may provide
//   bool emit_dense(size_t step_row0, size_t n, uint32_t my_bits, size_t my_pos, uint32_t step_total) const
// called by ALL 32 lanes (warp-uniform arguments except my_bits / my_pos); it returns false when it does not apply
// (few rows selected, ragged / unaligned step) and the per-lane emit_step runs instead.
template <typename Policy>
static __device__ __forceinline__ auto emit_dense(const Policy& pol, size_t step_row0, size_t n, uint32_t bits, size_t pos,
                                                  uint32_t step_total) -> decltype(pol.emit_dense(step_row0, n, bits, pos, step_total)) {
  return pol.emit_dense(step_row0, n, bits, pos, step_total);
}
template <typename Policy, typename... Rest>
static __device__ __forceinline__ bool emit_dense(const Policy&, size_t, size_t, uint32_t, size_t, uint32_t, Rest...) {
  return false;
}

template <typename Policy>
__global__ void __launch_bounds__(kThreads, 3)
select_chunked_kernel(Policy pol, size_t n, uint64_t* __restrict__ desc, unsigned long long* __restrict__ count_out,
                      unsigned* __restrict__ ticket) {
  constexpr int V = Policy::V, K = Policy::K, R = V * K;
  using G = Geom<Policy>;
  // The predicate bits of the chunk live in shared memory (16 KB), not in 16 registers per thread, and pass 2 is a
  // real loop: fully unrolled it was 55-70 K SASS instructions (13 % of the stall samples on instruction fetch, ncu)
  // and 126 registers (two CTAs per SM).
  __shared__ uint32_t fbits[kChunkTiles][kThreads];
  __shared__ uint32_t warp_tot[kChunkTiles][kWarps];
  __shared__ uint64_t chunk_excl;
  __shared__ unsigned chunk_id;
  const unsigned tid = threadIdx.x, warp = tid >> 5, lane = tid & 31u;
  const size_t chunks = (n + G::kChunkRows - 1) / G::kChunkRows;
  while (true) {
    if (tid == 0) chunk_id = atomicAdd(ticket, 1u);
    __syncthreads();
    const size_t chunk = chunk_id;
    if (chunk >= chunks) break;
#pragma unroll 4
    for (int k = 0; k < kChunkTiles; ++k) {
      const size_t warp_base = (chunk * kChunkTiles + k) * G::kTileRows + (size_t)warp * (32 * R);
      const uint32_t f = warp_base < n ? pol.flags(warp_base, n) : 0u;
      fbits[k][tid] = f;
      uint32_t wsum = __popc(f);
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
#pragma unroll 1
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
      const uint32_t f = fbits[k][tid];
      const size_t warp_base = (chunk * kChunkTiles + k) * G::kTileRows + (size_t)warp * (32 * R);
#pragma unroll
      for (int s = 0; s < K; ++s) {
        // rows of one step are ordered (lane, j): rank = selected rows of lower lanes (all j) + my lower j
        const uint32_t my_step = (f >> (s * V)) & ((V == 32) ? 0xffffffffu : ((1u << V) - 1u));
        const uint32_t cnt = __popc(my_step);
        uint32_t inc = cnt;   // inclusive scan of the lanes' counts: 5 shuffles instead of V ballots
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          const uint32_t o = __shfl_up_sync(0xffffffffu, inc, d);
          if (lane >= (unsigned)d) inc += o;
        }
        const uint32_t below = inc - cnt;
        const uint32_t step_total = __shfl_sync(0xffffffffu, inc, 31);
        const size_t step_row0 = warp_base + (size_t)s * (32 * V);
        if (step_total != 0 && !emit_dense(pol, step_row0, n, my_step, pos + below, step_total) && my_step)
          emit_step(pol, step_row0 + (size_t)lane * V, my_step, pos + below);
        pos += step_total;
      }
    }
    __syncthreads();  // fbits / warp_tot / chunk_excl / chunk_id are rewritten by the next chunk
  }
}

}  // namespace select_chunked
}  // namespace b200
