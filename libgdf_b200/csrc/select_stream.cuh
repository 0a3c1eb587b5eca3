// Streaming select (v2) for ONE aligned column: persistent CTAs, TMA-staged tiles, wide look-back.
//
// Why a second kernel (profiles/r01a_ncu_full_summary.md, select_kernel): with one tile per CTA and
// a 32-descriptor look-back window the kernel ran at 2.7 TB/s with 24 warps-per-issue stalled on the
// CTA barrier: ~450 tiles are in flight on 148 SMs, every tile has to walk back over the aggregates
// of the other in-flight tiles 32 at a time (one L2 round trip each), and nothing is loading while a
// CTA waits.  Here:
//   * CTAs are persistent and take tiles from a global ticket counter (so a tile's predecessors are
//     always held by running CTAs); tiles are 32 KB and arrive through a 3-stage shared-memory ring
//     filled by cp.async.bulk (stream.cuh) - the next two tiles are in flight while the current one
//     is ranked, looked back and written;
//   * the look-back reads 256 descriptors per round trip (8 per lane of warp 0), so one round
//     normally covers every in-flight tile;
//   * each thread owns R CONSECUTIVE rows of the tile (read from shared memory with an XOR swizzle
//     so that the 128-bit reads are bank-conflict free), which makes the output rank a plain
//     popc + warp scan instead of R ballots.
// Output order is ascending row order, exactly like the first kernel (and like thrust::copy_if in
// the reference, sqls_rtti_comp.hpp:358-365).
#pragma once
#include "select.cuh"
#include "stream.cuh"

namespace b200 {
namespace select_stream {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kStages = 3;

template <typename T>
struct Geom {
  static constexpr int R = sizeof(T) == 8 ? 16 : 32;          // rows per thread (flags fit 32 bits)
  static constexpr int kTileRows = kThreads * R;
  static constexpr int kTileBytes = kTileRows * (int)sizeof(T);  // 32 KB for 8- and 4-byte types
  static constexpr int C = R * (int)sizeof(T) / 16;            // 16-byte chunks per thread
  static constexpr int E = 16 / (int)sizeof(T);                // elements per chunk
};

struct Smem {
  uint64_t bar[kStages];
  unsigned tile[kStages];
  uint32_t warp_tot[2][kWarps];
  uint64_t tile_excl[2];
};

template <typename T>
constexpr size_t smem_bytes() {
  return (size_t)kStages * Geom<T>::kTileBytes + sizeof(Smem) + 128;
}

// Publishes this tile's total and returns the exclusive prefix of all earlier tiles; runs in warp 0.
static __device__ __forceinline__ uint64_t lookback_wide(uint64_t* desc, unsigned tile, uint64_t total) {
  using namespace select_detail;
  const unsigned lane = lane_id();
  if (tile == 0) {
    if (lane == 0) st_desc(desc, kPrefix | total);
    return 0;
  }
  if (lane == 0) st_desc(desc + tile, kAgg | total);
  uint64_t acc = 0;
  long long base = (long long)tile - 1;  // nearest predecessor
  bool done = false;
  while (!done) {
    uint64_t d[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {  // window j = 32 consecutive predecessors, nearest first
      const long long idx = base - (j * 32 + (int)lane);
      d[j] = idx >= 0 ? ld_desc(desc + idx) : kPrefix;  // before tile 0: prefix 0
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (done) break;
      const long long idx = base - (j * 32 + (int)lane);
      unsigned pre, first;
      while (true) {
        const unsigned st = (unsigned)(d[j] >> 62);
        const unsigned inv = __ballot_sync(0xffffffffu, st == 0);
        pre = __ballot_sync(0xffffffffu, st == 2);
        first = pre ? (unsigned)(__ffs(pre) - 1) : 31u;
        const unsigned need = pre ? ((2u << first) - 1u) : 0xffffffffu;  // lanes up to the nearest prefix
        if ((inv & need) == 0) break;
        if (st == 0) d[j] = ld_desc(desc + idx);  // not published yet: poll again
      }
      if (lane <= first) acc += d[j] & kValMask;
      if (pre) done = true;
    }
    base -= 256;
  }
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
  if (lane == 0) st_desc(desc + tile, kPrefix | (acc + total));
  return acc;
}

// Pred: void prepare() (once per thread, may read device scalars); bool operator()(T) const.
// Emit: void operator()(size_t row, size_t pos) const.
template <typename T, typename Pred, typename Emit>
__global__ void __launch_bounds__(kThreads)
select_stream_kernel(const T* __restrict__ data, size_t n, Pred pred, Emit emit, uint64_t* __restrict__ desc,
                     unsigned* __restrict__ ticket, unsigned long long* __restrict__ count_out) {
  using G = Geom<T>;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* ring = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~(uintptr_t)127);
  Smem& sm = *reinterpret_cast<Smem*>(ring + (size_t)kStages * G::kTileBytes);
  const unsigned tid = threadIdx.x, warp = tid >> 5, lane = tid & 31u;
  const size_t tiles = (n + G::kTileRows - 1) / G::kTileRows;
  pred.prepare();

  // one thread takes a ticket for stage s and starts the copy of that tile
  auto issue = [&](int s) {
    const unsigned t = atomicAdd(ticket, 1u);
    sm.tile[s] = t;
    if ((size_t)t < tiles && ((size_t)t + 1) * G::kTileRows <= n) {
      tma::mbar_expect_tx(&sm.bar[s], G::kTileBytes);
      tma::bulk_load(ring + (size_t)s * G::kTileBytes, data + (size_t)t * G::kTileRows, G::kTileBytes, &sm.bar[s]);
    }
  };

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < kStages; ++s) tma::mbar_init(&sm.bar[s], 1);
    tma::fence_barrier_init();
#pragma unroll
    for (int s = 0; s < kStages; ++s) issue(s);
  }
  __syncthreads();

  for (unsigned iter = 0;; ++iter) {
    const int s = (int)(iter % kStages);
    const unsigned t = sm.tile[s];
    if ((size_t)t >= tiles) break;  // tickets only grow: nothing of this CTA is in flight any more
    const size_t tile_row0 = (size_t)t * G::kTileRows;
    const bool full = tile_row0 + G::kTileRows <= n;
    uint32_t f = 0;
    if (full) {
      tma::mbar_wait(&sm.bar[s], (iter / kStages) & 1u);
      const uint4* mine = reinterpret_cast<const uint4*>(ring + (size_t)s * G::kTileBytes) + (size_t)tid * G::C;
#pragma unroll
      for (int j = 0; j < G::C; ++j) {
        const int c = j ^ (int)(lane & (G::C - 1));  // swizzle: a quarter-warp touches 8 different bank groups
        const uint4 raw = mine[c];
        const T* e = reinterpret_cast<const T*>(&raw);
#pragma unroll
        for (int k = 0; k < G::E; ++k) f |= (uint32_t)pred(e[k]) << (c * G::E + k);
      }
    } else {  // ragged last tile: direct loads
      const size_t row0 = tile_row0 + (size_t)tid * G::R;
#pragma unroll 4
      for (int k = 0; k < G::R; ++k)
        if (row0 + k < n) f |= (uint32_t)pred(data[row0 + k]) << k;
    }
    const uint32_t cnt = __popc(f);
    uint32_t inc = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t o = __shfl_up_sync(0xffffffffu, inc, d);
      if (lane >= (unsigned)d) inc += o;
    }
    const unsigned buf = iter & 1u;
    if (lane == 31) sm.warp_tot[buf][warp] = inc;
    __syncthreads();  // every thread is done with stage s; warp totals are visible
    if (tid == 32) issue(s);
    if (warp == 0) {
      uint32_t tot = 0;
#pragma unroll
      for (int w = 0; w < kWarps; ++w) tot += sm.warp_tot[buf][w];
      const uint64_t excl = lookback_wide(desc, t, tot);
      if (lane == 0) {
        sm.tile_excl[buf] = excl;
        if ((size_t)t == tiles - 1) *count_out = excl + tot;
      }
    }
    __syncthreads();
    if (cnt) {
      size_t pos = sm.tile_excl[buf] + (inc - cnt);
      for (unsigned w = 0; w < warp; ++w) pos += sm.warp_tot[buf][w];
      const size_t row0 = tile_row0 + (size_t)tid * G::R;
      while (f) {
        const int k = __ffs(f) - 1;
        f &= f - 1;
        emit(row0 + k, pos++);
      }
    }
  }
}

}  // namespace select_stream
}  // namespace b200
