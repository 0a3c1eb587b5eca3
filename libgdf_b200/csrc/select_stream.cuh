// Streaming select (v3) for ONE aligned column: persistent CTAs, TMA-staged tiles, CHUNKED look-back.
//
// History (profiles/): v1 (select.cuh, one 32 KB tile per CTA, 32-descriptor look-back) ran C2 at
// 2.7 TB/s with 24 warps-per-issue stalled on the CTA barrier.  The cause is structural: a tile cannot
// be written before the running total of ALL earlier tiles is known, and that total travels through
// L2 one look-back hop (~1 us under load) at a time.  A hop that covers 32 tiles of 32 KB moves the
// prefix 1 MB per microsecond - 1 TB/s.  v2 kept the per-tile protocol (persistent CTAs, 256-wide
// hops) and measured even worse (10 ms): with ~300 CTAs in lock step every iteration still waits
// for three serial hops.  v3 changes the granularity of the chain instead:
//   * a CTA processes CHUNKS of 16 consecutive tiles (64 K int64 rows, 512 KB).  Pass 1 streams the
//     16 tiles through a 3-stage cp.async.bulk ring (stream.cuh), and every thread keeps the predicate
//     bits of its rows - 16 consecutive rows per tile - in registers (the output of gdf_filter is the
//     row INDEX, so the data is not needed again);
//   * ONE descriptor per chunk is published and ONE look-back per chunk is done (a hop now moves the
//     prefix 256 x 512 KB); the next chunk's first tiles are already in flight while warp 0 waits;
//   * pass 2 ranks the kept bits (warp scan per tile done in pass 1, one pass over the 16 x 8 warp
//     totals) and writes the indices, ascending.
// Output order is ascending row order (stable), like thrust::copy_if in the reference
// (sqls_rtti_comp.hpp:358-365).
#pragma once
#include "select.cuh"
#include "stream.cuh"

namespace b200 {
namespace select_stream {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kStages = 3;
constexpr int kChunkTiles = 16;
constexpr int kTicketTile = kChunkTiles - kStages - 3;   // tile after which the next chunk's ticket is taken (see below)

template <typename T>
struct Geom {
  static constexpr int R = sizeof(T) == 8 ? 16 : 32;          // rows per thread per tile (flags fit 32 bits)
  static constexpr int kTileRows = kThreads * R;
  static constexpr int kTileBytes = kTileRows * (int)sizeof(T);  // 32 KB for 8- and 4-byte types
  static constexpr int kChunkRows = kTileRows * kChunkTiles;
  static constexpr int C = R * (int)sizeof(T) / 16;            // 16-byte chunks per thread
  static constexpr int E = 16 / (int)sizeof(T);                // elements per 16-byte chunk
};

struct Smem {
  uint64_t bar[kStages];
  uint32_t warp_tot[kChunkTiles][kWarps];
  uint64_t chunk_excl;
  unsigned chunk_cur;      // ticket of the chunk the CTA is about to process (written by thread 0 one barrier earlier)
};

template <typename T>
constexpr size_t smem_bytes() {
  return (size_t)kStages * Geom<T>::kTileBytes + sizeof(Smem);
}

// Publishes this chunk's total and returns the exclusive prefix of all earlier chunks; runs in warp 0.
// 256 descriptors (8 per lane) are read per round trip.
static __device__ __forceinline__ uint64_t lookback_wide(uint64_t* desc, unsigned idx0, uint64_t total) {
  using namespace select_detail;
  const unsigned lane = lane_id();
  if (idx0 == 0) {
    if (lane == 0) st_desc(desc, kPrefix | total);
    return 0;
  }
  if (lane == 0) st_desc(desc + idx0, kAgg | total);
  uint64_t acc = 0;
  long long base = (long long)idx0 - 1;  // nearest predecessor
  bool done = false;
  while (!done) {
    uint64_t d[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {  // window j = 32 consecutive predecessors, nearest first
      const long long idx = base - (j * 32 + (int)lane);
      d[j] = idx >= 0 ? ld_desc(desc + idx) : kPrefix;  // before the first: prefix 0
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
  if (lane == 0) st_desc(desc + idx0, kPrefix | (acc + total));
  return acc;
}

// Pred: void prepare() (once per thread, may read device scalars); bool operator()(T) const.
// Emit: void operator()(size_t row, size_t pos) const.
// Chunks are taken through an atomic TICKET, one chunk ahead of the one being processed (the ring prefetches
// across the chunk boundary).  Tickets are handed out in increasing order to CTAs that are already running, so
// every predecessor of a chunk is held by a running CTA: the look-back always makes progress, whether or not the
// whole grid is co-resident (another stream's persistent kernel may hold SMs).  Round 1 dealt chunks statically
// (CTA b: b, b + grid, ...), which was only safe with the full grid resident.
// WHEN the ticket is taken matters (profiles/r02_notes.md): a CTA that reserves its next chunk at the START of the
// current one holds that chunk idle for a whole chunk time while CTAs with later tickets already wait for its
// aggregate in their look-back - with two CTAs per SM the two waves ended up alternating (one streams, the other
// waits), C2 1.54 -> 1.87 ms.  The ticket is therefore taken as late as the prefetch allows: after tile
// kTicketTile = 10, three tiles (~5 us, well above the atomic's latency) before tile 13's prefetch needs it.  Only
// thread 0 starts copies, so the ticket lives in one of its registers and reaches the other threads through shared
// memory at the chunk's last barrier; nothing ever waits for the atomic.
// STATIC_DEAL = round 1's dealing, kept for A/B in the lab build only.
template <typename T, typename Pred, typename Emit, bool STATIC_DEAL = false>
__global__ void __launch_bounds__(kThreads)
select_stream_kernel(const T* __restrict__ data, size_t n, Pred pred, Emit emit, uint64_t* __restrict__ desc,
                     unsigned long long* __restrict__ count_out, unsigned* __restrict__ ticket) {
  using G = Geom<T>;
  extern __shared__ __align__(16) unsigned char select_smem[];
  unsigned char* const ring = select_smem;
  Smem& sm = *reinterpret_cast<Smem*>(select_smem + (size_t)kStages * G::kTileBytes);
  const unsigned tid = threadIdx.x, warp = tid >> 5, lane = tid & 31u;
  const size_t tiles = (n + G::kTileRows - 1) / G::kTileRows;
  const size_t chunks = (tiles + kChunkTiles - 1) / kChunkTiles;
  pred.prepare();

  // The CTA's tiles form one sequence q = 0, 1, 2, ...: tile q is tile (q % 16) of the CTA's (q / 16)-th chunk and
  // lives in ring stage q % kStages.  Only thread 0 starts copies; it holds the tickets of the CTA's current and
  // next chunk in registers (ids[CTA-local chunk number & 1]).
  unsigned id_cur = 0u, id_next = 0u;   // id_next is written by the atomic and not touched until tile 13's prefetch
  auto issue = [&](unsigned q, bool next_chunk) {  // thread 0 starts the copy of the CTA's q-th tile (full tiles only)
    const size_t t = (size_t)(next_chunk ? id_next : id_cur) * kChunkTiles + (q % kChunkTiles);
    const int s = (int)(q % kStages);
    if (t < tiles && (t + 1) * G::kTileRows <= n) {
      tma::mbar_expect_tx(&sm.bar[s], G::kTileBytes);
      tma::bulk_load(ring + (size_t)s * G::kTileBytes, data + t * G::kTileRows, G::kTileBytes, &sm.bar[s]);
    }
  };
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < kStages; ++s) tma::mbar_init(&sm.bar[s], 1);
    tma::fence_barrier_init();
    id_cur = STATIC_DEAL ? blockIdx.x : atomicAdd(ticket, 1u);
    sm.chunk_cur = id_cur;
#pragma unroll
    for (int s = 0; s < kStages; ++s) issue((unsigned)s, false);
  }
  __syncthreads();

  unsigned q = 0;                  // tiles consumed so far by this CTA
  unsigned phase_bits = 0;         // bit s = parity the next FULL tile of stage s completes
  for (unsigned j = 0;; ++j) {     // j = CTA-local chunk number
    const size_t chunk = sm.chunk_cur;
    if (chunk >= chunks) break;
    uint32_t f[kChunkTiles], excl[kChunkTiles];
#pragma unroll
    for (int k = 0; k < kChunkTiles; ++k, ++q) {
      const size_t t = chunk * kChunkTiles + k;
      const size_t tile_row0 = t * G::kTileRows;
      const int s = (int)(q % kStages);
      uint32_t bits = 0;
      if (tile_row0 + G::kTileRows <= n) {  // full tile: staged by TMA
        tma::mbar_wait(&sm.bar[s], (phase_bits >> s) & 1u);
        phase_bits ^= 1u << s;
        const uint4* mine = reinterpret_cast<const uint4*>(ring + (size_t)s * G::kTileBytes) + (size_t)tid * G::C;
#pragma unroll
        for (int j = 0; j < G::C; ++j) {
          const int c = j ^ (int)(lane & (G::C - 1));  // swizzle: a quarter-warp touches 8 different bank groups
          const uint4 raw = mine[c];
          const T* e = reinterpret_cast<const T*>(&raw);
#pragma unroll
          for (int i = 0; i < G::E; ++i) bits |= (uint32_t)pred(e[i]) << (c * G::E + i);
        }
      } else if (tile_row0 < n) {  // ragged last tile: direct loads
        const size_t row0 = tile_row0 + (size_t)tid * G::R;
#pragma unroll 4
        for (int i = 0; i < G::R; ++i)
          if (row0 + i < n) bits |= (uint32_t)pred(data[row0 + i]) << i;
      }
      f[k] = bits;
      uint32_t inc = __popc(bits);
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const uint32_t o = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= (unsigned)d) inc += o;
      }
      excl[k] = inc - __popc(bits);
      if (lane == 31) sm.warp_tot[k][warp] = inc;
      __syncthreads();  // every thread is done with stage s (and, for k = 15, all warp totals are visible)
      if (k == kTicketTile && tid == 0) id_next = STATIC_DEAL ? (unsigned)chunk + gridDim.x : atomicAdd(ticket, 1u);
      if (tid == 0) issue(q + kStages, k + kStages >= kChunkTiles);
    }
    if (warp == 0) {
      uint32_t part = 0;
#pragma unroll
      for (int j = 0; j < kChunkTiles * kWarps / 32; ++j) part += (&sm.warp_tot[0][0])[j * 32 + lane];
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) part += __shfl_xor_sync(0xffffffffu, part, d);
      const uint64_t before = lookback_wide(desc, (unsigned)chunk, part);
      if (lane == 0) {
        sm.chunk_excl = before;
        if (chunk == chunks - 1) *count_out = before + part;
      }
    }
    __syncthreads();
    // pass 2: ranks in (tile, warp, lane, bit) order = ascending row order
    size_t run = (size_t)sm.chunk_excl;
#pragma unroll
    for (int k = 0; k < kChunkTiles; ++k) {
      size_t pos = run + excl[k];
      uint32_t tile_total = 0;
#pragma unroll
      for (int w = 0; w < kWarps; ++w) {
        const uint32_t wt = sm.warp_tot[k][w];
        if ((unsigned)w < warp) pos += wt;
        tile_total += wt;
      }
      run += tile_total;
      uint32_t bits = f[k];
      const size_t row0 = (chunk * kChunkTiles + k) * G::kTileRows + (size_t)tid * G::R;
      while (bits) {
        const int i = __ffs(bits) - 1;
        bits &= bits - 1;
        emit(row0 + i, pos++);
      }
    }
    if (tid == 0) sm.chunk_cur = id_cur = id_next;
    __syncthreads();  // warp_tot / chunk_excl are rewritten by the next chunk; chunk_cur is visible
  }
}

}  // namespace select_stream
}  // namespace b200
