// Radix-partitioned hash join for ONE integer key column of 4 or 8 bytes - the shape of the headline
// benchmark (1e9 x 1e8 int64).  Semantics are those of join.cu (same null
// rule, same outputs); only the data movement differs.
//
// Why: a 1e8-row build side needs a ~3 GB table.  Probing it in place costs one random DRAM row
// activation per probe row (the reference pays two: slot, then the build key for rows_equal -
// ref hash/join_kernels.cuh:266-455).  B200 has 126 MB of L2, so instead:
//
//   1. histogram + scatter BOTH sides into NP = 2^k hash partitions of {key,row id} pairs
//      (NP chosen so one partition's table is <= 32 MB).  The scatter stages a 4096-row tile in
//      shared memory, ranks rows with warp-private counters + match.any (no atomics at all in the
//      ranking), reserves each (tile, partition) run with one global atomic and writes runs of
//      consecutive rows - coalesced stores instead of the reference's per-row 8-byte scatters
//      (ref gdf_table.cuh:1071-1192).  NULL-key rows are dropped here (build side, INNER probe side)
//      or tagged so that LEFT/FULL emit (l,-1) without touching a table;
//   2. build one open-addressing table per partition, slot = {key, row} in 16 bytes, so a probe
//      touches a single 32-byte sector; the partition id comes from the TOP hash bits and the slot
//      from the LOW bits, so they never alias;
//   3. probe partition after partition: the CTAs in flight at any moment work on one or two
//      partitions, whose tables therefore stay L2-resident.  Tile-wise count -> block scan -> one
//      cursor atomic per tile -> coalesced index stores.
//
// Algorithmic bytes (C3): 8*(P+B) key bytes in, 8 bytes per output pair out.  Actual DRAM traffic of
// this design: 2 reads of each key (histogram, scatter) + 12 B/row of pairs written and re-read.
#include <cstdlib>

#include "stream.cuh"
#include "table.cuh"

namespace b200 {
namespace {

constexpr int kThreads = 256;
constexpr unsigned kMaxParts = 256;
// Target build rows per partition (8-byte slots: tables <= 16 MB).  2^19 was measured against it at C3 (shipped build):
// probe 8.36 -> 7.82 ms (smaller tables hold on to L2 better while 124 MB of pairs and output stream past them), but
// scatter 4.36 -> 5.05 ms (256 bins: 16-pair runs, twice the bulk stores per row) and build 3.17 -> 3.49 ms: step 17.95 ->
// 18.10 ms.  At 2 GPUs the same change cost 1.7 ms in the exchange and bought nothing in the probe.
constexpr int kRowsPerPartitionLog2 = 20;
constexpr int kScatterRows = 16;                // rows per thread in the scatter tile
constexpr int kScatterTile = kThreads * kScatterRows;
constexpr int kBuildTile = kThreads * 8;
constexpr int kProbeRows = 4;
constexpr int kProbeTile = kThreads * kProbeRows;
constexpr unsigned long long kEmptyKey = ~0ull;

enum JoinKind { JOIN_INNER = 0, JOIN_LEFT = 1, JOIN_FULL = 2 };

struct alignas(16) Slot {
  unsigned long long key;
  int32_t row;
  int32_t pad;
};

// Internal position hash (partition id from its TOP bits, table slot from a re-mix).  Nothing
// observable depends on it - gdf_hash / gdf_hash_partition keep MurmurHash3 - so it is a 3-multiply
// mixer instead: MurmurHash3 cost ~25 of the ~35 instructions per row of the histogram pass, which made
// that pass issue-bound at 4.2 TB/s (profiles/r01a_ncu_full_summary.md).
template <typename KT> struct KeyBits;
template <> struct KeyBits<uint64_t> {
  static __device__ __forceinline__ uint32_t hash(uint64_t k) {
    uint32_t x = (uint32_t)k * 0x9E3779B1u ^ (uint32_t)(k >> 32) * 0x85EBCA77u;
    x ^= x >> 15;
    x *= 0x2C1B3C6Du;
    x ^= x >> 13;
    return x;
  }
};
template <> struct KeyBits<uint32_t> {
  static __device__ __forceinline__ uint32_t hash(uint32_t k) {
    uint32_t x = k * 0x9E3779B1u;
    x ^= x >> 15;
    x *= 0x2C1B3C6Du;
    x ^= x >> 13;
    return x;
  }
};

// Slot index inside a partition's table: all rows of one partition share the top bits of h, so the
// slot comes from a multiplicative re-mix whose middle bits depend on every bit of h.
// Composite (key, 4-byte second key) rows: the second key is folded into the position hash and stored in the
// slot's spare word, so (int64,int32) / (int32,int32) keys take the same partitioned path as single keys.
static __device__ __forceinline__ uint32_t with_k2(uint32_t h, uint32_t k2) {
  h = (h ^ (k2 * 0x9E3779B1u)) * 0x85EBCA77u;
  return h ^ (h >> 15);
}

// Home slots are EVEN: a probe reads the aligned 32-byte pair {home, home+1} in one sector.
static __device__ __forceinline__ uint32_t slot_hash(uint32_t h) { return ((h * 0x9E3779B1u) >> 7) & ~1u; }

struct PartGeom {
  unsigned nparts;  // power of two (local radix partitions) or any count (dest modes)
  unsigned shift;   // local partition = hash >> shift   (shift = 32 - log2(#local partitions)); one partition -> 0
  unsigned dest;    // 1 = "destination rank" mode of the multi-GPU layer: pid = mulhi(remix(hash), nparts), a
                    //     function that is independent of the top bits the receiver's local partitioning uses
                    // 2 = combined mode (fused exchange): pid = destination rank * nlocal + the receiver's local
                    //     partition, so that ONE pass on the sender does the work of both partition passes
  unsigned nlocal;  // dest == 2: local partitions per destination (power of two); nparts = ranks * nlocal
  __device__ __forceinline__ unsigned pid(uint32_t h) const {
    if (dest == 2u) return __umulhi(fmix32(h ^ 0x5bd1e995u), nparts / nlocal) * nlocal + (nlocal == 1 ? 0u : (h >> shift));
    if (dest & 1u) return __umulhi(fmix32(h ^ 0x5bd1e995u), nparts);
    return nparts == 1 ? 0u : (h >> shift);
  }
};

// ---- pass 1: per-partition row counts.  Shared-memory 32-bit atomics are the cheapest primitive on
// B200 for this (~2900 Gop/s measured, profiles/r02_microbench.txt; match.any-based ranking measured
// 20x slower), so the CTA histogram is plain atomicAdd on shared counters. ----
// COMPACT (join_compact.cuh): rows whose 64-bit key has a non-zero high word cannot match a 32-bit build side and
// are counted like NULL-key rows.  hi_or (if not null) receives the OR of the high words of all valid keys - the
// build-side launch uses it to find out whether the compact path applies at all.
// VECTOR = the mask-free single-key case (128-bit loads); a separate instantiation so that the registers of the general
// row loop (keys, second keys and validity bytes of eight rows) do not lower the occupancy of the C3 path.
template <typename KT, bool KEEP_NULLS, bool COMPACT, bool VECTOR>
__global__ void __launch_bounds__(kThreads)
part_hist_kernel(const KT* __restrict__ keys, const gdf_valid_type* __restrict__ valid, size_t n, PartGeom g,
                 unsigned long long* __restrict__ totals, const uint32_t* __restrict__ k2,
                 const gdf_valid_type* __restrict__ valid2, unsigned* __restrict__ hi_or) {
  __shared__ unsigned hist[kMaxParts];
  for (unsigned p = threadIdx.x; p < g.nparts; p += kThreads) hist[p] = 0;
  __syncthreads();
  const size_t stride = (size_t)gridDim.x * kThreads;
  uint32_t hi_acc = 0;
  auto wide = [&](KT k) -> bool {  // high word of an 8-byte key (always false for 4-byte keys)
    if (sizeof(KT) != 8) return false;
    const uint32_t hi = (uint32_t)((unsigned long long)k >> 32);
    hi_acc |= hi;
    return COMPACT && hi != 0;
  };
  if (VECTOR) {  // no mask, no second key, 16-byte aligned (checked by the host): 128-bit loads, 4 in flight per thread
    constexpr int VEC = 16 / (int)sizeof(KT), U = 4;
    const size_t nvec = n / VEC;
    const uint4* k4 = reinterpret_cast<const uint4*>(keys);
    for (size_t v0 = (size_t)blockIdx.x * kThreads + threadIdx.x; v0 < nvec; v0 += stride * U) {
      uint4 raw[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const size_t v = v0 + (size_t)u * stride;
        raw[u] = v < nvec ? ldg_stream(k4 + v) : make_uint4(0, 0, 0, 0);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const size_t v = v0 + (size_t)u * stride;
        if (v >= nvec) continue;
        const KT* e = reinterpret_cast<const KT*>(&raw[u]);
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
          if (!wide(e[j])) atomicAdd(&hist[g.pid(KeyBits<KT>::hash(e[j]))], 1u);
          else if (KEEP_NULLS) atomicAdd(&hist[(unsigned)(v * VEC + j) & (g.nparts - 1)], 1u);
        }
      }
    }
    if (blockIdx.x == 0 && threadIdx.x < VEC) {  // ragged tail
      const size_t r = nvec * VEC + threadIdx.x;
      if (r < n) {
        if (!wide(keys[r])) atomicAdd(&hist[g.pid(KeyBits<KT>::hash(keys[r]))], 1u);
        else if (KEEP_NULLS) atomicAdd(&hist[(unsigned)r & (g.nparts - 1)], 1u);
      }
    }
  } else {
    // Everything a row needs - key, second key, the two validity bytes - is loaded for all U rows BEFORE the first
    // shared-memory atomic: the compiler keeps loads behind an atomic, so fetching them row by row serialised eight
    // dependent round trips per thread (C5: 5.4 ms for the two histogram launches, 4x the bytes' worth).
    constexpr int U = 8;
    for (size_t r0 = (size_t)blockIdx.x * kThreads + threadIdx.x; r0 < n; r0 += stride * U) {
      KT k[U];
      uint32_t kk2[U];
      unsigned char v1[U], v2[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const size_t r = r0 + (size_t)u * stride;
        const bool in = r < n;
        k[u] = in ? keys[r] : (KT)0;
        kk2[u] = (in && k2) ? k2[r] : 0u;
        v1[u] = (in && valid) ? valid[r >> 3] : (unsigned char)0xff;
        v2[u] = (in && valid2) ? valid2[r >> 3] : (unsigned char)0xff;
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const size_t r = r0 + (size_t)u * stride;
        if (r >= n) continue;
        const bool ok = ((v1[u] >> (r & 7)) & 1u) && ((v2[u] >> (r & 7)) & 1u);
        if (ok && !wide(k[u])) {
          uint32_t h = KeyBits<KT>::hash(k[u]);
          if (k2) h = with_k2(h, kk2[u]);
          atomicAdd(&hist[g.pid(h)], 1u);
        } else if (KEEP_NULLS) {
          atomicAdd(&hist[(unsigned)r & (g.nparts - 1)], 1u);
        }
      }
    }
  }
  __syncthreads();
  for (unsigned p = threadIdx.x; p < g.nparts; p += kThreads)
    if (hist[p]) atomicAdd(&totals[p], (unsigned long long)hist[p]);
  if (hi_or != nullptr && sizeof(KT) == 8) {
    hi_acc = __reduce_or_sync(0xffffffffu, hi_acc);
    if (lane_id() == 0 && hi_acc) atomicOr(hi_or, hi_acc);
  }
}

// ---- pass 2: write-combining scatter of {key,row} pairs ----
// v2 (profiles/r01a: v1 reached 3.8 TB/s of DRAM traffic with six barriers per 4096-row tile, 24 warps
// per SM and the tile's loads issued only after the previous tile had been written):
//   * the NEXT tile's keys are loaded into registers before the current tile is ranked, so DRAM
//     latency overlaps the shared-memory phases;
//   * three barriers per tile: rank (shared-memory atomics) | reserve runs (one global atomic per
//     non-empty partition, issued early and only consumed after the staging phase) + scan | stage |
//     copy-out;
//   * copy-out is warp-per-partition: a warp copies one partition's run of the staged tile to its
//     reserved global range with consecutive lanes on consecutive elements - no per-element
//     partition lookup, coalesced stores.
template <typename KT>
struct ScatterSmem {
  KT keys[kScatterTile];
  int32_t rows[kScatterTile];
  unsigned short pid[kScatterTile];          // partition of every staged element (element-parallel copy-out)
  unsigned hist[2][kMaxParts];               // rows of this tile per partition (double-buffered)
  unsigned lstart[kMaxParts];                // start of partition p inside the staged tile
  unsigned long long gbase[kMaxParts];       // reserved global start of this tile's run
};

// Fused partition + exchange (multi-GPU layer): when `peer.on`, partition p's rows are written straight into
// destination p's receive buffers - peer-GPU memory mapped through CUDA IPC, i.e. the stores travel over
// NVLink from inside this kernel - instead of into one local array that an all-to-all would then move.
constexpr int kMaxPeers = 16;
struct PeerDst {
  void* keys[kMaxPeers];
  int32_t* ids[kMaxPeers];
  int on;
};

struct ScatterK2 {  // second key column of composite keys (HAS_K2)
  const uint32_t* in;
  const gdf_valid_type* valid2;
  uint32_t* out;
};

template <typename KT, bool KEEP_NULLS, bool PREFETCH, bool HAS_K2>
__global__ void __launch_bounds__(kThreads, (PREFETCH || HAS_K2) ? 2 : 3)
part_scatter_kernel(const KT* __restrict__ keys, const gdf_valid_type* __restrict__ valid, size_t n, PartGeom g,
                    unsigned long long* __restrict__ cursors, KT* __restrict__ out_keys,
                    int32_t* __restrict__ out_rows, const int32_t* __restrict__ payload, int32_t id_base,
                    const PeerDst peer, const ScatterK2 k2) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  ScatterSmem<KT>& sm = *reinterpret_cast<ScatterSmem<KT>*>(smem_raw);
  uint32_t* const sk2 = reinterpret_cast<uint32_t*>(smem_raw + sizeof(ScatterSmem<KT>));  // [kScatterTile], HAS_K2 only
  const unsigned warp = threadIdx.x >> 5, lane = lane_id();
  const size_t tiles = (n + kScatterTile - 1) / kScatterTile;
  for (unsigned p = threadIdx.x; p < 2 * kMaxParts; p += kThreads) (&sm.hist[0][0])[p] = 0;
  __syncthreads();
  KT knext[kScatterRows];
  // a warp owns 32*kScatterRows consecutive rows; step i covers 32 consecutive rows (coalesced)
  auto load_tile = [&](size_t tile) {
    const size_t wbase = tile * kScatterTile + (size_t)warp * (32 * kScatterRows);
#pragma unroll
    for (int i = 0; i < kScatterRows; ++i) {
      const size_t r = wbase + (size_t)i * 32 + lane;
      knext[i] = (tile < tiles && r < n) ? keys[r] : (KT)0;
    }
  };
  if (PREFETCH) load_tile(blockIdx.x);
  unsigned buf = 0;
  for (size_t tile = blockIdx.x; tile < tiles; tile += gridDim.x, buf ^= 1u) {
    const size_t wbase = tile * kScatterTile + (size_t)warp * (32 * kScatterRows);
    KT k[kScatterRows];
    if (!PREFETCH) load_tile(tile);
#pragma unroll
    for (int i = 0; i < kScatterRows; ++i) k[i] = knext[i];
    if (PREFETCH) load_tile(tile + gridDim.x);  // in flight during all the shared-memory phases below
    unsigned rp[kScatterRows];    // rank << 16 | pid, 0xffff = dropped, bit 15 = NULL-key row
    uint32_t k2v[HAS_K2 ? kScatterRows : 1];
#pragma unroll
    for (int i = 0; i < kScatterRows; ++i) {
      const size_t r = wbase + (size_t)i * 32 + lane;
      bool keep = r < n;
      unsigned p = 0, nullbit = 0;
      if (HAS_K2) k2v[i] = keep ? k2.in[r] : 0u;
      if (keep) {
        if (bit_valid(valid, r) && (!HAS_K2 || bit_valid(k2.valid2, r))) {
          uint32_t h = KeyBits<KT>::hash(k[i]);
          if (HAS_K2) h = with_k2(h, k2v[i]);
          p = g.pid(h);
        }
        else if (KEEP_NULLS) { p = (unsigned)r & (g.nparts - 1); nullbit = 0x8000u; }
        else keep = false;
      }
      rp[i] = keep ? ((atomicAdd(&sm.hist[buf][p], 1u) << 16) | p | nullbit) : 0xffffu;
    }
    __syncthreads();  // (1) tile histogram complete
    {                 // reserve this tile's runs: result lands in gbase[] before barrier (3)
      const unsigned p = threadIdx.x;
      if (p < g.nparts) {
        const unsigned run = sm.hist[buf][p];
        sm.gbase[p] = run ? atomicAdd(&cursors[p], (unsigned long long)run) : 0ull;
      }
      sm.hist[buf ^ 1u][threadIdx.x] = 0;  // kThreads == kMaxParts: clear the other buffer for the next tile
    }
    if (warp == kThreads / 32 - 1) {  // exclusive scan of the <= 256 partition counts (8 per lane)
      unsigned c[kMaxParts / 32], tot = 0;
#pragma unroll
      for (int j = 0; j < (int)(kMaxParts / 32); ++j) {
        const unsigned p = lane * (kMaxParts / 32) + j;
        c[j] = p < g.nparts ? sm.hist[buf][p] : 0;
        tot += c[j];
      }
      unsigned inc = tot;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const unsigned o = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= (unsigned)d) inc += o;
      }
      unsigned run = inc - tot;
#pragma unroll
      for (int j = 0; j < (int)(kMaxParts / 32); ++j) {
        const unsigned p = lane * (kMaxParts / 32) + j;
        if (p < g.nparts) sm.lstart[p] = run;
        run += c[j];
      }
    }
    __syncthreads();  // (2) lstart visible (previous tile's copy-out finished before barrier (1))
#pragma unroll
    for (int i = 0; i < kScatterRows; ++i) {
      if ((rp[i] & 0xffffu) == 0xffffu) continue;
      const unsigned p = rp[i] & 0x7fffu;
      const unsigned at = sm.lstart[p] + (rp[i] >> 16);
      const size_t row = wbase + (size_t)i * 32 + lane;
      // row tag: position in the column, or the caller's payload (global row id of an exchanged row)
      const int32_t r = payload ? payload[row] : (int32_t)row + id_base;
      sm.keys[at] = k[i];
      sm.rows[at] = (rp[i] & 0x8000u) ? ~r : r;
      sm.pid[at] = (unsigned short)p;
      if (HAS_K2) sk2[at] = k2v[i];
    }
    __syncthreads();  // (3) staged tile + gbase complete
    {  // element-parallel copy-out: consecutive threads take consecutive staged elements, which are
       // consecutive in the output wherever they belong to the same partition
      const unsigned kept = sm.lstart[g.nparts - 1] + sm.hist[buf][g.nparts - 1];
      for (unsigned j = threadIdx.x; j < kept; j += kThreads) {
        const unsigned p = sm.pid[j];
        const unsigned long long gidx = sm.gbase[p] + (j - sm.lstart[p]);
        KT* const ok = peer.on ? static_cast<KT*>(peer.keys[p]) : out_keys;
        int32_t* const orow = peer.on ? peer.ids[p] : out_rows;
        ok[gidx] = sm.keys[j];
        orow[gidx] = sm.rows[j];
        if (HAS_K2) k2.out[gidx] = sk2[j];
      }
    }
    // the next iteration's barrier (1) separates this copy-out from the next staging phase; the
    // histogram of this tile (hist[buf]) is cleared during the NEXT tile's phase (2)
  }
}

// ---- compact scatter: {key32, tag32} pairs in ONE 8-byte array (join_compact.cuh) ----
// Same tile protocol as part_scatter_kernel; the staged tile is 32 KB of pairs + a partition byte per element, so
// four 256-thread CTAs are resident per SM instead of three (128 KB of input in flight per SM instead of 96).  Rows with a NULL key or a
// key that does not fit 32 bits are dropped (build side, INNER probe side) or kept with tag = ~row (LEFT / FULL).
constexpr int kS32Threads = 256;
constexpr int kS32Rows = 16;
constexpr int kS32Tile = kS32Threads * kS32Rows;  // 4096

struct Scatter32Smem {
  uint2 pairs[kS32Tile];
  unsigned char pid[kS32Tile];
  unsigned hist[2][kMaxParts];
  unsigned lstart[kMaxParts];
  unsigned long long gbase[kMaxParts];
};

// Fused exchange (multi-GPU layer, combined geometry): bin = destination rank * nlocal + local partition, and the
// bin's run goes straight into THAT rank's pair buffer - peer memory mapped through CUDA IPC, so the stores cross
// NVLink from inside this kernel and arrive already in the receiver's partition order.
struct PeerPairs {
  uint2* base[kMaxPeers];
  unsigned on;      // 0: everything goes to out_pairs
  unsigned shift;   // destination rank = bin >> shift  (shift = log2(nlocal))
  const int* status;  // asynchronous exchange: device flags written by xjoin_plan_kernel; non-zero = write nothing
  // histogram-free partitioning (single-GPU INNER probe side): bin p owns the fixed region [p * region_cap,
  // (p + 1) * region_cap) of the destination, cursors start at the region starts; a run that would cross its region's
  // end is dropped and *overflow is raised (the caller then repeats the side with exact counts)
  unsigned long long region_cap;
  int* overflow;
  // multi-GPU exchange (part_scatter32_even_kernel): this rank's rows per bin and its start offset inside the
  // destinations' buffers (device arrays); every (sender, bin) slot there holds even(count) pairs
  const unsigned long long* counts;
  const unsigned long long* offsets0;
};

template <typename KT, bool KEEP_NULLS>
__global__ void __launch_bounds__(kS32Threads, sizeof(KT) == 8 ? 3 : 4)
part_scatter32_kernel(const KT* __restrict__ keys, const gdf_valid_type* __restrict__ valid, size_t n, PartGeom g,
                      unsigned long long* __restrict__ cursors, uint2* __restrict__ out_pairs,
                      const int32_t* __restrict__ payload, int32_t id_base, const PeerPairs peer) {
  extern __shared__ __align__(16) unsigned char scatter32_smem[];
  Scatter32Smem& sm = *reinterpret_cast<Scatter32Smem*>(scatter32_smem);
  if (peer.status != nullptr && (peer.status[0] | peer.status[1]) != 0) return;  // the device-side plan said "do not write"

  const unsigned warp = threadIdx.x >> 5, lane = lane_id();
  const size_t tiles = (n + kS32Tile - 1) / kS32Tile;
  for (unsigned p = threadIdx.x; p < 2 * kMaxParts; p += kS32Threads) (&sm.hist[0][0])[p] = 0;
  __syncthreads();
  unsigned buf = 0;
  for (size_t tile = blockIdx.x; tile < tiles; tile += gridDim.x, buf ^= 1u) {
    const size_t wbase = tile * kS32Tile + (size_t)warp * (32 * kS32Rows);
    uint32_t k[kS32Rows];  // low key words; bit i of `wide` = row i's key does not fit 32 bits
    unsigned wide = 0;
#pragma unroll
    for (int i = 0; i < kS32Rows; ++i) {
      const size_t r = wbase + (size_t)i * 32 + lane;
      const KT kk = r < n ? keys[r] : (KT)0;
      k[i] = (uint32_t)kk;
      if (sizeof(KT) == 8 && (uint32_t)((unsigned long long)kk >> 32) != 0) wide |= 1u << i;
    }
    unsigned rp[kS32Rows];  // rank << 16 | pid, 0xffff = dropped, bit 15 = row that can never match
#pragma unroll
    for (int i = 0; i < kS32Rows; ++i) {
      const size_t r = wbase + (size_t)i * 32 + lane;
      bool keep = r < n;
      unsigned p = 0, nullbit = 0;
      if (keep) {
        if (!((wide >> i) & 1u) && bit_valid(valid, r)) p = g.pid(KeyBits<uint32_t>::hash(k[i]));
        else if (KEEP_NULLS) { p = (unsigned)r & (g.nparts - 1); nullbit = 0x8000u; }
        else keep = false;
      }
      rp[i] = keep ? ((atomicAdd(&sm.hist[buf][p], 1u) << 16) | p | nullbit) : 0xffffu;
    }
    __syncthreads();  // (1) tile histogram complete
    {
      const unsigned p = threadIdx.x;
      if (p < g.nparts) {
        const unsigned run = sm.hist[buf][p];
        sm.gbase[p] = run ? atomicAdd(&cursors[p], (unsigned long long)run) : 0ull;
      }
      if (p < kMaxParts) sm.hist[buf ^ 1u][p] = 0;  // the other buffer, for the next tile
    }
    if (warp == kS32Threads / 32 - 1) {  // exclusive scan of the <= 256 partition counts (8 per lane)
      unsigned c[kMaxParts / 32], tot = 0;
#pragma unroll
      for (int j = 0; j < (int)(kMaxParts / 32); ++j) {
        const unsigned p = lane * (kMaxParts / 32) + j;
        c[j] = p < g.nparts ? sm.hist[buf][p] : 0;
        tot += c[j];
      }
      unsigned inc = tot;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const unsigned o = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= (unsigned)d) inc += o;
      }
      unsigned run = inc - tot;
#pragma unroll
      for (int j = 0; j < (int)(kMaxParts / 32); ++j) {
        const unsigned p = lane * (kMaxParts / 32) + j;
        if (p < g.nparts) sm.lstart[p] = run;
        run += c[j];
      }
    }
    __syncthreads();  // (2) lstart visible (the previous tile's copy-out finished before barrier (1))
#pragma unroll
    for (int i = 0; i < kS32Rows; ++i) {
      if ((rp[i] & 0xffffu) == 0xffffu) continue;
      const unsigned p = rp[i] & 0x7fffu;
      const unsigned at = sm.lstart[p] + (rp[i] >> 16);
      const size_t row = wbase + (size_t)i * 32 + lane;
      const int32_t r = payload ? payload[row] : (int32_t)row + id_base;
      sm.pairs[at] = make_uint2(k[i], (uint32_t)((rp[i] & 0x8000u) ? ~r : r));
      sm.pid[at] = (unsigned char)p;
    }
    __syncthreads();  // (3) staged tile + gbase complete
    const unsigned kept = sm.lstart[g.nparts - 1] + sm.hist[buf][g.nparts - 1];
    for (unsigned j = threadIdx.x; j < kept; j += kS32Threads) {
      const unsigned p = sm.pid[j];
      uint2* const dst = peer.on ? peer.base[p >> peer.shift] : out_pairs;
      dst[sm.gbase[p] + (j - sm.lstart[p])] = sm.pairs[j];
    }
  }
}

// ---- compact scatter for the fused exchange: per-bin runs leave the SM as TMA bulk stores ----
// tools/p2p_bench.cu (profiles/r02_p2p_store_bench.txt): the element-parallel copy-out above - a warp instruction stores
// 32 consecutive staged pairs, which straddle bin boundaries and start at arbitrary 8-byte offsets - reaches 446 GB/s
// of peer stores over NVLink (the SM's outstanding store requests carry ~70 useful bytes each), exactly what the
// exchange measured; cp.async.bulk shared -> global reaches 716 GB/s (the copy-engine rate) for runs as short as
// 128 bytes, because the TMA engine emits whole lines.  So here ONE thread per bin issues ONE bulk store per tile.
// Bulk copies need 16-byte aligned addresses on both sides and a multiple of 16 bytes; pairs are 8 bytes, so:
//   * the bin's reserved global position (gbase, from the cursor atomic) must be known BEFORE the tile is staged: its
//     parity decides where the run sits in shared memory (slot start even, run at slot start + parity), so that the
//     even-aligned body of the run is 16-byte aligned in shared AND global memory; an odd head / tail pair is one
//     plain 8-byte store.  The atomic's latency is exposed once per tile - immaterial here, the kernel is NVLink-bound;
//   * a slot is the run rounded up to an even length (+ parity), so the staged tile needs up to 2 extra pairs per bin.
struct Scatter32BulkSmem {
  uint2 pairs[kS32Tile + 2 * kMaxParts];
  unsigned hist[2][kMaxParts];
  unsigned lstart[kMaxParts];
  unsigned long long gbase[kMaxParts];
};

template <typename KT>
__global__ void __launch_bounds__(kS32Threads, 3)
part_scatter32_bulk_kernel(const KT* __restrict__ keys, size_t n, PartGeom g, unsigned long long* __restrict__ cursors,
                           int32_t id_base, const PeerPairs peer) {
  extern __shared__ __align__(128) unsigned char scatter32b_smem[];
  Scatter32BulkSmem& sm = *reinterpret_cast<Scatter32BulkSmem*>(scatter32b_smem);
  if (peer.status != nullptr && (peer.status[0] | peer.status[1]) != 0) return;  // the device-side plan said "do not write"
  const unsigned warp = threadIdx.x >> 5, lane = lane_id();
  const size_t tiles = (n + kS32Tile - 1) / kS32Tile;
  for (unsigned p = threadIdx.x; p < 2 * kMaxParts; p += kS32Threads) (&sm.hist[0][0])[p] = 0;
  __syncthreads();
  unsigned buf = 0;
  for (size_t tile = blockIdx.x; tile < tiles; tile += gridDim.x, buf ^= 1u) {
    const size_t wbase = tile * kS32Tile + (size_t)warp * (32 * kS32Rows);
    uint32_t k[kS32Rows];
    unsigned wide = 0;
#pragma unroll
    for (int i = 0; i < kS32Rows; ++i) {
      const size_t r = wbase + (size_t)i * 32 + lane;
      const KT kk = r < n ? keys[r] : (KT)0;
      k[i] = (uint32_t)kk;
      if (sizeof(KT) == 8 && (uint32_t)((unsigned long long)kk >> 32) != 0) wide |= 1u << i;
    }
    unsigned rp[kS32Rows];  // rank << 16 | bin, 0xffff = dropped (past the end / key wider than 32 bits: it cannot match)
#pragma unroll
    for (int i = 0; i < kS32Rows; ++i) {
      const size_t r = wbase + (size_t)i * 32 + lane;
      const bool keep = r < n && !((wide >> i) & 1u);
      const unsigned p = keep ? g.pid(KeyBits<uint32_t>::hash(k[i])) : 0u;
      rp[i] = keep ? ((atomicAdd(&sm.hist[buf][p], 1u) << 16) | p) : 0xffffu;
    }
    __syncthreads();  // (1) tile histogram complete
    {
      const unsigned p = threadIdx.x;
      if (p < g.nparts) {
        const unsigned run = sm.hist[buf][p];
        unsigned long long at = run ? atomicAdd(&cursors[p], (unsigned long long)run) : 0ull;
        if (peer.region_cap && run && at + run > (unsigned long long)(p + 1) * peer.region_cap) {
          *peer.overflow = 1;
          at = ~0ull;  // dropped
        }
        sm.gbase[p] = at;
        tma::bulk_wait_read_all();  // this thread's bulk store of the previous tile has finished reading shared memory
      }
      if (p < kMaxParts) sm.hist[buf ^ 1u][p] = 0;  // the other buffer, for the next tile
    }
    __syncthreads();  // (2a) reserved positions known, the previous tile's staged pairs are free
    if (warp == kS32Threads / 32 - 1) {  // slot starts: exclusive scan of the even slot lengths (8 bins per lane)
      unsigned c[kMaxParts / 32], tot = 0;
#pragma unroll
      for (int j = 0; j < (int)(kMaxParts / 32); ++j) {
        const unsigned p = lane * (kMaxParts / 32) + j;
        c[j] = p < g.nparts ? ((sm.hist[buf][p] + ((unsigned)sm.gbase[p] & 1u) + 1u) & ~1u) : 0;
        tot += c[j];
      }
      unsigned inc = tot;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const unsigned o = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= (unsigned)d) inc += o;
      }
      unsigned run = inc - tot;
#pragma unroll
      for (int j = 0; j < (int)(kMaxParts / 32); ++j) {
        const unsigned p = lane * (kMaxParts / 32) + j;
        if (p < g.nparts) sm.lstart[p] = run + ((unsigned)sm.gbase[p] & 1u);  // the run starts at the slot + parity
        run += c[j];
      }
    }
    __syncthreads();  // (2b) run starts visible
#pragma unroll
    for (int i = 0; i < kS32Rows; ++i) {
      if ((rp[i] & 0xffffu) == 0xffffu) continue;
      const unsigned p = rp[i] & 0xffffu;
      const size_t row = wbase + (size_t)i * 32 + lane;
      sm.pairs[sm.lstart[p] + (rp[i] >> 16)] = make_uint2(k[i], (uint32_t)((int32_t)row + id_base));
    }
    tma::fence_proxy_async();  // the staged pairs are read by the copy engine
    __syncthreads();           // (3) staged tile complete
    {
      const unsigned p = threadIdx.x;
      unsigned len = p < g.nparts ? sm.hist[buf][p] : 0u;
      if (len && sm.gbase[p] != ~0ull) {
        uint2* dst = peer.base[p >> peer.shift] + sm.gbase[p];
        const uint2* src = &sm.pairs[sm.lstart[p]];
        if ((unsigned)sm.gbase[p] & 1u) {  // odd head pair: plain store; what follows is 16-byte aligned on both sides
          *dst = *src;
          ++dst, ++src, --len;
        }
        const unsigned body = len & ~1u;
        if (body) tma::bulk_store(dst, src, body * (unsigned)sizeof(uint2));
        if (len & 1u) dst[body] = src[body];
        tma::bulk_commit();
      }
    }
  }
  tma::bulk_wait_all();
}

// ---- the exchange's scatter: SECTOR-ALIGNED runs only ----
// tools/p2p_bench.cu (profiles/r02_p2p_store_bench.txt): NVLink is bound by the NUMBER of write packets, not by bytes.
// One plain 8-byte peer store next to every 256-byte bulk store takes it from 695 to 487 GB/s, two take it to 394 (the
// odd head / tail pairs of part_scatter32_bulk_kernel were exactly such packets: the exchange measured 0.42 TB/s); a bulk
// store that starts or ends inside a 32-byte sector costs too (256-byte runs: 700 GB/s aligned, 618 at 32-byte, 497 at
// 16-byte alignment).  Here nothing but whole-sector bulk stores leaves the SM in the steady state:
//   * the receiver's layout gives every (sender, bin) slot a length that is a multiple of kXG = 4 pairs = 32 bytes (the
//     plan rounds counts up, dist.py plan_fused_exchange / xjoin_plan_kernel), so a sender's cursor for a bin starts
//     sector-aligned and - because every tile reserves a multiple of kXG pairs - stays so;
//   * the up to kXG - 1 pairs of a bin that a tile cannot emit are CARRIED in shared memory into the CTA's next tile
//     (staged first there);
//   * what is still carried when the CTA runs out of tiles is written with single stores (<= 3 per bin and CTA, once,
//     filling the slot from its back so that the front cursor other CTAs still use stays aligned), and CTA 0 fills the
//     rest of every slot with {0, INT_MIN} "no row" pairs, which build and probe skip.
constexpr unsigned kXG = 4;   // pairs per exchange granule (one 32-byte sector); dist.py EXCHANGE_GRANULE must agree

struct Scatter32EvenSmem {
  uint2 pairs[kS32Tile + 2 * (kXG - 1) * kMaxParts];
  uint2 carry[kMaxParts][kXG - 1];
  unsigned carried[kMaxParts];   // 0 .. kXG - 1
  unsigned hist[2][kMaxParts];
  unsigned lstart[kMaxParts];
  unsigned long long gbase[kMaxParts];
};

template <typename KT>
__global__ void __launch_bounds__(kS32Threads, 3)
part_scatter32_even_kernel(const KT* __restrict__ keys, size_t n, PartGeom g, unsigned long long* __restrict__ cursors,
                           int32_t id_base, const PeerPairs peer, const unsigned long long* __restrict__ counts,
                           const unsigned long long* __restrict__ offsets0, unsigned* __restrict__ tail_cursors) {
  extern __shared__ __align__(128) unsigned char scatter32e_smem[];
  Scatter32EvenSmem& sm = *reinterpret_cast<Scatter32EvenSmem*>(scatter32e_smem);
  if (peer.status != nullptr && (peer.status[0] | peer.status[1]) != 0) return;  // the device-side plan said "do not write"
  const unsigned warp = threadIdx.x >> 5, lane = lane_id();
  const size_t tiles = (n + kS32Tile - 1) / kS32Tile;
  for (unsigned p = threadIdx.x; p < 2 * kMaxParts; p += kS32Threads) (&sm.hist[0][0])[p] = 0;
  for (unsigned p = threadIdx.x; p < kMaxParts; p += kS32Threads) sm.carried[p] = 0;
  if (blockIdx.x == 0)  // pad every slot up to a whole granule: "no row" pairs behind this rank's pairs of the bin
    for (unsigned p = threadIdx.x; p < g.nparts; p += kS32Threads) {
      const unsigned long long c = counts[p], end = (c + kXG - 1) & ~(unsigned long long)(kXG - 1);
      for (unsigned long long i = c; i < end; ++i) peer.base[p >> peer.shift][offsets0[p] + i] = make_uint2(0u, 0x80000000u);
    }
  __syncthreads();
  unsigned buf = 0;
  for (size_t tile = blockIdx.x; tile < tiles; tile += gridDim.x, buf ^= 1u) {
    const size_t wbase = tile * kS32Tile + (size_t)warp * (32 * kS32Rows);
    uint32_t k[kS32Rows];
    unsigned wide = 0;
#pragma unroll
    for (int i = 0; i < kS32Rows; ++i) {
      const size_t r = wbase + (size_t)i * 32 + lane;
      const KT kk = r < n ? keys[r] : (KT)0;
      k[i] = (uint32_t)kk;
      if (sizeof(KT) == 8 && (uint32_t)((unsigned long long)kk >> 32) != 0) wide |= 1u << i;
    }
    unsigned rp[kS32Rows];  // rank << 16 | bin, 0xffff = dropped (past the end / key wider than 32 bits: it cannot match)
#pragma unroll
    for (int i = 0; i < kS32Rows; ++i) {
      const size_t r = wbase + (size_t)i * 32 + lane;
      const bool keep = r < n && !((wide >> i) & 1u);
      const unsigned p = keep ? g.pid(KeyBits<uint32_t>::hash(k[i])) : 0u;
      rp[i] = keep ? ((atomicAdd(&sm.hist[buf][p], 1u) << 16) | p) : 0xffffu;
    }
    __syncthreads();  // (1) tile histogram complete
    {
      const unsigned p = threadIdx.x;
      if (p < g.nparts) {
        const unsigned have = sm.hist[buf][p] + sm.carried[p];   // the carried pairs are staged first
        const unsigned emit = have & ~(kXG - 1);
        sm.gbase[p] = emit ? atomicAdd(&cursors[p], (unsigned long long)emit) : 0ull;
        tma::bulk_wait_read_all();  // this thread's bulk store of the previous tile has finished reading shared memory
      }
      if (p < kMaxParts) sm.hist[buf ^ 1u][p] = 0;  // the other buffer, for the next tile
    }
    __syncthreads();  // (2a) reserved positions known, the previous tile's staged pairs are free
    if (warp == kS32Threads / 32 - 1) {  // slot starts: exclusive scan of the slot lengths rounded up to granules (8 bins per lane)
      unsigned c[kMaxParts / 32], tot = 0;
#pragma unroll
      for (int j = 0; j < (int)(kMaxParts / 32); ++j) {
        const unsigned p = lane * (kMaxParts / 32) + j;
        c[j] = p < g.nparts ? ((sm.hist[buf][p] + sm.carried[p] + kXG - 1) & ~(kXG - 1)) : 0;
        tot += c[j];
      }
      unsigned inc = tot;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const unsigned o = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= (unsigned)d) inc += o;
      }
      unsigned run = inc - tot;
#pragma unroll
      for (int j = 0; j < (int)(kMaxParts / 32); ++j) {
        const unsigned p = lane * (kMaxParts / 32) + j;
        if (p < g.nparts) sm.lstart[p] = run;
        run += c[j];
      }
    }
    __syncthreads();  // (2b) slot starts visible
    if (threadIdx.x < g.nparts)
      for (unsigned i = 0; i < sm.carried[threadIdx.x]; ++i) sm.pairs[sm.lstart[threadIdx.x] + i] = sm.carry[threadIdx.x][i];
#pragma unroll
    for (int i = 0; i < kS32Rows; ++i) {
      if ((rp[i] & 0xffffu) == 0xffffu) continue;
      const unsigned p = rp[i] & 0xffffu;
      const size_t row = wbase + (size_t)i * 32 + lane;
      sm.pairs[sm.lstart[p] + sm.carried[p] + (rp[i] >> 16)] = make_uint2(k[i], (uint32_t)((int32_t)row + id_base));
    }
    tma::fence_proxy_async();  // the staged pairs are read by the copy engine
    __syncthreads();           // (3) staged tile complete
    {
      const unsigned p = threadIdx.x;
      if (p < g.nparts) {
        const unsigned have = sm.hist[buf][p] + sm.carried[p];
        const unsigned emit = have & ~(kXG - 1), rest = have & (kXG - 1);
        const uint2* src = &sm.pairs[sm.lstart[p]];
        if (emit) {
          tma::bulk_store(peer.base[p >> peer.shift] + sm.gbase[p], src, emit * (unsigned)sizeof(uint2));
          tma::bulk_commit();
        }
        for (unsigned i = 0; i < rest; ++i) sm.carry[p][i] = src[emit + i];  // not part of the bulk store's range
        sm.carried[p] = rest;
      }
    }
    // the next tile's barriers order these writes of carry / carried before their next readers
  }
  __syncthreads();
  // What is still carried: single stores, <= kXG - 1 per bin and CTA.  They fill the slot from its BACK (just before the
  // pad pairs), so that the front cursor - still used by other CTAs' bulk stores - stays sector-aligned; the runs and
  // the singles meet exactly, because the slot holds the count rounded up to a granule.
  if (threadIdx.x < g.nparts) {
    const unsigned p = threadIdx.x, rest = sm.carried[p];
    if (rest) {
      const unsigned long long last_real = offsets0[p] + counts[p] - 1ull;   // the pads sit right behind it
      const unsigned at = atomicAdd(&tail_cursors[p], rest);
      for (unsigned i = 0; i < rest; ++i) peer.base[p >> peer.shift][last_real - (at + i)] = sm.carry[p][i];
    }
  }
  tma::bulk_wait_all();
}

#include "join_compact.cuh"

// Where the pairs of one side live.  rows == nullptr means "not partitioned": key i belongs to row i
// and `valid` (if any) still applies.
template <typename KT>
struct Pairs {
  const KT* keys;
  const int32_t* rows;
  const gdf_valid_type* valid;
  size_t n;
  const uint32_t* k2;             // second key of a composite key (nullptr: single key)
  const gdf_valid_type* valid2;   // its mask (unpartitioned pairs only)
};

struct Tables {
  Slot* slots;
  const unsigned long long* offset;  // [nparts] first slot of partition p
  const unsigned* mask;              // [nparts] slots_p - 1
};

template <typename KT, bool K2>   // K2: composite key, second key in Pairs::k2 (kept out of the single-key code)
__global__ void __launch_bounds__(kThreads)
part_build_kernel(Pairs<KT> b, PartGeom g, Tables t, int* __restrict__ flags /*[0]=dup [1]=sentinel key*/) {
  // One CTA = one contiguous tile of pairs, tiles dispatched in index order: the pairs are
  // partition-contiguous, so the CTAs in flight insert into one or two partitions' tables at a time
  // and those tables stay L2-resident (a grid-stride loop would touch every table at once).
  // Each thread keeps kBuildU inserts in flight: the first CAS of every row is issued before any result
  // is looked at (the v1 kernel did one dependent CAS round trip per row and sat at 12 % issue
  // utilisation with 104 warps-per-issue stalled on the scoreboard).
  constexpr int U = kBuildTile / kThreads;
  const size_t tile_lo = (size_t)blockIdx.x * kBuildTile;
  unsigned long long key[U], prev[U];
  int32_t row[U];
  uint32_t pad[K2 ? U : 1];
  Slot* tab[U];
  unsigned s[U], mask[U];
  bool live[U];
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const size_t i = tile_lo + (size_t)u * kThreads + threadIdx.x;
    live[u] = i < b.n;
    key[u] = 0;
    row[u] = 0;
    if (K2) pad[u] = 0;
    if (live[u]) {
      row[u] = b.rows ? b.rows[i] : (int32_t)i;
      if (!b.rows && !(bit_valid(b.valid, i) && (!K2 || bit_valid(b.valid2, i)))) live[u] = false;
      key[u] = (unsigned long long)b.keys[i];
      if (K2) pad[u] = b.k2[i];
    }
    if (live[u] && key[u] == kEmptyKey) {
      flags[1] = 1;
      live[u] = false;
    }
    uint32_t h = KeyBits<KT>::hash((KT)key[u]);
    if (K2) h = with_k2(h, pad[u]);
    const unsigned p = g.pid(h);
    tab[u] = t.slots + t.offset[p];
    mask[u] = t.mask[p];
    s[u] = slot_hash(h) & mask[u];
  }
  // Rounds: every still-pending insert of the thread issues its CAS, then all results are examined.
  // (Resolving row after row makes a warp pay max-over-32-lanes of the probe length once PER ROW: the
  // first version of this loop spent 34 warps-per-issue on the scoreboard.)
  unsigned pend = 0;
#pragma unroll
  for (int u = 0; u < U; ++u)
    if (live[u]) pend |= 1u << u;
  while (pend) {
#pragma unroll
    for (int u = 0; u < U; ++u)
      if ((pend >> u) & 1u) prev[u] = atomicCAS(&tab[u][s[u]].key, kEmptyKey, key[u]);
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (!((pend >> u) & 1u)) continue;
      if (prev[u] == kEmptyKey) {  // claimed
        tab[u][s[u]].row = row[u];
        if (K2) tab[u][s[u]].pad = (int32_t)pad[u];
        pend &= ~(1u << u);
      } else {
        if (prev[u] == key[u]) flags[0] = 1;  // duplicate build key
        s[u] = (s[u] + 1) & mask[u];         // linear probing
      }
    }
  }
}

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

static __device__ __forceinline__ Slot ld_slot(const Slot* p) {
  const uint4 raw = *reinterpret_cast<const uint4*>(p);  // cached load: tables are meant to live in L2
  Slot s;
  s.key = ((unsigned long long)raw.y << 32) | raw.x;
  s.row = (int32_t)raw.z;
  s.pad = (int32_t)raw.w;
  return s;
}

// LEFT_LIKE: unmatched / NULL probe rows emit (row,-1).  UNIQUE: build keys are unique, stop at the
// first match.  WRITE=false: count only.
template <typename KT, bool LEFT_LIKE, bool UNIQUE, bool WRITE, bool K2>
__global__ void __launch_bounds__(kThreads)
part_probe_kernel(Pairs<KT> pr, PartGeom g, Tables t, int32_t* __restrict__ out_probe,
                  int32_t* __restrict__ out_build, unsigned long long* __restrict__ cursor) {
  __shared__ unsigned warp_sums[kThreads / 32];
  __shared__ unsigned long long tile_out;
  const size_t tile_base = (size_t)blockIdx.x * kProbeTile;
  unsigned long long key[kProbeRows];
  int32_t prow[kProbeRows], first[kProbeRows];
  unsigned cnt[kProbeRows], start[kProbeRows], mask[kProbeRows];
  int32_t k2v[K2 ? kProbeRows : 1];
  const Slot* tab[kProbeRows];
  bool lookup[kProbeRows];
#pragma unroll
  for (int i = 0; i < kProbeRows; ++i) {
    const size_t j = tile_base + (size_t)i * kThreads + threadIdx.x;
    if (K2) k2v[i] = 0;
    cnt[i] = 0;
    first[i] = -1;
    lookup[i] = false;
    prow[i] = -1;
    key[i] = 0;
    tab[i] = t.slots;
    start[i] = 0;
    mask[i] = 0;
    if (j < pr.n) {
      const KT kraw = pr.keys[j];
      key[i] = (unsigned long long)kraw;
      bool ok = true;
      if (pr.rows) {
        const int32_t tag = pr.rows[j];
        ok = tag >= 0;  // negative tag = NULL-key row kept for LEFT/FULL
        prow[i] = ok ? tag : ~tag;
      } else {
        prow[i] = (int32_t)j;
        ok = bit_valid(pr.valid, j) && (!K2 || bit_valid(pr.valid2, j));
      }
      if (K2) k2v[i] = (int32_t)pr.k2[j];
      if (ok && key[i] != kEmptyKey) {
        uint32_t h = KeyBits<KT>::hash(kraw);
        if (K2) h = with_k2(h, (uint32_t)k2v[i]);
        const unsigned p = g.pid(h);
        tab[i] = t.slots + t.offset[p];
        mask[i] = t.mask[p];
        start[i] = slot_hash(h) & mask[i];
        lookup[i] = true;
      }
      if (LEFT_LIKE) cnt[i] = 1;  // at least (row,-1)
    }
  }
  // first slot of every row fetched before any chain is walked (independent L2 requests)
  Slot s0[kProbeRows];
#pragma unroll
  for (int i = 0; i < kProbeRows; ++i)
    if (lookup[i]) s0[i] = ld_slot(tab[i] + start[i]);
#pragma unroll
  for (int i = 0; i < kProbeRows; ++i) {
    if (!lookup[i]) continue;
    unsigned s = start[i], c = 0;
    Slot cur = s0[i];
    while (cur.key != kEmptyKey) {
      if (cur.key == key[i] && (!K2 || cur.pad == k2v[i])) {
        if (c == 0) first[i] = cur.row;
        ++c;
        if (UNIQUE) break;
      }
      s = (s + 1) & mask[i];
      cur = ld_slot(tab[i] + s);
    }
    if (c) cnt[i] = c;
  }
  unsigned excl[kProbeRows], slice_total[kProbeRows], tile_total = 0;
#pragma unroll
  for (int i = 0; i < kProbeRows; ++i) {
    excl[i] = block_exclusive_scan(cnt[i], warp_sums, &slice_total[i]);
    tile_total += slice_total[i];
  }
  if (threadIdx.x == 0) tile_out = tile_total ? atomicAdd(cursor, (unsigned long long)tile_total) : 0ull;
  if (!WRITE) return;
  __syncthreads();
  size_t pos0 = (size_t)tile_out;
#pragma unroll
  for (int i = 0; i < kProbeRows; ++i) {
    size_t pos = pos0 + excl[i];
    if (cnt[i] == 1) {
      out_probe[pos] = prow[i];
      out_build[pos] = first[i];
    } else if (cnt[i] > 1) {
      unsigned s = start[i];
      Slot cur = ld_slot(tab[i] + s);
      while (cur.key != kEmptyKey) {
        if (cur.key == key[i] && (!K2 || cur.pad == k2v[i])) {
          out_probe[pos] = prow[i];
          out_build[pos] = cur.row;
          ++pos;
        }
        s = (s + 1) & mask[i];
        cur = ld_slot(tab[i] + s);
      }
    }
    pos0 += slice_total[i];
  }
}

// ---- probe v2: persistent CTAs, TMA-staged pair tiles, two barriers per 2048-row tile ----
//
// profiles/r01a_ncu_full_summary.md, part_probe_kernel: 13.3 ms of the 25.2 ms C3 step, DRAM at 23 %,
// issue slots 40 % busy (~195 instructions per row: MurmurHash3, four block-wide scans with two
// barriers each per 1024-row tile), 14 warps-per-issue stalled on L2/DRAM latency with nothing
// prefetched across tiles.  v2:
//   * {key,tag} tiles arrive through a 3-stage cp.async.bulk ring (stream.cuh), L2 evict_first;
//     table slots are read with an L2 evict_last hint so that streaming traffic does not push the
//     hot partitions' tables out of L2;
//   * a warp owns 256 consecutive rows of the tile, lane l handles rows i*32+l (i = 0..7): all eight
//     slot lookups of a thread are issued before the first is consumed; ranks inside a warp come from
//     ballots (0/1 matches) or a shuffle scan (duplicate build keys);
//   * one cursor atomic and two barriers per tile; output stores of a warp step are consecutive ints.
constexpr int kP2Threads = 512;
constexpr int kP2Warps = kP2Threads / 32;
constexpr int kP2Rows = 4;
constexpr int kP2Tile = kP2Threads * kP2Rows;
constexpr int kP2Stages = 3;

template <typename KT>
struct Probe2Geom {
  static constexpr int kKeyBytes = kP2Tile * (int)sizeof(KT);
  static constexpr int kTagBytes = kP2Tile * 4;
  static constexpr int kStageBytes = kKeyBytes + kTagBytes;
};
struct Probe2Smem {
  uint64_t bar[kP2Stages];
  unsigned tile[kP2Stages];
  unsigned warp_tot[2][kP2Warps];
  unsigned long long tile_out[2];
  unsigned long long part_off[kMaxParts];  // copies of Tables::offset / mask: the probe loop must not pay
  unsigned part_mask[kMaxParts];           // an extra L2 round trip per round to fetch them
  // per-warp straggler queue (UNIQUE path): rows whose first bucket held neither their key nor an EMPTY slot
  unsigned long long q_key[kP2Warps][32 * kP2Rows];  // key; after resolution: output rank (or ~0 = no output)
  unsigned q_where[kP2Warps][32 * kP2Rows];          // partition << 24 | slot; after resolution: build row
  int32_t q_prow[kP2Warps][32 * kP2Rows];
};
template <typename KT>
constexpr size_t probe2_smem_bytes() {
  return (size_t)kP2Stages * Probe2Geom<KT>::kStageBytes + sizeof(Probe2Smem) + 128;
}

static __device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
static __device__ __forceinline__ uint64_t l2_policy_evict_normal() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(p));
  return p;
}
static __device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
static __device__ __forceinline__ Slot ld_slot_hint(const Slot* p, uint64_t policy) {
  uint4 raw;
  asm volatile("ld.global.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;"
               : "=r"(raw.x), "=r"(raw.y), "=r"(raw.z), "=r"(raw.w)
               : "l"(p), "l"(policy));
  Slot s;
  s.key = ((unsigned long long)raw.y << 32) | raw.x;
  s.row = (int32_t)raw.z;
  s.pad = 0;
  return s;
}
struct Bucket2 {  // two adjacent slots = one 32-byte sector
  unsigned long long k0, k1;
  int32_t r0, r1;
};
static __device__ __forceinline__ Bucket2 ld_bucket_hint(const Slot* p, uint64_t policy) {
  unsigned long long a, b, c, d;
  asm volatile("ld.global.L2::cache_hint.v4.u64 {%0,%1,%2,%3}, [%4], %5;"
               : "=l"(a), "=l"(b), "=l"(c), "=l"(d)
               : "l"(p), "l"(policy));
  Bucket2 r;
  r.k0 = a;
  r.r0 = (int32_t)(unsigned)b;
  r.k1 = c;
  r.r1 = (int32_t)(unsigned)d;
  return r;
}
static __device__ __forceinline__ void st_i32_hint(int32_t* p, int32_t v, uint64_t policy) {
  asm volatile("st.global.L2::cache_hint.b32 [%0], %1, %2;" ::"l"(p), "r"(v), "l"(policy) : "memory");
}
static __device__ __forceinline__ void bulk_load_hint(void* smem_dst, const void* gmem_src, uint32_t bytes,
                                                      uint64_t* bar, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
          tma::smem_u32(smem_dst)),
      "l"(gmem_src), "r"(bytes), "r"(tma::smem_u32(bar)), "l"(policy)
      : "memory");
}

template <typename KT, bool LEFT_LIKE, bool UNIQUE, bool WRITE>
__global__ void __launch_bounds__(kP2Threads, 2)
part_probe_stream_kernel(Pairs<KT> pr /* rows != nullptr */, PartGeom g, Tables t, int32_t* __restrict__ out_probe,
                         int32_t* __restrict__ out_build, unsigned long long* __restrict__ cursor,
                         unsigned* __restrict__ ticket, int hint_mode) {
  using G = Probe2Geom<KT>;
  extern __shared__ __align__(16) unsigned char smem_raw[];  // indexed directly so that ptxas emits LDS/STS
  unsigned char* const ring = smem_raw;
  Probe2Smem& sm = *reinterpret_cast<Probe2Smem*>(smem_raw + (size_t)kP2Stages * G::kStageBytes);
  const unsigned tid = threadIdx.x, warp = tid >> 5, lane = tid & 31u;
  const size_t tiles = (pr.n + kP2Tile - 1) / kP2Tile;
  // hint_mode bit 0: streams evict_first, bit 1: tables evict_last bit 2: L2 prefetch of the next tile's buckets (lab knob B200_PROBE_HINTS, default 7)
  const uint64_t pol_stream = (hint_mode & 1) ? l2_policy_evict_first() : l2_policy_evict_normal();
  const uint64_t pol_table = (hint_mode & 2) ? l2_policy_evict_last() : l2_policy_evict_normal();

  auto issue = [&](int s) {
    const unsigned tl = atomicAdd(ticket, 1u);
    sm.tile[s] = tl;
    if ((size_t)tl < tiles && ((size_t)tl + 1) * kP2Tile <= pr.n) {
      unsigned char* dst = ring + (size_t)s * G::kStageBytes;
      tma::mbar_expect_tx(&sm.bar[s], G::kStageBytes);
      bulk_load_hint(dst, pr.keys + (size_t)tl * kP2Tile, G::kKeyBytes, &sm.bar[s], pol_stream);
      bulk_load_hint(dst + G::kKeyBytes, pr.rows + (size_t)tl * kP2Tile, G::kTagBytes, &sm.bar[s], pol_stream);
    }
  };
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < kP2Stages; ++s) tma::mbar_init(&sm.bar[s], 1);
    tma::fence_barrier_init();
#pragma unroll
    for (int s = 0; s < kP2Stages; ++s) issue(s);
  }
  for (unsigned p = tid; p < g.nparts; p += kP2Threads) {
    sm.part_off[p] = t.offset[p];
    sm.part_mask[p] = t.mask[p];
  }
  __syncthreads();

  for (unsigned iter = 0;; ++iter) {
    const int s = (int)(iter % kP2Stages);
    const unsigned tl = sm.tile[s];
    if ((size_t)tl >= tiles) break;
    const size_t tile_row0 = (size_t)tl * kP2Tile;
    const bool full = tile_row0 + kP2Tile <= pr.n;
    unsigned long long key[kP2Rows];
    int32_t prow[kP2Rows], first[kP2Rows];
    unsigned cnt[kP2Rows], where[kP2Rows];  // where = partition id << 24 | current slot (slots <= 2^24 per table)
    unsigned home[UNIQUE ? 1 : kP2Rows];    // first slot of the row's probe sequence (multi-match write pass)
    bool lookup[kP2Rows];
    if (full) tma::mbar_wait(&sm.bar[s], (iter / kP2Stages) & 1u);
    if (hint_mode & 4) {
      // Software pipelining of the table misses: the NEXT tile's keys are already in the ring, so its
      // first buckets are prefetched into L2 now (fire and forget) and are L2 hits one tile later.
      // ~13 % of the look-ups are compulsory DRAM misses (every table sector is read once per ~8
      // probes); without this a round of 2048 look-ups always waits for its slowest DRAM miss.
      const int s1 = (int)((iter + 1) % kP2Stages);
      const unsigned tn = sm.tile[s1];
      if ((size_t)tn < tiles && ((size_t)tn + 1) * kP2Tile <= pr.n) {
        tma::mbar_wait(&sm.bar[s1], ((iter + 1) / kP2Stages) & 1u);
        const KT* nkeys = reinterpret_cast<const KT*>(ring + (size_t)s1 * G::kStageBytes);
#pragma unroll
        for (int i = 0; i < kP2Rows; ++i) {
          const uint32_t h = KeyBits<KT>::hash(nkeys[warp * (32 * kP2Rows) + i * 32 + lane]);
          const unsigned p = g.pid(h);
          const Slot* a = t.slots + sm.part_off[p] + (slot_hash(h) & sm.part_mask[p]);
          asm volatile("prefetch.global.L2::evict_last [%0];" ::"l"(a));
        }
      }
    }
    const KT* skeys = reinterpret_cast<const KT*>(ring + (size_t)s * G::kStageBytes);
    const int32_t* stags = reinterpret_cast<const int32_t*>(ring + (size_t)s * G::kStageBytes + G::kKeyBytes);
#pragma unroll
    for (int i = 0; i < kP2Rows; ++i) {
      const unsigned local = warp * (32 * kP2Rows) + i * 32 + lane;
      const size_t j = tile_row0 + local;
      bool have = full || j < pr.n;
      KT kraw = 0;
      int32_t tag = 0;
      if (full) {
        kraw = skeys[local];
        tag = stags[local];
      } else if (have) {
        kraw = pr.keys[j];
        tag = pr.rows[j];
      }
      key[i] = (unsigned long long)kraw;
      const bool ok = have && tag >= 0;  // negative tag = NULL-key row kept for LEFT/FULL
      prow[i] = tag >= 0 ? tag : ~tag;
      first[i] = -1;
      cnt[i] = (LEFT_LIKE && have) ? 1u : 0u;
      lookup[i] = ok && key[i] != kEmptyKey;
      const uint32_t h = KeyBits<KT>::hash(kraw);
      const unsigned p = g.pid(h);
      where[i] = (p << 24) | (slot_hash(h) & sm.part_mask[p]);
      if (!UNIQUE) home[i] = where[i];
    }
    // Resolve in ROUNDS: all pending look-ups of the thread are in flight together, then examined;
    // rows whose slot held another key advance one slot and go again.  Round 0 settles ~80 % of the
    // rows, and the number of rounds is the longest probe sequence among the warp's 256 rows, paid once
    // per tile (resolving row by row paid max-over-32-lanes per ROW: 26 dependent L2 round trips per
    // tile in the first version of this kernel, profiles/r01b).
    Bucket2 cur[kP2Rows];
    unsigned pend = 0;
#pragma unroll
    for (int i = 0; i < kP2Rows; ++i) {
      if (lookup[i]) {
        pend |= 1u << i;
        cur[i] = ld_bucket_hint(t.slots + sm.part_off[where[i] >> 24] + (where[i] & 0xffffffu), pol_table);
      }
    }
    unsigned qn = 0, q_emitted = 0;  // warp-uniform: queued stragglers / how many of them produce a pair
    if (UNIQUE) {
      // One look at the first bucket settles ~93 % of the rows.  The rest are compacted into the
      // warp's queue and resolved one per LANE (a lane walks its own straggler's probe sequence), so
      // the long tail of linear probing is paid once per tile by a few lanes instead of once per row
      // by the whole warp.
#pragma unroll
      for (int i = 0; i < kP2Rows; ++i) {
        bool pending = false;
        if ((pend >> i) & 1u) {
          if (cur[i].k0 == key[i]) { first[i] = cur[i].r0; cnt[i] = 1; }
          else if (cur[i].k0 == kEmptyKey) {}
          else if (cur[i].k1 == key[i]) { first[i] = cur[i].r1; cnt[i] = 1; }
          else if (cur[i].k1 == kEmptyKey) {}
          else pending = true;
        }
        const unsigned bq = __ballot_sync(0xffffffffu, pending);
        if (bq) {
          if (pending) {
            const unsigned e = qn + __popc(bq & lanemask_lt());
            sm.q_key[warp][e] = key[i];
            sm.q_where[warp][e] = where[i];
            sm.q_prow[warp][e] = prow[i];
            cnt[i] = 0;  // the queue entry produces this row's output
          }
          qn += __popc(bq);
        }
      }
      pend = 0;
      __syncwarp();
      for (unsigned base = 0; base < qn; base += 32) {
        const unsigned e = base + lane;
        bool emit_it = false;
        if (e < qn) {
          const unsigned long long k = sm.q_key[warp][e];
          const unsigned w = sm.q_where[warp][e];
          const Slot* tb = t.slots + sm.part_off[w >> 24];
          const unsigned m = sm.part_mask[w >> 24];
          unsigned at = w & 0xffffffu;
          int32_t found = -1;
          while (true) {
            at = (at + 2u) & m;
            const Bucket2 c = ld_bucket_hint(tb + at, pol_table);
            if (c.k0 == k) { found = c.r0; break; }
            if (c.k0 == kEmptyKey) break;
            if (c.k1 == k) { found = c.r1; break; }
            if (c.k1 == kEmptyKey) break;
          }
          emit_it = LEFT_LIKE || found >= 0;
          sm.q_where[warp][e] = (unsigned)found;
        }
        const unsigned be = __ballot_sync(0xffffffffu, emit_it);
        if (e < qn) sm.q_key[warp][e] = emit_it ? (unsigned long long)(q_emitted + __popc(be & lanemask_lt())) : ~0ull;
        q_emitted += __popc(be);
      }
    }
    while (true) {
      // rows still pending in ANY lane (warp-uniform): later rounds touch only those, so a round costs
      // what its stragglers cost instead of a full pass over the 8 rows
      const unsigned any_check = __reduce_or_sync(0xffffffffu, pend);
      if (any_check == 0) break;
#pragma unroll
      for (int i = 0; i < kP2Rows; ++i) {
        if (!((any_check >> i) & 1u)) continue;
        if (!((pend >> i) & 1u)) continue;
        bool stop = false;
        if (cur[i].k0 == key[i]) {
          if (UNIQUE || first[i] < 0) first[i] = cur[i].r0;
          if (UNIQUE) { cnt[i] = 1; stop = true; }
          else cnt[i] = (cnt[i] & 0x80000000u) ? cnt[i] + 1 : 0x80000001u;  // bit 31: "matched at least once"
        } else if (cur[i].k0 == kEmptyKey) {
          stop = true;
        }
        if (!stop) {
          if (cur[i].k1 == key[i]) {
            if (UNIQUE || first[i] < 0) first[i] = cur[i].r1;
            if (UNIQUE) { cnt[i] = 1; stop = true; }
            else cnt[i] = (cnt[i] & 0x80000000u) ? cnt[i] + 1 : 0x80000001u;
          } else if (cur[i].k1 == kEmptyKey) {
            stop = true;
          }
        }
        if (stop) pend &= ~(1u << i);
      }
      const unsigned any_issue = __reduce_or_sync(0xffffffffu, pend);
#pragma unroll
      for (int i = 0; i < kP2Rows; ++i) {
        if (!((any_issue >> i) & 1u)) continue;
        if (!((pend >> i) & 1u)) continue;
        const unsigned p = where[i] >> 24;
        const unsigned at = ((where[i] & 0xffffffu) + 2u) & sm.part_mask[p];
        where[i] = (p << 24) | at;
        cur[i] = ld_bucket_hint(t.slots + sm.part_off[p] + at, pol_table);
      }
    }
    if (!UNIQUE) {
#pragma unroll
      for (int i = 0; i < kP2Rows; ++i)
        if (cnt[i] & 0x80000000u) cnt[i] &= 0x7fffffffu;  // number of matches (replaces LEFT's provisional 1)
    }
    // ranks: step-major inside the warp, then warps, then the tile's reservation
    unsigned rank[kP2Rows], warp_total = 0;
#pragma unroll
    for (int i = 0; i < kP2Rows; ++i) {
      if (UNIQUE) {  // counts are 0/1
        const unsigned b = __ballot_sync(0xffffffffu, cnt[i] != 0);
        rank[i] = warp_total + __popc(b & lanemask_lt());
        warp_total += __popc(b);
      } else {
        unsigned inc = cnt[i];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          const unsigned o = __shfl_up_sync(0xffffffffu, inc, d);
          if (lane >= (unsigned)d) inc += o;
        }
        rank[i] = warp_total + inc - cnt[i];
        warp_total += __shfl_sync(0xffffffffu, inc, 31);
      }
    }
    const unsigned regular_total = warp_total;
    warp_total += q_emitted;
    const unsigned buf = iter & 1u;
    if (lane == 0) sm.warp_tot[buf][warp] = warp_total;
    __syncthreads();  // stage s is consumed; warp totals visible
    if (tid == 32) issue(s);
    if (tid == 0) {
      unsigned tot = 0;
#pragma unroll
      for (int w = 0; w < kP2Warps; ++w) tot += sm.warp_tot[buf][w];
      sm.tile_out[buf] = tot ? atomicAdd(cursor, (unsigned long long)tot) : 0ull;
    }
    if (!WRITE) continue;  // count-only pass: the cursor is the result (tile_out is double-buffered)
    __syncthreads();
    size_t pos0 = (size_t)sm.tile_out[buf];
    for (unsigned w = 0; w < warp; ++w) pos0 += sm.warp_tot[buf][w];
#pragma unroll
    for (int i = 0; i < kP2Rows; ++i) {
      size_t pos = pos0 + rank[i];
      if (cnt[i] == 1) {
        st_i32_hint(out_probe + pos, prow[i], pol_stream);
        st_i32_hint(out_build + pos, first[i], pol_stream);
      } else if (cnt[i] > 1) {
        const unsigned hw = home[UNIQUE ? 0 : i];
        const Slot* tb = t.slots + sm.part_off[hw >> 24];
        const unsigned m = sm.part_mask[hw >> 24];
        unsigned at = hw & 0xffffffu;
        Slot c2 = ld_slot_hint(tb + at, pol_table);
        while (c2.key != kEmptyKey) {
          if (c2.key == key[i]) {
            out_probe[pos] = prow[i];
            out_build[pos] = c2.row;
            ++pos;
          }
          at = (at + 1) & m;
          c2 = ld_slot_hint(tb + at, pol_table);
        }
      }
    }
    if (UNIQUE) {  // the warp's stragglers follow its regular rows
      for (unsigned e = lane; e < qn; e += 32) {
        const unsigned long long r = sm.q_key[warp][e];
        if (r == ~0ull) continue;
        const size_t pos = pos0 + regular_total + (size_t)r;
        st_i32_hint(out_probe + pos, sm.q_prow[warp][e], pol_stream);
        st_i32_hint(out_build + pos, (int32_t)sm.q_where[warp][e], pol_stream);
      }
      __syncwarp();  // the queue is rewritten by the next tile
    }
  }
}

__global__ void mark_rows_kernel(const int32_t* __restrict__ idx, size_t n, unsigned char* __restrict__ marks,
                                 size_t limit) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int32_t v = idx[i];
    if (v >= 0 && (size_t)v < limit) marks[v] = 1;
  }
}

// append (-1, r) for every unmarked build row; positions claimed with one atomic per warp
__global__ void append_unmatched_kernel(const unsigned char* __restrict__ marks, size_t build_rows,
                                        int32_t* __restrict__ out_probe, int32_t* __restrict__ out_build,
                                        unsigned long long* __restrict__ cursor) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const size_t n_round = (build_rows + stride - 1) / stride * stride;
  for (size_t r = (size_t)blockIdx.x * blockDim.x + threadIdx.x; r < n_round; r += stride) {
    const bool have = r < build_rows && marks[r] == 0;
    const unsigned m = __ballot_sync(0xffffffffu, have);
    if (m == 0) continue;
    unsigned long long base = 0;
    const int leader = __ffs(m) - 1;
    if ((int)lane_id() == leader) base = atomicAdd(cursor, (unsigned long long)__popc(m));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (have) {
      const size_t at = (size_t)(base + __popc(m & lanemask_lt()));
      out_probe[at] = -1;
      out_build[at] = (int32_t)r;
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

template <typename KT, bool KEEP_NULLS, bool COMPACT = false>
gdf_error partition_hist(const gdf_column* col, PartGeom g, unsigned long long* d_totals /*device [nparts + 1]*/,
                         unsigned long long* h_totals, const gdf_column* col2 = nullptr, unsigned* h_hi_or = nullptr) {
  const KT* keys = static_cast<const KT*>(col->data);
  const size_t n = col->size;
  // d_totals[nparts] doubles as the "OR of the high key words" cell
  B200_CUDA_TRY(cudaMemsetAsync(d_totals, 0, (g.nparts + 1) * sizeof(unsigned long long), 0));
  const int blocks = sm_count() * 4;
  {
    B200_TIMED("join_part_hist");
    const bool vec = col->valid == nullptr && col2 == nullptr && aligned16(keys);
    auto kern = vec ? part_hist_kernel<KT, KEEP_NULLS, COMPACT, true> : part_hist_kernel<KT, KEEP_NULLS, COMPACT, false>;
    kern<<<blocks, kThreads>>>(keys, col->valid, n, g, d_totals, col2 ? static_cast<const uint32_t*>(col2->data) : nullptr,
                               col2 ? col2->valid : nullptr, reinterpret_cast<unsigned*>(d_totals + g.nparts));
  }
  B200_CHECK_LAST();
  if (h_totals == nullptr) return GDF_SUCCESS;  // device-only caller (asynchronous exchange): no read-back, no host sync
  unsigned long long h_all[kMaxParts + 1];
  B200_CUDA_TRY(cudaMemcpy(h_all, d_totals, (g.nparts + 1) * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  for (unsigned p = 0; p < g.nparts; ++p) h_totals[p] = h_all[p];
  if (h_hi_or) *h_hi_or = (unsigned)h_all[g.nparts];
  return GDF_SUCCESS;
}

template <typename KT, bool KEEP_NULLS>
gdf_error partition_scatter32(const gdf_column* col, PartGeom g, const unsigned long long* h_cursors,
                              unsigned long long* d_cursors, uint2* out_pairs, const int32_t* payload, int32_t id_base,
                              const PeerPairs* peer_dst = nullptr, bool cursors_on_device = false, int ctas_per_sm = 0) {
  PeerPairs peer;
  peer.on = 0;
  peer.shift = 0;
  peer.status = nullptr;
  peer.region_cap = 0;
  peer.overflow = nullptr;
  peer.counts = nullptr;
  peer.offsets0 = nullptr;
  for (int r = 0; r < kMaxPeers; ++r) peer.base[r] = nullptr;
  if (peer_dst) peer = *peer_dst;
  const KT* keys = static_cast<const KT*>(col->data);
  const size_t n = col->size;
  if (cursors_on_device)  // asynchronous exchange: the plan lives on the device, nothing here waits for the GPU
    B200_CUDA_TRY(cudaMemcpyAsync(d_cursors, h_cursors, g.nparts * sizeof(unsigned long long), cudaMemcpyDeviceToDevice, 0));
  else
    B200_CUDA_TRY(cudaMemcpy(d_cursors, h_cursors, g.nparts * sizeof(unsigned long long), cudaMemcpyHostToDevice));
  auto kern = part_scatter32_kernel<KT, KEEP_NULLS>;
  const size_t smem_bytes = sizeof(Scatter32Smem);
  B200_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
  int per_sm = 1;
  B200_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kS32Threads, smem_bytes));
  const size_t tiles = (n + kS32Tile - 1) / kS32Tile;
  // exactly one resident wave; a caller that runs another kernel beside this one (the multi-GPU layer fills its hash
  // tables while the probe side crosses NVLink) asks for fewer CTAs per SM, which leaves registers for the neighbour
  if (ctas_per_sm > 0 && ctas_per_sm < per_sm) per_sm = ctas_per_sm;
  const size_t cap = (size_t)sm_count() * (size_t)(per_sm > 0 ? per_sm : 1);
  const int sblocks = (int)(tiles < cap ? (tiles ? tiles : 1) : cap);
  // single-GPU path: the same bulk-store kernel with one destination (measured at C3: 4.83 -> 4.38 ms for both sides;
  // the element-parallel copy-out kept the LSU at 74 %, profiles/r02a_ncu_join_full.md)
  if (!peer.on && out_pairs != nullptr && lab_knob("B200_SCATTER_BULK", 1) != 0) {
    peer.on = 1;
    peer.shift = 8;  // bin >> 8 == 0: everything goes to base[0]
    peer.base[0] = out_pairs;
  }
  if (peer.on && peer.counts != nullptr && !KEEP_NULLS && col->valid == nullptr && payload == nullptr) {
    // the multi-GPU exchange: even runs only (no 8-byte peer stores in the steady state)
    auto ekern = part_scatter32_even_kernel<KT>;
    const size_t esmem = sizeof(Scatter32EvenSmem);
    B200_CUDA_TRY(cudaFuncSetAttribute(ekern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)esmem));
    int eper_sm = 1;
    B200_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&eper_sm, ekern, kS32Threads, esmem));
    if (ctas_per_sm > 0 && ctas_per_sm < eper_sm) eper_sm = ctas_per_sm;
    const size_t ecap = (size_t)sm_count() * (size_t)(eper_sm > 0 ? eper_sm : 1);
    const int eblocks = (int)(tiles < ecap ? (tiles ? tiles : 1) : ecap);
    unsigned* d_tail = reinterpret_cast<unsigned*>(d_cursors + g.nparts);   // the caller's scratch holds 2 x nparts cells
    B200_CUDA_TRY(cudaMemsetAsync(d_tail, 0, g.nparts * sizeof(unsigned), 0));
    B200_TIMED("join_part_scatter");
    ekern<<<eblocks, kS32Threads, esmem>>>(keys, n, g, d_cursors, id_base, peer, peer.counts, peer.offsets0, d_tail);
  } else if (peer.on && !KEEP_NULLS && col->valid == nullptr && payload == nullptr) {  // single destination: TMA bulk stores per bin
    auto bkern = part_scatter32_bulk_kernel<KT>;
    const size_t bsmem = sizeof(Scatter32BulkSmem);
    B200_CUDA_TRY(cudaFuncSetAttribute(bkern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bsmem));
    int bper_sm = 1;
    B200_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bper_sm, bkern, kS32Threads, bsmem));
    if (ctas_per_sm > 0 && ctas_per_sm < bper_sm) bper_sm = ctas_per_sm;
    const size_t bcap = (size_t)sm_count() * (size_t)(bper_sm > 0 ? bper_sm : 1);
    const int bblocks = (int)(tiles < bcap ? (tiles ? tiles : 1) : bcap);
    B200_TIMED("join_part_scatter");
    bkern<<<bblocks, kS32Threads, bsmem>>>(keys, n, g, d_cursors, id_base, peer);
  } else {
    B200_TIMED("join_part_scatter");
    kern<<<sblocks, kS32Threads, smem_bytes>>>(keys, col->valid, n, g, d_cursors, out_pairs, payload, id_base, peer);
  }
  B200_CHECK_LAST();
  return GDF_SUCCESS;
}

// h_cursors[p] = index of partition p's first row in its destination array (one shared array: the exclusive
// scan of the counts; peer mode: this rank's offset inside destination p's receive buffer)
template <typename KT, bool KEEP_NULLS>
gdf_error partition_scatter(const gdf_column* col, PartGeom g, const unsigned long long* h_cursors,
                            unsigned long long* d_cursors, KT* out_keys, int32_t* out_rows, const int32_t* payload,
                            int32_t id_base, const PeerDst& peer, const gdf_column* col2 = nullptr,
                            uint32_t* out_k2 = nullptr) {
  const KT* keys = static_cast<const KT*>(col->data);
  const size_t n = col->size;
  B200_CUDA_TRY(cudaMemcpy(d_cursors, h_cursors, g.nparts * sizeof(unsigned long long), cudaMemcpyHostToDevice));
  const bool prefetch = lab_knob("B200_SCATTER_PREFETCH", 0) != 0;
  auto kern = col2 ? part_scatter_kernel<KT, KEEP_NULLS, false, true>
                   : (prefetch ? part_scatter_kernel<KT, KEEP_NULLS, true, false> : part_scatter_kernel<KT, KEEP_NULLS, false, false>);
  const size_t smem_bytes = sizeof(ScatterSmem<KT>) + (col2 ? kScatterTile * sizeof(uint32_t) : 0);
  ScatterK2 k2{col2 ? static_cast<const uint32_t*>(col2->data) : nullptr, col2 ? col2->valid : nullptr, out_k2};
  B200_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
  int per_sm = 1;
  B200_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kThreads, smem_bytes));
  const size_t tiles = (n + kScatterTile - 1) / kScatterTile;
  const size_t cap = (size_t)sm_count() * (size_t)(per_sm > 0 ? per_sm : 1);  // exactly one resident wave
  const int sblocks = (int)(tiles < cap ? (tiles ? tiles : 1) : cap);
  {
    B200_TIMED("join_part_scatter");
    kern<<<sblocks, kThreads, smem_bytes>>>(keys, col->valid, n, g, d_cursors, out_keys, out_rows, payload, id_base, peer,
                                            k2);
  }
  B200_CHECK_LAST();
  return GDF_SUCCESS;
}

template <typename KT, bool KEEP_NULLS>
gdf_error partition_side(const gdf_column* col, PartGeom g, Scratch& keys_out, Scratch& rows_out,
                         unsigned long long* d_totals /*device [nparts]*/, unsigned long long* d_cursors,
                         unsigned long long* h_totals, size_t* kept, const int32_t* payload = nullptr,
                         int32_t id_base = 0, KT* ext_keys = nullptr, int32_t* ext_rows = nullptr,
                         const gdf_column* col2 = nullptr, Scratch* k2_out = nullptr) {
  gdf_error e = partition_hist<KT, KEEP_NULLS>(col, g, d_totals, h_totals, col2);
  if (e != GDF_SUCCESS) return e;
  unsigned long long h_cursors[kMaxParts], run = 0;
  for (unsigned p = 0; p < g.nparts; ++p) {
    h_cursors[p] = run;
    run += h_totals[p];
  }
  *kept = (size_t)run;
  if (!ext_keys) {
    B200_CUDA_TRY(keys_out.alloc((run ? run : 1) * sizeof(KT)));
    B200_CUDA_TRY(rows_out.alloc((run ? run : 1) * sizeof(int32_t)));
    ext_keys = keys_out.as<KT>();
    ext_rows = rows_out.as<int32_t>();
  }
  PeerDst none;
  none.on = 0;
  uint32_t* k2_ptr = nullptr;
  if (col2) {
    B200_CUDA_TRY(k2_out->alloc((run ? run : 1) * sizeof(uint32_t)));
    k2_ptr = k2_out->as<uint32_t>();
  }
  return partition_scatter<KT, KEEP_NULLS>(col, g, h_cursors, d_cursors, ext_keys, ext_rows, payload, id_base, none, col2,
                                           k2_ptr);
}

gdf_error read_u64(const unsigned long long* d, unsigned long long* h) {
  unsigned long long* box = static_cast<unsigned long long*>(pinned_mailbox());
  B200_REQUIRE(box != nullptr, GDF_CUDA_ERROR);
  B200_CUDA_TRY(cudaMemcpyAsync(box, d, sizeof(unsigned long long), cudaMemcpyDeviceToHost, 0));
  B200_CUDA_TRY(cudaStreamSynchronize(0));
  *h = *box;
  return GDF_SUCCESS;
}

template <typename KT, bool LEFT_LIKE, bool UNIQUE, bool WRITE>
gdf_error launch_probe_stream(const Pairs<KT>& pr, PartGeom g, const Tables& t, int32_t* op, int32_t* ob,
                              unsigned long long* cursor, unsigned* ticket) {
  auto kern = part_probe_stream_kernel<KT, LEFT_LIKE, UNIQUE, WRITE>;
  const int smem = (int)probe2_smem_bytes<KT>();
  B200_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  B200_CUDA_TRY(cudaMemsetAsync(ticket, 0, sizeof(unsigned), 0));
  const size_t tiles = (pr.n + kP2Tile - 1) / kP2Tile;
  const size_t resident = (size_t)sm_count() * 2;
  const unsigned blocks = (unsigned)(tiles < resident ? tiles : resident);
  const int hint_mode = lab_knob("B200_PROBE_HINTS", 7);
  kern<<<blocks, kP2Threads, smem>>>(pr, g, t, op, ob, cursor, ticket, hint_mode);
  B200_CHECK_LAST();
  return GDF_SUCCESS;
}

template <typename KT, bool LEFT_LIKE>
gdf_error launch_probe(bool unique, bool write, const Pairs<KT>& pr, PartGeom g, const Tables& t, int32_t* op,
                       int32_t* ob, unsigned long long* cursor, unsigned* ticket, bool stream_ok) {
  if (pr.n == 0) return GDF_SUCCESS;
  B200_TIMED(write ? "join_part_probe" : "join_part_count");
  const bool force_v1 = lab_knob("B200_PROBE_V1", 1) != 0;  // the one-tile-per-CTA kernel is the fastest measured for 16-byte slots (profiles/r01b)
  if (pr.rows != nullptr && stream_ok && !force_v1) {  // partitioned {key,tag} pairs: streaming kernel
    if (unique) {
      if (write) return launch_probe_stream<KT, LEFT_LIKE, true, true>(pr, g, t, op, ob, cursor, ticket);
      return launch_probe_stream<KT, LEFT_LIKE, true, false>(pr, g, t, op, ob, cursor, ticket);
    }
    if (write) return launch_probe_stream<KT, LEFT_LIKE, false, true>(pr, g, t, op, ob, cursor, ticket);
    return launch_probe_stream<KT, LEFT_LIKE, false, false>(pr, g, t, op, ob, cursor, ticket);
  }
  const unsigned tiles = (unsigned)((pr.n + kProbeTile - 1) / kProbeTile);
  void (*kern)(Pairs<KT>, PartGeom, Tables, int32_t*, int32_t*, unsigned long long*) = nullptr;
  if (pr.k2) {
    kern = unique ? (write ? part_probe_kernel<KT, LEFT_LIKE, true, true, true> : part_probe_kernel<KT, LEFT_LIKE, true, false, true>)
                  : (write ? part_probe_kernel<KT, LEFT_LIKE, false, true, true> : part_probe_kernel<KT, LEFT_LIKE, false, false, true>);
  } else {
    kern = unique ? (write ? part_probe_kernel<KT, LEFT_LIKE, true, true, false> : part_probe_kernel<KT, LEFT_LIKE, true, false, false>)
                  : (write ? part_probe_kernel<KT, LEFT_LIKE, false, true, false> : part_probe_kernel<KT, LEFT_LIKE, false, false, false>);
  }
  kern<<<tiles, kThreads>>>(pr, g, t, op, ob, cursor);
  B200_CHECK_LAST();
  return GDF_SUCCESS;
}

void view_indices(gdf_column* c, int32_t* data, size_t n) {
  if (n == 0 && data == nullptr) gdf_column_view(c, nullptr, nullptr, 0, N_GDF_TYPES);
  else gdf_column_view(c, data, nullptr, n, GDF_INT32);
}

// ---- compact path (join_compact.cuh): tables, build, probe, fix-up for {key32, tag32} pairs ----
template <bool LEFT_LIKE, bool UNIQUE, int MODE>
gdf_error launch_probe32(const Pairs32& pr, PartGeom g, const Tables32& t, const Probe32Out& out, unsigned blocks) {
  auto kern = probe32_kernel<LEFT_LIKE, UNIQUE, MODE>;
  const int smem = (int)probe32_smem_bytes();
  B200_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  kern<<<blocks, kC32Threads, smem>>>(pr, g, t, out, (unsigned)lab_knob("B200_LAB_PROBE", 0));
  B200_CHECK_LAST();
  return GDF_SUCCESS;
}

template <bool LEFT_LIKE>
gdf_error launch_probe32_unique(const Pairs32& pr, PartGeom g, const Tables32& t, const Probe32Out& out,
                                const unsigned long long* d_pstart, unsigned blocks, unsigned cap_tiles = 0) {
  auto kern = cap_tiles ? probe32_unique_kernel<LEFT_LIKE, true> : probe32_unique_kernel<LEFT_LIKE, false>;
  const int smem = (int)probe32u_smem_bytes();
  B200_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  kern<<<blocks, kC32Threads, smem>>>(pr, g, t, out, d_pstart, cap_tiles, (unsigned)lab_knob("B200_LAB_PROBE", 0));
  B200_CHECK_LAST();
  return GDF_SUCCESS;
}

// Table geometry of the compact path, passed to the device BY VALUE (3 KB of kernel parameters): no host-to-device copy,
// so nothing synchronises and the stage can run on any stream.
struct TableMeta {
  unsigned long long off[kMaxParts];
  unsigned mask[kMaxParts];
  unsigned n;
};
__global__ void table_meta_kernel(const TableMeta m, unsigned long long* __restrict__ d_toffset, unsigned* __restrict__ d_tmask) {
  const unsigned p = threadIdx.x;
  if (p < m.n) {
    d_toffset[p] = m.off[p];
    d_tmask[p] = m.mask[p];
  }
}

// The library's private non-blocking stream: table builds that run beside a scatter on the legacy stream.
static cudaStream_t xjoin_side_stream() {
  static cudaStream_t s = nullptr;
  if (s == nullptr && cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking) != cudaSuccess) s = nullptr;
  return s;
}

// Stage 1 of the compact join: size, initialise and fill the per-partition tables, all on stream `s` (the legacy stream
// for a single-GPU join; a private stream when the multi-GPU layer overlaps it with the probe side's exchange).
gdf_error compact_build(PartGeom g, const Pairs32& bp, const unsigned long long* h_btot, unsigned long long* d_toffset,
                        unsigned* d_tmask, int* d_flags, Scratch& table, Tables32* t_out, cudaStream_t s) {
  TableMeta meta;
  unsigned long long total_slots = 0;
  for (unsigned p = 0; p < g.nparts; ++p) {  // load factor <= 0.5, whole buckets
    unsigned slots = pow2_at_least(h_btot[p] ? 2 * h_btot[p] : 4);
    if (slots < 4) slots = 4;
    meta.off[p] = total_slots;
    meta.mask[p] = slots - 1;
    total_slots += slots;
  }
  meta.n = g.nparts;
  table_meta_kernel<<<1, kMaxParts, 0, s>>>(meta, d_toffset, d_tmask);
  B200_CHECK_LAST();
  B200_CUDA_TRY(table.alloc(total_slots * sizeof(unsigned long long)));
  const Tables32 t{table.as<unsigned long long>(), d_toffset, d_tmask};
  *t_out = t;
  // ONE memset + ONE build launch: filling the tables a few partitions at a time (so that the EMPTY pattern is still in
  // L2 when the inserts arrive) measured slower - 32 rounds of 48 MB: 3.70 ms against 3.17 ms, the gaps and tails of 64
  // small launches outweigh the L2 hits (profiles/r02_notes.md; lab knob B200_BUILD_ROUND_MB).
  const unsigned long long kRoundSlots = ((unsigned long long)lab_knob("B200_BUILD_ROUND_MB", 1 << 20) << 20) / sizeof(unsigned long long);
  unsigned long long pair_lo = 0;
  for (unsigned p = 0; p < g.nparts;) {
    unsigned q = p;
    unsigned long long slots = 0, pairs = 0;
    do {
      slots += (unsigned long long)meta.mask[q] + 1;
      pairs += h_btot[q];
      ++q;
    } while (q < g.nparts && slots + meta.mask[q] + 1 <= kRoundSlots);
    B200_CUDA_TRY(cudaMemsetAsync(t.slots + meta.off[p], 0xff, slots * sizeof(unsigned long long), s));
    if (pairs) {
      const Pairs32 part{bp.pairs + pair_lo, (size_t)pairs};
      build32_kernel<<<(unsigned)((pairs + kB32Tile - 1) / kB32Tile), kB32Threads, 0, s>>>(part, g, t, d_flags);
      B200_CHECK_LAST();
    }
    pair_lo += pairs;
    p = q;
  }
  return GDF_SUCCESS;
}

gdf_error compact_probe(int kind, bool flip, PartGeom g, const Tables32& t, const Pairs32& pp, size_t build_rows,
                        unsigned long long* d_cursor, int* d_flags, const unsigned long long* d_pstart, gdf_column* out_l,
                        gdf_column* out_r, unsigned cap_tiles = 0, size_t probe_rows = 0);

// Stage 2: probe + output (legacy stream).  `t` was filled by compact_build; the caller has ordered this stream after it.
// cap_tiles != 0: padded probe layout (join_compact.cuh) - pp.n is the padded length, probe_rows the real row count.
gdf_error compact_probe(int kind, bool flip, PartGeom g, const Tables32& t, const Pairs32& pp, size_t build_rows,
                        unsigned long long* d_cursor, int* d_flags, const unsigned long long* d_pstart, gdf_column* out_l,
                        gdf_column* out_r, unsigned cap_tiles, size_t probe_rows) {
  const bool left_like = kind != JOIN_INNER;
  const size_t real_rows = cap_tiles ? probe_rows : pp.n;
  int h_flags[2] = {0, 0};
  B200_CUDA_TRY(cudaMemcpy(h_flags, d_flags, sizeof(h_flags), cudaMemcpyDeviceToHost));
  const bool unique = h_flags[0] == 0;
  if (pp.n == 0 && kind != JOIN_FULL) {
    view_indices(out_l, nullptr, 0);
    view_indices(out_r, nullptr, 0);
    return GDF_SUCCESS;
  }
  const size_t tiles = (pp.n + kC32Tile - 1) / kC32Tile;
  size_t want = (tiles + kC32Warps - 1) / kC32Warps;
  const size_t resident = (size_t)sm_count();  // one 512-thread CTA (128 KB of rings) per SM
  unsigned blocks = (unsigned)(want < resident ? (want ? want : 1) : resident);
  if (2u * blocks * kC32Warps > (unsigned)kFixCap) blocks = kFixCap / (2 * kC32Warps);
  const unsigned warps = blocks * kC32Warps;

  Probe32Out out{nullptr, nullptr, d_cursor, nullptr, nullptr};
  gdf_error e = GDF_SUCCESS;
  size_t capacity = real_rows;
  if (cap_tiles && (!unique || left_like)) return GDF_INVALID_API_CALL;  // the padded layout is for INNER + unique keys only
  if (!unique) {  // exact count instead of the reference's estimate / retry loop
    B200_TIMED("join_part_count");
    e = left_like ? launch_probe32<true, false, P32_COUNT>(pp, g, t, out, blocks)
                  : launch_probe32<false, false, P32_COUNT>(pp, g, t, out, blocks);
    if (e != GDF_SUCCESS) return e;
    unsigned long long exact = 0;
    if ((e = read_u64(d_cursor, &exact)) != GDF_SUCCESS) return e;
    capacity = (size_t)exact;
    B200_CUDA_TRY(cudaMemsetAsync(d_cursor, 0, sizeof(unsigned long long), 0));
  } else if (!left_like) {
    capacity = real_rows + (size_t)2 * warps * kC32Chunk;  // every warp may leave two partly filled chunks behind
  }
  if (kind == JOIN_FULL) capacity += build_rows;
  if (capacity == 0) {
    view_indices(out_l, nullptr, 0);
    view_indices(out_r, nullptr, 0);
    return GDF_SUCCESS;
  }
  int32_t *op = nullptr, *ob = nullptr;
  B200_RMM_TRY(output_alloc((void**)&op, capacity * sizeof(int32_t)));
  if (output_alloc((void**)&ob, capacity * sizeof(int32_t)) != RMM_SUCCESS) {
    rmmFree(op, 0);
    return GDF_MEMORYMANAGER_ERROR;
  }
  out.probe = op;
  out.build = ob;
  unsigned long long found = 0;
  Scratch holes, plan;
  if (pp.n) {
    if (!unique) {
      B200_TIMED("join_part_probe");
      e = left_like ? launch_probe32<true, false, P32_CURSOR>(pp, g, t, out, blocks)
                    : launch_probe32<false, false, P32_CURSOR>(pp, g, t, out, blocks);
      if (e == GDF_SUCCESS) e = read_u64(d_cursor, &found);
    } else if (left_like) {  // one pair per probe row, at the row's own position
      B200_TIMED("join_part_probe");
      e = launch_probe32_unique<true>(pp, g, t, out, d_pstart, blocks);
      found = pp.n;
    } else {
      cudaError_t ce = holes.alloc((size_t)2 * warps * (sizeof(unsigned long long) + sizeof(unsigned)));
      if (ce == cudaSuccess) ce = plan.alloc(sizeof(FixPlan));
      if (ce != cudaSuccess) e = GDF_CUDA_ERROR;
      if (e == GDF_SUCCESS) {
        out.hole_start = holes.as<unsigned long long>();
        out.hole_len = reinterpret_cast<unsigned*>(out.hole_start + (size_t)2 * warps);
        {
          B200_TIMED("join_part_probe");
          e = launch_probe32_unique<false>(pp, g, t, out, d_pstart, blocks, cap_tiles);
        }
        if (e == GDF_SUCCESS) {
          B200_TIMED("join_output_fixup");
          const int fsmem = kFixCap * (int)(sizeof(unsigned long long) + 3 * sizeof(unsigned));
          cudaFuncSetAttribute(fixup_plan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, fsmem);
          fixup_plan_kernel<<<1, kFixThreads, fsmem>>>(out.hole_start, out.hole_len, 2 * warps, d_cursor, plan.as<FixPlan>());
          fixup_move_kernel<<<sm_count() * 32, 256>>>(plan.as<FixPlan>(), op, ob);  // two dependent binary searches per pair: latency is hidden by threads, not by a loop
          if (cudaPeekAtLastError() != cudaSuccess) e = GDF_CUDA_ERROR;
        }
        if (e == GDF_SUCCESS) e = read_u64(&plan.as<FixPlan>()->found, &found);
      }
    }
  }
  if (e == GDF_SUCCESS && kind == JOIN_FULL) {
    Scratch marks;
    cudaError_t ce = marks.alloc(build_rows);
    if (ce == cudaSuccess) ce = cudaMemsetAsync(marks.ptr, 0, build_rows, 0);
    if (ce == cudaSuccess) {
      ce = cudaMemcpy(d_cursor, &found, sizeof(unsigned long long), cudaMemcpyHostToDevice);  // append position
    }
    if (ce == cudaSuccess) {
      if (found) mark_rows_kernel<<<grid_for((size_t)found), kThreads>>>(ob, (size_t)found, marks.as<unsigned char>(), build_rows);
      append_unmatched_kernel<<<grid_for(build_rows), kThreads>>>(marks.as<unsigned char>(), build_rows, op, ob, d_cursor);
      ce = cudaPeekAtLastError();
    }
    if (ce != cudaSuccess) e = GDF_CUDA_ERROR;
    else e = read_u64(d_cursor, &found);
  }
  if (e != GDF_SUCCESS || found == 0) {
    rmmFree(op, 0);
    rmmFree(ob, 0);
    if (e == GDF_SUCCESS) {
      view_indices(out_l, nullptr, 0);
      view_indices(out_r, nullptr, 0);
    }
    return e;
  }
  view_indices(flip ? out_r : out_l, op, (size_t)found);
  view_indices(flip ? out_l : out_r, ob, (size_t)found);
  return GDF_SUCCESS;
}

template <typename KT>
gdf_error run_partitioned(int kind, const gdf_column* probe_col, const gdf_column* build_col, bool flip,
                          gdf_column* out_l, gdf_column* out_r, bool* handled, const int32_t* probe_payload,
                          const int32_t* build_payload, const gdf_column* probe_col2, const gdf_column* build_col2) {
  const size_t P = probe_col->size, B = build_col->size;
  const bool left_like = kind != JOIN_INNER;
  PartGeom g;
  g.dest = 0;
  g.nlocal = 0;
  {
    const size_t rows_per_part = (size_t)1 << lab_knob("B200_ROWS_PER_PART_LOG2", kRowsPerPartitionLog2);
    unsigned np = pow2_at_least((B + rows_per_part - 1) / rows_per_part);
    if (np > kMaxParts) np = kMaxParts;
    unsigned lg = 0;
    while ((1u << lg) < np) ++lg;
    g.nparts = np;
    g.shift = 32 - lg;
  }
  Scratch small;  // totals[np + 1] | cursors[np] | toffset[np] | cursor, flags, ticket (64 bytes) | pstart[np + 1] | tmask[np]
  const size_t small_bytes = (g.nparts * 4 + 2) * sizeof(unsigned long long) + 64 + g.nparts * sizeof(unsigned);
  B200_CUDA_TRY(small.alloc(small_bytes));
  unsigned long long* d_totals = small.as<unsigned long long>();
  unsigned long long* d_cursors = d_totals + g.nparts + 1;
  unsigned long long* d_toffset = d_cursors + g.nparts;
  unsigned long long* d_cursor = d_toffset + g.nparts;
  int* d_flags = reinterpret_cast<int*>(d_cursor + 1);
  unsigned* d_ticket = reinterpret_cast<unsigned*>(d_cursor + 3);
  unsigned long long* d_pstart = d_cursor + 8;
  unsigned* d_tmask = reinterpret_cast<unsigned*>(d_pstart + g.nparts + 1);
  B200_CUDA_TRY(cudaMemsetAsync(d_cursor, 0, 64, 0));

  Scratch bkeys, brows, pkeys, prows;
  // unpartitioned pairs: row tag = position, or the caller's payload (then masks are not supported)
  Pairs<KT> bp{static_cast<const KT*>(build_col->data), build_payload, build_col->valid, B,
               build_col2 ? static_cast<const uint32_t*>(build_col2->data) : nullptr, build_col2 ? build_col2->valid : nullptr};
  Pairs<KT> pp{static_cast<const KT*>(probe_col->data), probe_payload, probe_col->valid, P,
               probe_col2 ? static_cast<const uint32_t*>(probe_col2->data) : nullptr, probe_col2 ? probe_col2->valid : nullptr};
  Scratch bk2, pk2;
  unsigned long long h_btot[kMaxParts];
  if (g.nparts > 1) {
    unsigned long long h_ptot[kMaxParts], h_cursors[kMaxParts + 1];
    auto scan = [&](const unsigned long long* tot) {
      unsigned long long run = 0;
      for (unsigned p = 0; p < g.nparts; ++p) {
        h_cursors[p] = run;
        run += tot[p];
      }
      return (size_t)run;
    };
    // build-side histogram first: it also tells whether every valid build key fits 32 bits
    unsigned hi_or = 0;
    gdf_error e = partition_hist<KT, false>(build_col, g, d_totals, h_btot, build_col2, &hi_or);
    if (e != GDF_SUCCESS) return e;
    size_t kept = scan(h_btot);
    bool compact = build_col2 == nullptr && (sizeof(KT) == 4 || hi_or == 0) && lab_knob("B200_JOIN_COMPACT", 1) != 0;
    unsigned long long slots32 = 0;
    for (unsigned p = 0; p < g.nparts; ++p) {
      if (h_btot[p] > (1u << 22)) compact = false;  // the compact probe packs {partition, slot < 2^24} into 32 bits
      slots32 += pow2_at_least(h_btot[p] ? 2 * h_btot[p] : 4);
    }
    if (slots32 >= (1ull << 32) - 8) compact = false;  // absolute slot indices are 32-bit in the lean probe
    if (compact) {
      Scratch bpairs, ppairs;
      B200_CUDA_TRY(bpairs.alloc((kept ? kept : 1) * sizeof(uint2)));
      e = partition_scatter32<KT, false>(build_col, g, h_cursors, d_cursors, bpairs.as<uint2>(), build_payload, 0);
      if (e != GDF_SUCCESS) return e;
      const Pairs32 bp32{bpairs.as<uint2>(), kept};
      // INNER join with a plain probe column: NO histogram pass over the probe side, and the tables are filled on a
      // private stream WHILE the probe side is scattered (the build is bound by DRAM-cold random sector reads at low
      // occupancy, the scatter by the LSU and bandwidth: they overlap well; the scatter runs at 2 CTAs per SM so that a
      // build CTA fits beside it).  Partition p gets the fixed region [p * cap, (p + 1) * cap) of the pair array (cap =
      // the even share + 1/16 + 4096 pairs), the scatter's cursors start at the region starts, the probe kernel reads the
      // region ends from the cursors (device to device) and skips the padding.  Both shortcuts are optimistic: if the
      // build keys turn out to have duplicates (the count / cursor probe kernels need the contiguous layout) or a
      // skewed probe column overflows a region (the run is dropped and flagged, never written out of bounds), the
      // probe side is partitioned again with exact counts.  C3: 17.9 -> 15.7 ms (no histogram) -> see DESIGN.md.
      bool padded = kind == JOIN_INNER && probe_col->valid == nullptr && probe_payload == nullptr && P >= (1u << 20) &&
                    lab_knob("B200_PADDED_PROBE", 1) != 0;
      const bool overlap = padded && lab_knob("B200_BUILD_OVERLAP", 1) != 0 && xjoin_side_stream() != nullptr;
      Scratch table;
      Tables32 t{nullptr, nullptr, nullptr};
      cudaEvent_t built = nullptr;
      if (overlap) {
        cudaStream_t side = xjoin_side_stream();
        cudaEvent_t issued = nullptr;
        B200_CUDA_TRY(cudaEventCreateWithFlags(&issued, cudaEventDisableTiming));
        cudaError_t ce = cudaEventRecord(issued, 0);   // the private stream starts after the build side's scatter
        if (ce == cudaSuccess) ce = cudaStreamWaitEvent(side, issued, 0);
        cudaEventDestroy(issued);
        if (ce != cudaSuccess) return GDF_CUDA_ERROR;
        {
          B200_TIMED_ON("join_part_build", side);   // elapsed time on the private stream: it overlaps the probe-side scatter
          e = compact_build(g, bp32, h_btot, d_toffset, d_tmask, d_flags, table, &t, side);
        }
        if (e == GDF_SUCCESS && (cudaEventCreateWithFlags(&built, cudaEventDisableTiming) != cudaSuccess ||
                                 cudaEventRecord(built, side) != cudaSuccess))
          e = GDF_CUDA_ERROR;
        if (e != GDF_SUCCESS) {
          cudaStreamSynchronize(side);
          if (built) cudaEventDestroy(built);
          return e;
        }
      } else {
        B200_TIMED("join_part_build");
        e = compact_build(g, bp32, h_btot, d_toffset, d_tmask, d_flags, table, &t, 0);
        if (e != GDF_SUCCESS) return e;
      }
      *handled = true;
      // everything below that returns early must first make the legacy stream wait for the private one
      auto join_streams = [&]() -> gdf_error {
        if (built == nullptr) return GDF_SUCCESS;
        const cudaError_t ce = cudaStreamWaitEvent(0, built, 0);
        cudaEventDestroy(built);
        built = nullptr;
        return ce == cudaSuccess ? GDF_SUCCESS : GDF_CUDA_ERROR;
      };
      int h_flags[2] = {0, 0};
      if (padded && !overlap) {
        B200_CUDA_TRY(cudaMemcpy(h_flags, d_flags, sizeof(h_flags), cudaMemcpyDeviceToHost));
        if (h_flags[0] != 0) padded = false;  // duplicate build keys
      }
      if (padded) {
        const size_t share = (P + g.nparts - 1) / g.nparts;
        const size_t cap = (share + share / 16 + 4096 + kC32Tile - 1) / kC32Tile * kC32Tile;
        int* d_overflow = reinterpret_cast<int*>(d_cursor + 2);
        B200_CUDA_TRY(ppairs.alloc((size_t)g.nparts * cap * sizeof(uint2)));
        for (unsigned p = 0; p < g.nparts; ++p) h_cursors[p] = (unsigned long long)p * cap;
        PeerPairs reg;
        reg.on = 1;
        reg.shift = 8;  // bin >> 8 == 0: one destination
        reg.status = nullptr;
        reg.region_cap = cap;
        reg.overflow = d_overflow;
        reg.counts = nullptr;
        reg.offsets0 = nullptr;
        for (int r = 0; r < kMaxPeers; ++r) reg.base[r] = nullptr;
        reg.base[0] = ppairs.as<uint2>();
        e = partition_scatter32<KT, false>(probe_col, g, h_cursors, d_cursors, ppairs.as<uint2>(), nullptr, 0, &reg, false,
                                           overlap ? lab_knob("B200_OVERLAP_SCATTER_CTAS", 2) : 0);
        if (e == GDF_SUCCESS) e = join_streams();
        if (e != GDF_SUCCESS) {
          cudaDeviceSynchronize();
          if (built) cudaEventDestroy(built);
          return e;
        }
        // region ends = the cursors after the scatter; [nparts] is not read in this layout
        B200_CUDA_TRY(cudaMemcpyAsync(d_pstart, d_cursors, g.nparts * sizeof(unsigned long long), cudaMemcpyDeviceToDevice, 0));
        bool redo = false;
        if (overlap) {  // the build ran beside the scatter: only now are its flags known
          B200_CUDA_TRY(cudaMemcpy(h_flags, d_flags, sizeof(h_flags), cudaMemcpyDeviceToHost));
          redo = h_flags[0] != 0;
        }
        if (!redo) {
          const Pairs32 pp32{ppairs.as<uint2>(), (size_t)g.nparts * cap};
          e = compact_probe(kind, flip, g, t, pp32, B, d_cursor, d_flags, d_pstart, out_l, out_r, (unsigned)(cap / kC32Tile), P);
          if (e != GDF_SUCCESS) return e;
          int overflow = 0;
          B200_CUDA_TRY(cudaMemcpy(&overflow, d_overflow, sizeof(int), cudaMemcpyDeviceToHost));
          if (!overflow) return GDF_SUCCESS;
          // skewed probe keys: discard the partial result, fall through to the exact path
          if (out_l->data) rmmFree(out_l->data, 0);
          if (out_r->data) rmmFree(out_r->data, 0);
          view_indices(out_l, nullptr, 0);
          view_indices(out_r, nullptr, 0);
        }
        ppairs.release();
        // cursor, overflow flag and ticket start again; the duplicate flag (first int after the cursor) is kept
        B200_CUDA_TRY(cudaMemsetAsync(d_cursor, 0, sizeof(unsigned long long), 0));
        B200_CUDA_TRY(cudaMemsetAsync(d_cursor + 2, 0, 64 - 2 * sizeof(unsigned long long), 0));
      }
      {
        const gdf_error je = join_streams();
        if (je != GDF_SUCCESS) return je;
      }
      e = left_like ? partition_hist<KT, true, true>(probe_col, g, d_totals, h_ptot)
                    : partition_hist<KT, false, true>(probe_col, g, d_totals, h_ptot);
      if (e != GDF_SUCCESS) return e;
      kept = scan(h_ptot);
      B200_CUDA_TRY(ppairs.alloc((kept ? kept : 1) * sizeof(uint2)));
      e = left_like ? partition_scatter32<KT, true>(probe_col, g, h_cursors, d_cursors, ppairs.as<uint2>(), probe_payload, 0)
                    : partition_scatter32<KT, false>(probe_col, g, h_cursors, d_cursors, ppairs.as<uint2>(), probe_payload, 0);
      if (e != GDF_SUCCESS) return e;
      const Pairs32 pp32{ppairs.as<uint2>(), kept};
      h_cursors[g.nparts] = kept;  // h_cursors = first pair of every probe partition
      B200_CUDA_TRY(cudaMemcpy(d_pstart, h_cursors, (g.nparts + 1) * sizeof(unsigned long long), cudaMemcpyHostToDevice));
      return compact_probe(kind, flip, g, t, pp32, B, d_cursor, d_flags, d_pstart, out_l, out_r);
    }
    // general path: {key, tag} (+ second key) in separate arrays, 16-byte slots
    PeerDst none;
    none.on = 0;
    B200_CUDA_TRY(bkeys.alloc((kept ? kept : 1) * sizeof(KT)));
    B200_CUDA_TRY(brows.alloc((kept ? kept : 1) * sizeof(int32_t)));
    if (build_col2) B200_CUDA_TRY(bk2.alloc((kept ? kept : 1) * sizeof(uint32_t)));
    e = partition_scatter<KT, false>(build_col, g, h_cursors, d_cursors, bkeys.as<KT>(), brows.as<int32_t>(), build_payload, 0,
                                     none, build_col2, build_col2 ? bk2.as<uint32_t>() : nullptr);
    if (e != GDF_SUCCESS) return e;
    bp = Pairs<KT>{bkeys.as<KT>(), brows.as<int32_t>(), nullptr, kept, build_col2 ? bk2.as<uint32_t>() : nullptr, nullptr};
    e = left_like ? partition_side<KT, true>(probe_col, g, pkeys, prows, d_totals, d_cursors, h_ptot, &kept, probe_payload, 0,
                                             nullptr, nullptr, probe_col2, &pk2)
                  : partition_side<KT, false>(probe_col, g, pkeys, prows, d_totals, d_cursors, h_ptot, &kept, probe_payload, 0,
                                              nullptr, nullptr, probe_col2, &pk2);
    if (e != GDF_SUCCESS) return e;
    pp = Pairs<KT>{pkeys.as<KT>(), prows.as<int32_t>(), nullptr, kept, probe_col2 ? pk2.as<uint32_t>() : nullptr, nullptr};
  } else {
    h_btot[0] = B;
  }
  // one table per partition, sized from the actual counts (load factor <= 0.5)
  unsigned long long h_toffset[kMaxParts], total_slots = 0;
  unsigned h_tmask[kMaxParts];
  bool stream_ok = probe_col2 == nullptr;  // the streaming probe packs {partition, slot} into 8 + 24 bits; single keys only
  for (unsigned p = 0; p < g.nparts; ++p) {
    const unsigned slots = pow2_at_least(h_btot[p] ? 2 * h_btot[p] : 2);
    if (slots > (1u << 24)) stream_ok = false;
    h_toffset[p] = total_slots;
    h_tmask[p] = slots - 1;
    total_slots += slots;
  }
  B200_CUDA_TRY(cudaMemcpy(d_toffset, h_toffset, g.nparts * sizeof(unsigned long long), cudaMemcpyHostToDevice));
  B200_CUDA_TRY(cudaMemcpy(d_tmask, h_tmask, g.nparts * sizeof(unsigned), cudaMemcpyHostToDevice));
  Scratch table;
  B200_CUDA_TRY(table.alloc(total_slots * sizeof(Slot)));
  {
    B200_TIMED("join_table_init");
    B200_CUDA_TRY(cudaMemsetAsync(table.ptr, 0xff, total_slots * sizeof(Slot), 0));
  }
  Tables t{table.as<Slot>(), d_toffset, d_tmask};
  if (bp.n) {
    B200_TIMED("join_part_build");
    if (bp.k2) part_build_kernel<KT, true><<<(unsigned)((bp.n + kBuildTile - 1) / kBuildTile), kThreads>>>(bp, g, t, d_flags);
    else part_build_kernel<KT, false><<<(unsigned)((bp.n + kBuildTile - 1) / kBuildTile), kThreads>>>(bp, g, t, d_flags);
    B200_CHECK_LAST();
  }
  int h_flags[2] = {0, 0};
  B200_CUDA_TRY(cudaMemcpy(h_flags, d_flags, sizeof(h_flags), cudaMemcpyDeviceToHost));
  if (h_flags[1]) {  // a build key equals the EMPTY pattern: let the generic path handle this input
    *handled = false;
    return GDF_SUCCESS;
  }
  brows.release();  // build pairs are in the tables now
  bkeys.release();
  bk2.release();
  const bool unique = h_flags[0] == 0;
  size_t bound = pp.n;
  gdf_error e = GDF_SUCCESS;
  if (!unique) {
    e = left_like ? launch_probe<KT, true>(false, false, pp, g, t, nullptr, nullptr, d_cursor, d_ticket, stream_ok)
                  : launch_probe<KT, false>(false, false, pp, g, t, nullptr, nullptr, d_cursor, d_ticket, stream_ok);
    if (e != GDF_SUCCESS) return e;
    unsigned long long exact = 0;
    if ((e = read_u64(d_cursor, &exact)) != GDF_SUCCESS) return e;
    bound = (size_t)exact;
    B200_CUDA_TRY(cudaMemsetAsync(d_cursor, 0, sizeof(unsigned long long), 0));
  }
  *handled = true;
  const size_t capacity = bound + (kind == JOIN_FULL ? B : 0);
  if (capacity == 0) {
    view_indices(out_l, nullptr, 0);
    view_indices(out_r, nullptr, 0);
    return GDF_SUCCESS;
  }
  int32_t *op = nullptr, *ob = nullptr;
  B200_RMM_TRY(output_alloc((void**)&op, capacity * sizeof(int32_t)));
  if (output_alloc((void**)&ob, capacity * sizeof(int32_t)) != RMM_SUCCESS) {
    rmmFree(op, 0);
    return GDF_MEMORYMANAGER_ERROR;
  }
  e = left_like ? launch_probe<KT, true>(unique, true, pp, g, t, op, ob, d_cursor, d_ticket, stream_ok)
                : launch_probe<KT, false>(unique, true, pp, g, t, op, ob, d_cursor, d_ticket, stream_ok);
  unsigned long long found = 0;
  if (e == GDF_SUCCESS) e = read_u64(d_cursor, &found);
  if (e == GDF_SUCCESS && kind == JOIN_FULL) {
    Scratch marks;
    cudaError_t ce = marks.alloc(B);
    if (ce == cudaSuccess) ce = cudaMemsetAsync(marks.ptr, 0, B, 0);
    if (ce == cudaSuccess) {
      if (found) mark_rows_kernel<<<grid_for((size_t)found), kThreads>>>(ob, (size_t)found, marks.as<unsigned char>(), B);
      append_unmatched_kernel<<<grid_for(B), kThreads>>>(marks.as<unsigned char>(), B, op, ob, d_cursor);
      ce = cudaPeekAtLastError();
    }
    if (ce != cudaSuccess) e = GDF_CUDA_ERROR;
    else e = read_u64(d_cursor, &found);
  }
  if (e != GDF_SUCCESS || found == 0) {
    rmmFree(op, 0);
    rmmFree(ob, 0);
    if (e == GDF_SUCCESS) {
      view_indices(out_l, nullptr, 0);
      view_indices(out_r, nullptr, 0);
    }
    return e;
  }
  view_indices(flip ? out_r : out_l, op, (size_t)found);
  view_indices(flip ? out_l : out_r, ob, (size_t)found);
  return GDF_SUCCESS;
}

}  // namespace

gdf_error partitioned_join(int kind, const gdf_column* probe_key, const gdf_column* build_key, bool flip,
                           gdf_column* out_l, gdf_column* out_r, bool* handled, const int32_t* probe_payload,
                           const int32_t* build_payload, const gdf_column* probe_key2, const gdf_column* build_key2) {
  // probe_key2 / build_key2: optional SECOND key column of 4-byte integers (composite keys such as C5's
  // (int64,int32)); it rides in the slot's spare word.  Anything else goes to the generic path.
  *handled = false;
  if (probe_key2 && !(probe_key2->dtype == GDF_INT32 || probe_key2->dtype == GDF_DATE32)) return GDF_SUCCESS;
  switch (probe_key->dtype) {
    case GDF_INT64: case GDF_DATE64: case GDF_TIMESTAMP:
      return run_partitioned<uint64_t>(kind, probe_key, build_key, flip, out_l, out_r, handled, probe_payload, build_payload,
                                       probe_key2, build_key2);
    case GDF_INT32: case GDF_DATE32:
      return run_partitioned<uint32_t>(kind, probe_key, build_key, flip, out_l, out_r, handled, probe_payload, build_payload,
                                       probe_key2, build_key2);
    default: return GDF_SUCCESS;  // floats / narrow ints: generic path
  }
}

// Multi-GPU layer: split one key column into `num_partitions` destination ranges of {key, global row id}
// pairs (ids = id_base + position).  h_offsets receives num_partitions + 1 start offsets (host).
template <typename KT>
gdf_error partition_pairs_typed(const gdf_column* key, int32_t id_base, unsigned num_partitions, void* out_keys,
                                int32_t* out_ids, unsigned long long* h_offsets) {
  PartGeom g;
  g.nparts = num_partitions;
  g.shift = 0;
  g.dest = 1;
  g.nlocal = 0;
  Scratch small, unused_a, unused_b;
  B200_CUDA_TRY(small.alloc((2 * (size_t)num_partitions + 1) * sizeof(unsigned long long)));
  unsigned long long h_tot[kMaxParts];
  size_t kept = 0;
  gdf_error e = partition_side<KT, false>(key, g, unused_a, unused_b, small.as<unsigned long long>(),
                                          small.as<unsigned long long>() + num_partitions + 1, h_tot, &kept, nullptr, id_base,
                                          static_cast<KT*>(out_keys), out_ids);
  if (e != GDF_SUCCESS) return e;
  unsigned long long run = 0;
  for (unsigned p = 0; p < num_partitions; ++p) {
    h_offsets[p] = run;
    run += h_tot[p];
  }
  h_offsets[num_partitions] = run;
  return GDF_SUCCESS;
}

// Multi-GPU layer, fused path, step 1: rows per destination (host array [num_partitions]).
gdf_error partition_count(const gdf_column* key, unsigned num_partitions, unsigned long long* h_counts) {
  B200_REQUIRE(num_partitions >= 1 && num_partitions <= kMaxParts, GDF_INVALID_API_CALL);
  PartGeom g;
  g.nparts = num_partitions;
  g.shift = 0;
  g.dest = 1;
  g.nlocal = 0;
  Scratch small;
  B200_CUDA_TRY(small.alloc(((size_t)num_partitions + 1) * sizeof(unsigned long long)));
  switch (key->dtype) {
    case GDF_INT64: case GDF_DATE64: case GDF_TIMESTAMP:
      return partition_hist<uint64_t, false>(key, g, small.as<unsigned long long>(), h_counts);
    case GDF_INT32: case GDF_DATE32:
      return partition_hist<uint32_t, false>(key, g, small.as<unsigned long long>(), h_counts);
    default: return GDF_UNSUPPORTED_DTYPE;
  }
}

// Step 2: scatter {key, id} pairs straight into the destinations' buffers (local or peer pointers).
gdf_error partition_scatter_peer(const gdf_column* key, int32_t id_base, unsigned num_partitions, void* const* dst_keys,
                                 int32_t* const* dst_ids, const unsigned long long* dst_offsets) {
  B200_REQUIRE(num_partitions >= 1 && num_partitions <= (unsigned)kMaxPeers, GDF_INVALID_API_CALL);
  PartGeom g;
  g.nparts = num_partitions;
  g.shift = 0;
  g.dest = 1;
  g.nlocal = 0;
  PeerDst peer;
  peer.on = 1;
  for (int p = 0; p < kMaxPeers; ++p) {
    peer.keys[p] = p < (int)num_partitions ? dst_keys[p] : nullptr;
    peer.ids[p] = p < (int)num_partitions ? dst_ids[p] : nullptr;
  }
  Scratch small;
  B200_CUDA_TRY(small.alloc((size_t)num_partitions * sizeof(unsigned long long)));
  switch (key->dtype) {
    case GDF_INT64: case GDF_DATE64: case GDF_TIMESTAMP:
      return partition_scatter<uint64_t, false>(key, g, dst_offsets, small.as<unsigned long long>(), nullptr, nullptr,
                                                nullptr, id_base, peer);
    case GDF_INT32: case GDF_DATE32:
      return partition_scatter<uint32_t, false>(key, g, dst_offsets, small.as<unsigned long long>(), nullptr, nullptr,
                                                nullptr, id_base, peer);
    default: return GDF_UNSUPPORTED_DTYPE;
  }
}

// ---- fused exchange, second generation (multi-GPU layer): ONE partition pass per side ----
// The sender's histogram / scatter use the COMBINED geometry (destination rank x the receiver's local partition), so
// the pairs arrive in the receiver's buffers already partition-contiguous and in the compact {key32, tag32} form;
// the receiver goes straight to table build + probe (run_compact) without a histogram or scatter of its own.
// Round 1 scattered every row twice (by destination over NVLink, then again locally): 4.2 of 7.8 ms at 8 GPUs.
PartGeom combined_geom(unsigned ranks, unsigned nlocal) {
  PartGeom g;
  g.nparts = ranks * nlocal;
  g.dest = 2;
  g.nlocal = nlocal;
  unsigned lg = 0;
  while ((1u << lg) < nlocal) ++lg;
  g.shift = 32 - lg;
  return g;
}

// counts[ranks * nlocal] (host) = rows of `key` per (destination, local partition); *hi_or = OR of the high words of
// the keys (0 for 4-byte keys): the compact exchange applies when it is 0 on the build side of every rank.
gdf_error xjoin_count(const gdf_column* key, unsigned ranks, unsigned nlocal, unsigned long long* h_counts, unsigned* hi_or) {
  B200_REQUIRE(ranks >= 1 && ranks <= (unsigned)kMaxPeers && nlocal >= 1 && (nlocal & (nlocal - 1)) == 0 &&
                   ranks * nlocal <= kMaxParts, GDF_INVALID_API_CALL);
  const PartGeom g = combined_geom(ranks, nlocal);
  Scratch small;
  B200_CUDA_TRY(small.alloc(((size_t)g.nparts + 1) * sizeof(unsigned long long)));
  switch (key->dtype) {
    case GDF_INT64: case GDF_DATE64: case GDF_TIMESTAMP:
      return partition_hist<uint64_t, false, true>(key, g, small.as<unsigned long long>(), h_counts, nullptr, hi_or);
    case GDF_INT32: case GDF_DATE32:
      return partition_hist<uint32_t, false, true>(key, g, small.as<unsigned long long>(), h_counts, nullptr, hi_or);
    default: return GDF_UNSUPPORTED_DTYPE;
  }
}

// dst_pairs[r] = rank r's pair buffer as mapped on THIS device; h_offsets[r * nlocal + p] = position inside it of this
// rank's first pair of local partition p.  Rows whose key does not fit 32 bits are dropped (they cannot match).
gdf_error xjoin_scatter(const gdf_column* key, int32_t id_base, unsigned ranks, unsigned nlocal, void* const* dst_pairs,
                        const unsigned long long* offsets, const unsigned long long* counts,
                        const int* d_status /* non-null: offsets and counts are DEVICE arrays */, int ctas_per_sm) {
  B200_REQUIRE(ranks >= 1 && ranks <= (unsigned)kMaxPeers && nlocal >= 1 && (nlocal & (nlocal - 1)) == 0 &&
                   ranks * nlocal <= kMaxParts, GDF_INVALID_API_CALL);
  const PartGeom g = combined_geom(ranks, nlocal);
  PeerPairs peer;
  peer.on = 1;
  peer.shift = 32 - g.shift;
  peer.status = d_status;
  peer.region_cap = 0;
  peer.overflow = nullptr;
  for (int r = 0; r < kMaxPeers; ++r) peer.base[r] = r < (int)ranks ? static_cast<uint2*>(dst_pairs[r]) : nullptr;
  Scratch small;  // cursors[np] | tail cursors[np] | counts[np] | offsets[np]
  B200_CUDA_TRY(small.alloc((size_t)g.nparts * 4 * sizeof(unsigned long long)));
  const bool dev = d_status != nullptr;  // asynchronous exchange: the plan lives on the device
  if (dev) {
    peer.counts = counts;
    peer.offsets0 = offsets;
  } else {  // host plan: the kernel still wants its own device copies (the cursors are consumed)
    unsigned long long* d_counts = small.as<unsigned long long>() + 2 * (size_t)g.nparts;
    unsigned long long* d_off0 = d_counts + g.nparts;
    B200_CUDA_TRY(cudaMemcpy(d_counts, counts, g.nparts * sizeof(unsigned long long), cudaMemcpyHostToDevice));
    B200_CUDA_TRY(cudaMemcpy(d_off0, offsets, g.nparts * sizeof(unsigned long long), cudaMemcpyHostToDevice));
    peer.counts = d_counts;
    peer.offsets0 = d_off0;
  }
  switch (key->dtype) {
    case GDF_INT64: case GDF_DATE64: case GDF_TIMESTAMP:
      return partition_scatter32<uint64_t, false>(key, g, offsets, small.as<unsigned long long>(), nullptr, nullptr, id_base, &peer, dev, ctas_per_sm);
    case GDF_INT32: case GDF_DATE32:
      return partition_scatter32<uint32_t, false>(key, g, offsets, small.as<unsigned long long>(), nullptr, nullptr, id_base, &peer, dev, ctas_per_sm);
    default: return GDF_UNSUPPORTED_DTYPE;
  }
}

// ---- asynchronous variant of the exchange: counts and plan stay on the device, the host never waits between the
// histogram and the scatter (round 2, first version: ~1.2 of 5.5 ms at 8 GPUs were host round trips) ----
// d_counts[ranks * nlocal + 1]: rows per bin, then the OR of the keys' high words (as a 64-bit cell).
gdf_error xjoin_count_dev(const gdf_column* key, unsigned ranks, unsigned nlocal, unsigned long long* d_counts) {
  B200_REQUIRE(ranks >= 1 && ranks <= (unsigned)kMaxPeers && nlocal >= 1 && (nlocal & (nlocal - 1)) == 0 &&
                   ranks * nlocal <= kMaxParts, GDF_INVALID_API_CALL);
  const PartGeom g = combined_geom(ranks, nlocal);
  if (key->size == 0) {
    B200_CUDA_TRY(cudaMemsetAsync(d_counts, 0, ((size_t)g.nparts + 1) * sizeof(unsigned long long), 0));
    return GDF_SUCCESS;
  }
  switch (key->dtype) {
    case GDF_INT64: case GDF_DATE64: case GDF_TIMESTAMP:
      return partition_hist<uint64_t, false, true>(key, g, d_counts, nullptr);
    case GDF_INT32: case GDF_DATE32:
      return partition_hist<uint32_t, false, true>(key, g, d_counts, nullptr);
    default: return GDF_UNSUPPORTED_DTYPE;
  }
}

// all[s * stride + side * (bins + 1) + b] = pairs rank s sends to bin b of `side` (0 build, 1 probe), cell [bins] = OR of
// the side's high key words.  One CTA, one thread per bin.  Receiver d lays its buffer out partition-major and, inside a
// partition, in sender order (dist.plan_fused_exchange is the host statement of the same arithmetic):
//   off[side][b]   where THIS rank's pairs of bin b start inside the destination's buffer
//   status[0] = 1  a build key somewhere does not fit 32 bits (compact exchange not applicable)
//   status[1] = 1  some rank would receive more pairs than its buffer holds (cap_build / cap_probe)
__global__ void xjoin_plan_kernel(const unsigned long long* __restrict__ all, unsigned ranks, unsigned nlocal, unsigned rank,
                                  unsigned long long cap_build, unsigned long long cap_probe,
                                  unsigned long long* __restrict__ off_build, unsigned long long* __restrict__ off_probe,
                                  int* __restrict__ status) {
  __shared__ unsigned long long tot[kMaxParts];
  const unsigned bins = ranks * nlocal, stride = 2 * (bins + 1);
  const unsigned b = threadIdx.x;
  if (b == 0) {
    unsigned long long wide = 0;
    for (unsigned s = 0; s < ranks; ++s) wide |= all[(size_t)s * stride + bins];
    if (wide) status[0] = 1;
  }
  for (unsigned side = 0; side < 2; ++side) {
    unsigned long long before_me = 0, total = 0;
    if (b < bins)
      for (unsigned s = 0; s < ranks; ++s) {  // every (sender, bin) slot holds whole 32-byte granules (part_scatter32_even_kernel)
        const unsigned long long c = (all[(size_t)s * stride + side * (bins + 1) + b] + kXG - 1) & ~(unsigned long long)(kXG - 1);
        if (s < rank) before_me += c;
        total += c;
      }
    __syncthreads();  // the previous side's reads of tot[] are done
    if (b < bins) tot[b] = total;
    __syncthreads();
    if (b < bins) {
      const unsigned d = b / nlocal;
      unsigned long long start = 0;
      for (unsigned q = d * nlocal; q < b; ++q) start += tot[q];
      (side ? off_probe : off_build)[b] = start + before_me;
      if (b % nlocal == nlocal - 1 && start + total > (side ? cap_probe : cap_build)) status[1] = 1;
    }
  }
}

gdf_error xjoin_plan_dev(const unsigned long long* d_all, unsigned ranks, unsigned nlocal, unsigned rank, unsigned long long cap_build,
                         unsigned long long cap_probe, unsigned long long* d_off_build, unsigned long long* d_off_probe,
                         int* d_status) {
  B200_REQUIRE(ranks >= 1 && ranks <= (unsigned)kMaxPeers && nlocal >= 1 && ranks * nlocal <= kMaxParts && rank < ranks,
               GDF_INVALID_API_CALL);
  B200_CUDA_TRY(cudaMemsetAsync(d_status, 0, 2 * sizeof(int), 0));
  xjoin_plan_kernel<<<1, kMaxParts>>>(d_all, ranks, nlocal, rank, cap_build, cap_probe, d_off_build, d_off_probe, d_status);
  B200_CHECK_LAST();
  return GDF_SUCCESS;
}

// INNER join of partition-contiguous compact pairs (this rank's receive buffers): counts[p] = pairs of local partition p.
// Two stages so that the multi-GPU layer can fill the tables while the probe side is still crossing NVLink:
//   xjoin_build  ordered after everything issued so far on the legacy stream, runs on a private non-blocking stream
//                (side = true) or on the legacy stream itself, returns a handle
//   xjoin_probe  makes the legacy stream wait for the build, probes, writes the outputs, releases the handle
struct XJoinHandle {
  Scratch small, table;
  Tables32 t{nullptr, nullptr, nullptr};
  PartGeom g;
  unsigned long long build_rows = 0;
  unsigned long long* d_cursor = nullptr;
  int* d_flags = nullptr;
  unsigned long long* d_pstart = nullptr;
  cudaEvent_t done = nullptr;
  bool side = false;
};

gdf_error xjoin_build(const void* build_pairs, const unsigned long long* build_counts, unsigned nlocal, bool side, void** handle) {
  B200_REQUIRE(nlocal >= 1 && nlocal <= kMaxParts && (nlocal & (nlocal - 1)) == 0, GDF_INVALID_API_CALL);
  *handle = nullptr;
  unsigned long long build_rows = 0, slots32 = 0;
  for (unsigned p = 0; p < nlocal; ++p) {
    build_rows += build_counts[p];
    B200_REQUIRE(build_counts[p] <= (1u << 22), GDF_COLUMN_SIZE_TOO_BIG);  // limits of the compact probe (run_partitioned)
    slots32 += pow2_at_least(build_counts[p] ? 2 * build_counts[p] : 4);
  }
  B200_REQUIRE(slots32 < (1ull << 32) - 8, GDF_COLUMN_SIZE_TOO_BIG);
  XJoinHandle* h = new XJoinHandle();
  h->g.nparts = nlocal;
  h->g.dest = 0;
  h->g.nlocal = 0;
  unsigned lg = 0;
  while ((1u << lg) < nlocal) ++lg;
  h->g.shift = 32 - lg;
  h->build_rows = build_rows;
  h->side = side;
  // toffset[np] | cursor, flags (64 bytes) | pstart[np + 1] | tmask[np]
  cudaError_t ce = h->small.alloc(((size_t)nlocal * 2 + 1) * sizeof(unsigned long long) + 64 + nlocal * sizeof(unsigned));
  cudaStream_t s = 0;
  if (ce == cudaSuccess && side) {
    s = xjoin_side_stream();
    if (s == nullptr) ce = cudaErrorUnknown;
    if (ce == cudaSuccess) ce = cudaEventCreateWithFlags(&h->done, cudaEventDisableTiming);
    if (ce == cudaSuccess) {   // order the private stream after everything the caller has issued (the build side's exchange)
      cudaEvent_t issued = nullptr;
      ce = cudaEventCreateWithFlags(&issued, cudaEventDisableTiming);
      if (ce == cudaSuccess) ce = cudaEventRecord(issued, 0);
      if (ce == cudaSuccess) ce = cudaStreamWaitEvent(s, issued, 0);
      if (issued) cudaEventDestroy(issued);
    }
  }
  if (ce != cudaSuccess) {
    delete h;
    return GDF_CUDA_ERROR;
  }
  unsigned long long* d_toffset = h->small.as<unsigned long long>();
  h->d_cursor = d_toffset + nlocal;
  h->d_flags = reinterpret_cast<int*>(h->d_cursor + 1);
  h->d_pstart = h->d_cursor + 8;
  unsigned* d_tmask = reinterpret_cast<unsigned*>(h->d_pstart + nlocal + 1);
  gdf_error e = GDF_SUCCESS;
  if (cudaMemsetAsync(h->d_cursor, 0, 64, s) != cudaSuccess) e = GDF_CUDA_ERROR;
  if (e == GDF_SUCCESS && build_rows) {
    const Pairs32 bp{static_cast<const uint2*>(build_pairs), (size_t)build_rows};
    if (side) {
      B200_TIMED_ON("join_part_build", s);
      e = compact_build(h->g, bp, build_counts, d_toffset, d_tmask, h->d_flags, h->table, &h->t, s);
    } else {
      B200_TIMED("join_part_build");
      e = compact_build(h->g, bp, build_counts, d_toffset, d_tmask, h->d_flags, h->table, &h->t, s);
    }
  }
  if (e == GDF_SUCCESS && side && cudaEventRecord(h->done, s) != cudaSuccess) e = GDF_CUDA_ERROR;
  if (e != GDF_SUCCESS) {
    if (side) cudaStreamSynchronize(s);   // nothing of the handle's memory may still be in use when it is released
    if (h->done) cudaEventDestroy(h->done);
    delete h;
    return e;
  }
  *handle = h;
  return GDF_SUCCESS;
}

gdf_error xjoin_probe(void* handle, const void* probe_pairs, const unsigned long long* probe_counts, gdf_column* out_l,
                      gdf_column* out_r) {
  XJoinHandle* h = static_cast<XJoinHandle*>(handle);
  const unsigned nlocal = h->g.nparts;
  unsigned long long h_pstart[kMaxParts + 1], probe_rows = 0;
  for (unsigned p = 0; p < nlocal; ++p) {
    h_pstart[p] = probe_rows;
    probe_rows += probe_counts[p];
  }
  h_pstart[nlocal] = probe_rows;
  gdf_error e = GDF_SUCCESS;
  if (h->side && cudaStreamWaitEvent(0, h->done, 0) != cudaSuccess) e = GDF_CUDA_ERROR;
  view_indices(out_l, nullptr, 0);
  view_indices(out_r, nullptr, 0);
  if (e == GDF_SUCCESS && probe_rows != 0 && h->build_rows != 0) {
    if (cudaMemcpy(h->d_pstart, h_pstart, (nlocal + 1) * sizeof(unsigned long long), cudaMemcpyHostToDevice) != cudaSuccess)
      e = GDF_CUDA_ERROR;
    const Pairs32 pp{static_cast<const uint2*>(probe_pairs), (size_t)probe_rows};
    if (e == GDF_SUCCESS)
      e = compact_probe(JOIN_INNER, false, h->g, h->t, pp, (size_t)h->build_rows, h->d_cursor, h->d_flags, h->d_pstart, out_l, out_r);
  }
  if (h->side) {
    cudaStreamSynchronize(xjoin_side_stream());  // (already complete: the legacy stream waited for it)
    cudaEventDestroy(h->done);
  }
  delete h;   // scratch goes back to the cache; every user of it was ordered on, or awaited by, the legacy stream
  return e;
}

gdf_error xjoin_local(const void* probe_pairs, const unsigned long long* probe_counts, const void* build_pairs,
                      const unsigned long long* build_counts, unsigned nlocal, gdf_column* out_l, gdf_column* out_r) {
  void* h = nullptr;
  const gdf_error e = xjoin_build(build_pairs, build_counts, nlocal, false, &h);
  if (e != GDF_SUCCESS) return e;
  return xjoin_probe(h, probe_pairs, probe_counts, out_l, out_r);
}

gdf_error partition_pairs(const gdf_column* key, int32_t id_base, unsigned num_partitions, void* out_keys,
                          int32_t* out_ids, unsigned long long* h_offsets) {
  B200_REQUIRE(num_partitions >= 1 && num_partitions <= kMaxParts, GDF_INVALID_API_CALL);
  switch (key->dtype) {
    case GDF_INT64: case GDF_DATE64: case GDF_TIMESTAMP:
      return partition_pairs_typed<uint64_t>(key, id_base, num_partitions, out_keys, out_ids, h_offsets);
    case GDF_INT32: case GDF_DATE32:
      return partition_pairs_typed<uint32_t>(key, id_base, num_partitions, out_keys, out_ids, h_offsets);
    default: return GDF_UNSUPPORTED_DTYPE;
  }
}

}  // namespace b200
