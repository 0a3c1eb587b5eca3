// Compact (32-bit key) kernels of the radix-partitioned join.  Included by join_part.cu inside its anonymous
// namespace, after KeyBits / slot_hash / PartGeom.
//
// When every valid BUILD key fits in 32 bits (int32 key columns always; int64 key columns whose high words are all
// zero - detected for free by the histogram pass: ids, C3) a probe key with a non-zero high word cannot match, so
//   * the partitioned pairs are {key32, tag32} = 8 bytes instead of 12 (scatter writes and probe reads shrink by a
//     third),
//   * a table slot is the 8-byte word {key32 | row32 << 32}: insert = ONE 64-bit CAS (no second store), and four
//     slots fill one 32-byte L2 sector.  Tables are BUCKETISED: a key's home is an aligned group of four slots
//     fetched with one 256-bit load; at load <= 0.5 the key (or the EMPTY slot that proves its absence) is in
//     the home bucket for ~98 % of the rows, so a probe is exactly one L2 round trip.  EMPTY = all ones: the row
//     word of a real entry is a row id < 2^31.
//
// Probe kernel (profiles/r02_notes.md; what round 1 established: a CTA barrier + one returning cursor atomic per
// 1-2 K-row tile caps a kernel at ~2 TB/s of its stream however the tile is fed):
//   * WARP-INDEPENDENT: no __syncthreads after the prologue.  Every warp owns a 4-stage ring of 256-pair (2 KB)
//     tiles filled by cp.async.bulk (TMA) with an mbarrier per stage; lane 0 is the producer, the warp releases a
//     stage with __syncwarp once all lanes hold their pairs in registers.  Tiles are dealt round-robin over all
//     warps of the grid, so the warps in flight work inside a window of ~600 K consecutive pairs - one or two
//     partitions, whose tables stay L2-resident;
//   * 8 bucket loads in flight per lane; stragglers (home bucket full of other keys, ~2 %) are resolved in warp-wide
//     rounds;
//   * ranks inside the warp come from ballots; output space comes from PER-WARP CHUNKS of the output arrays
//     (2048 pairs, one cursor atomic per chunk, requested one tile before it is needed, so no atomic is on a warp's
//     critical path).  A chunk is filled densely, a tile that crosses a chunk boundary is split.  What stays
//     unused at the end (at most two partial chunks per warp) is recorded as HOLES, and a fix-up pass moves the
//     tail of the output into them (<= 4736 x 2048 pairs, ~20 us) so that the caller sees one dense column;
//   * LEFT joins with unique build keys emit exactly one pair per probe row: the output position is the pair's
//     position, no allocation at all.  Duplicate build keys: exact count pass, then one cursor atomic per warp tile.
#pragma once

constexpr int kC32Threads = 512;
constexpr int kC32Warps = kC32Threads / 32;
constexpr int kC32Rows = 8;                      // pairs per lane per tile
constexpr int kC32Tile = 32 * kC32Rows;          // 256 pairs = 2 KB
constexpr int kC32Stages = 4;
constexpr unsigned kC32Chunk = 2048;             // output pairs per chunk
constexpr unsigned long long kEmpty32 = ~0ull;

struct Tables32 {
  unsigned long long* slots;           // {key32 | row32 << 32}
  const unsigned long long* offset;    // [nparts] first slot of partition p
  const unsigned* mask;                // [nparts] slots_p - 1, slots_p a power of two >= 4
};

struct Pairs32 {
  const uint2* pairs;  // .x = key, .y = tag (row id; ~row id for rows that can never match: NULL key / wide key)
  size_t n;
};

struct Bucket32 {
  unsigned long long w[4];
};
static __device__ __forceinline__ Bucket32 ld_bucket32(const unsigned long long* p) {  // 32-byte aligned; L2 only (tables
  Bucket32 b;                                                                         // are written by the build kernel)
  asm volatile("ld.global.cg.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(b.w[0]), "=l"(b.w[1]), "=l"(b.w[2]), "=l"(b.w[3]) : "l"(p));
  return b;
}
static __device__ __forceinline__ Bucket32 ld_bucket32_hint(const unsigned long long* p, uint64_t policy) {
  Bucket32 b;
  asm volatile("ld.global.L2::cache_hint.v4.u64 {%0,%1,%2,%3}, [%4], %5;"
               : "=l"(b.w[0]), "=l"(b.w[1]), "=l"(b.w[2]), "=l"(b.w[3])
               : "l"(p), "l"(policy));
  return b;
}
static __device__ __forceinline__ bool slot_empty(unsigned long long w) { return (uint32_t)(w >> 32) == 0xffffffffu; }

// ---- build: one 64-bit CAS per row into the first EMPTY slot of its bucket ----
// One CTA = one contiguous tile of pairs, tiles dispatched in index order (the pairs are partition-contiguous, so
// the CTAs in flight insert into one or two partitions' tables, which stay L2-resident).  Every thread keeps four
// inserts in flight and resolves them in rounds: (a) load the bucket of every row that needs one, (b) pick the
// first EMPTY slot of the local copy, (c) issue all CAS, (d) examine - a failed CAS returns the occupant, which
// is checked for "same key" = duplicate build key and marks the slot as taken (no reload).  Two equal keys walk the
// same slot sequence and slots never empty, so the later one always sees the earlier one: flags[0] is exact.
constexpr int kB32Threads = 256;
constexpr int kB32U = 4;
constexpr int kB32Tile = kB32Threads * kB32U;

__global__ void __launch_bounds__(kB32Threads)
build32_kernel(Pairs32 b, PartGeom g, Tables32 t, int* __restrict__ flags /*[0]=dup*/) {
  const size_t tile_lo = (size_t)blockIdx.x * kB32Tile;
  unsigned long long mine[kB32U], prev[kB32U];
  unsigned long long* tab[kB32U];
  unsigned at[kB32U], mask[kB32U];
  unsigned taken[kB32U];  // slots of the current bucket that a failed CAS showed to be occupied (bit j)
  int cand[kB32U];
  Bucket32 bk[kB32U];
  unsigned pend = 0, need = 0;
#pragma unroll
  for (int u = 0; u < kB32U; ++u) {
    const size_t i = tile_lo + (size_t)u * kB32Threads + threadIdx.x;
    mine[u] = 0;
    tab[u] = t.slots;
    at[u] = 0;
    mask[u] = 3;
    taken[u] = 0;
    if (i < b.n && b.pairs[i].y != 0x80000000u) {  // tag INT_MIN = the exchange's "no row" pad pair
      const uint2 pr = b.pairs[i];
      mine[u] = ((unsigned long long)pr.y << 32) | pr.x;
      const uint32_t h = KeyBits<uint32_t>::hash(pr.x);
      const unsigned p = g.pid(h);
      tab[u] = t.slots + t.offset[p];
      mask[u] = t.mask[p];
      at[u] = (slot_hash(h) & mask[u]) & ~3u;
      pend |= 1u << u;
    }
  }
  need = pend;
  bool dup = false;
  while (pend) {
#pragma unroll
    for (int u = 0; u < kB32U; ++u)
      if ((need >> u) & 1u) {
        bk[u] = ld_bucket32(tab[u] + at[u]);
        taken[u] = 0;
      }
    // keys already in the bucket are compared once, right after the load
#pragma unroll
    for (int u = 0; u < kB32U; ++u) {
      if (!((need >> u) & 1u)) continue;
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (!slot_empty(bk[u].w[j]) && (uint32_t)bk[u].w[j] == (uint32_t)mine[u]) dup = true;
    }
    need = 0;
#pragma unroll
    for (int u = 0; u < kB32U; ++u) {
      cand[u] = -1;
      if (!((pend >> u) & 1u)) continue;
#pragma unroll
      for (int j = 3; j >= 0; --j)
        if (slot_empty(bk[u].w[j]) && !((taken[u] >> j) & 1u)) cand[u] = j;
      if (cand[u] < 0) {  // bucket full of other keys: next bucket, loaded in the next round
        at[u] = (at[u] + 4u) & mask[u];
        need |= 1u << u;
      }
    }
#pragma unroll
    for (int u = 0; u < kB32U; ++u)
      if (((pend >> u) & 1u) && cand[u] >= 0) prev[u] = atomicCAS(tab[u] + at[u] + cand[u], kEmpty32, mine[u]);
#pragma unroll
    for (int u = 0; u < kB32U; ++u) {
      if (!((pend >> u) & 1u) || cand[u] < 0) continue;
      if (prev[u] == kEmpty32) {
        pend &= ~(1u << u);
      } else {  // somebody else took the slot since the bucket was loaded: note it, compare, try the next EMPTY slot
        taken[u] |= 1u << cand[u];
        if ((uint32_t)prev[u] == (uint32_t)mine[u]) dup = true;
      }
    }
  }
  if (dup) flags[0] = 1;
}

// ---- probe ----
enum Probe32Mode { P32_COUNT = 0, P32_CURSOR = 1, P32_CHUNK = 2, P32_POSITIONAL = 3 };

struct Probe32Out {
  int32_t* probe;
  int32_t* build;
  unsigned long long* cursor;        // COUNT: total matches; CURSOR: output cursor; CHUNK: chunk cursor (multiples of kC32Chunk)
  unsigned long long* hole_start;    // CHUNK: [2 * warps of the grid]
  unsigned* hole_len;
};

struct Probe32Smem {
  uint64_t bar[kC32Warps][kC32Stages];
  unsigned long long part_off[kMaxParts];
  unsigned part_mask[kMaxParts];
};
constexpr size_t probe32_smem_bytes() {
  return (size_t)kC32Warps * kC32Stages * kC32Tile * sizeof(uint2) + sizeof(Probe32Smem) + 128;
}

static __device__ __forceinline__ void st_i32_stream(int32_t* p, int32_t v, uint64_t policy) {
  asm volatile("st.global.L2::cache_hint.b32 [%0], %1, %2;" ::"l"(p), "r"(v), "l"(policy) : "memory");
}
static __device__ __forceinline__ void bulk_load_policy(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar,
                                                        uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
          tma::smem_u32(smem_dst)),
      "l"(gmem_src), "r"(bytes), "r"(tma::smem_u32(bar)), "l"(policy)
      : "memory");
}

template <bool LEFT_LIKE, bool UNIQUE, int MODE>
__global__ void __launch_bounds__(kC32Threads, 1)
probe32_kernel(Pairs32 pr, PartGeom g, Tables32 t, Probe32Out out, unsigned lab) {
  extern __shared__ __align__(128) unsigned char probe32_smem[];
  uint2* const ring_all = reinterpret_cast<uint2*>(probe32_smem);
  Probe32Smem& sm = *reinterpret_cast<Probe32Smem*>(probe32_smem + (size_t)kC32Warps * kC32Stages * kC32Tile * sizeof(uint2));
  const unsigned tid = threadIdx.x, warp = tid >> 5, lane = tid & 31u;
  const size_t tiles = (pr.n + kC32Tile - 1) / kC32Tile;
  const size_t gw = (size_t)blockIdx.x * kC32Warps + warp;       // this warp's index in the grid
  const size_t W = (size_t)gridDim.x * kC32Warps;                // warps of the grid: tile of iteration i = i * W + gw
  uint2* const ring = ring_all + (size_t)warp * kC32Stages * kC32Tile;
  uint64_t* const bar = sm.bar[warp];
  uint64_t pol_stream, pol_table;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_stream));
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol_table));

  for (unsigned p = tid; p < g.nparts; p += kC32Threads) {
    sm.part_off[p] = t.offset[p];
    sm.part_mask[p] = t.mask[p];
  }
  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < kC32Stages; ++s) tma::mbar_init(&bar[s], 1);
    tma::fence_barrier_init();
  }
  __syncthreads();  // the only CTA-wide barrier: partition geometry + barriers are set up

  auto issue = [&](size_t it) {  // lane 0: start the copy of this warp's tile of iteration `it` (full tiles only)
    const size_t tl = it * W + gw;
    if (tl < tiles && (tl + 1) * kC32Tile <= pr.n) {
      const int s = (int)(it % kC32Stages);
      tma::mbar_expect_tx(&bar[s], kC32Tile * (uint32_t)sizeof(uint2));
      bulk_load_policy(ring + (size_t)s * kC32Tile, pr.pairs + tl * kC32Tile, kC32Tile * (uint32_t)sizeof(uint2), &bar[s],
                       pol_stream);
    }
  };
  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < kC32Stages; ++s) issue((size_t)s);
  }

  // per-warp output chunk state (MODE == P32_CHUNK); identical in all lanes except `next_base` (lane 0 holds it)
  unsigned long long chunk_base = 0, next_base = 0;
  unsigned chunk_used = kC32Chunk;  // "no chunk yet"
  bool have_next = false;
  unsigned long long counted = 0;   // P32_COUNT: matches seen by this lane

  for (size_t it = 0;; ++it) {
    const size_t tl = it * W + gw;
    if (tl >= tiles) break;
    const size_t row0 = tl * kC32Tile;
    const bool full = row0 + kC32Tile <= pr.n;
    const int s = (int)(it % kC32Stages);
    if (MODE == P32_CHUNK && !have_next && chunk_used + kC32Tile > kC32Chunk) {
      // the tile may overflow the current chunk: ask for the next one now, use it (if at all) after the look-ups
      if (lane == 0) next_base = atomicAdd(out.cursor, (unsigned long long)kC32Chunk);
      have_next = true;
    }
    uint32_t key[kC32Rows];
    int32_t tag[kC32Rows];
    if (full) {
      tma::mbar_wait(&bar[s], (unsigned)(it / kC32Stages) & 1u);
#pragma unroll
      for (int i = 0; i < kC32Rows; ++i) {
        const uint2 v = ring[(size_t)s * kC32Tile + i * 32 + lane];
        key[i] = v.x;
        tag[i] = (int32_t)v.y;
      }
      __syncwarp();  // every lane holds its pairs: the stage can be refilled
      if (lane == 0) {
        tma::fence_proxy_async();
        issue(it + kC32Stages);
      }
    } else {
#pragma unroll
      for (int i = 0; i < kC32Rows; ++i) {
        const size_t j = row0 + i * 32 + lane;
        uint2 v = make_uint2(0u, 0x80000000u);  // past the end: tag INT_MIN = "no row"
        if (j < pr.n) v = pr.pairs[j];
        key[i] = v.x;
        tag[i] = (int32_t)v.y;
      }
    }
    // ---- look-ups: every row's home bucket in flight, then examined ----
    int32_t first[kC32Rows];
    unsigned cnt[kC32Rows];
    unsigned where[kC32Rows];   // partition << 24 | bucket's first slot   (slots per partition <= 2^24)
    Bucket32 bk[kC32Rows];
    unsigned pend = 0;
#pragma unroll
    for (int i = 0; i < kC32Rows; ++i) {
      first[i] = -1;
      const bool have = full || (tag[i] != (int32_t)0x80000000);
      cnt[i] = (LEFT_LIKE && have) ? 1u : 0u;
      const uint32_t h = KeyBits<uint32_t>::hash(key[i]);
      const unsigned p = g.pid(h);
      where[i] = (p << 24) | ((slot_hash(h) & sm.part_mask[p]) & ~3u);
      bool look = have && tag[i] >= 0;  // negative tag: the row can never match (NULL key / wide key)
#ifdef B200_LAB
      if ((lab & 1u) && look) { look = false; cnt[i] = 1; first[i] = 0; }   // ablation: no table look-up
#endif
      if (look) {
        pend |= 1u << i;
        bk[i] = ld_bucket32_hint(t.slots + sm.part_off[p] + (where[i] & 0xffffffu), pol_table);
      }
    }
    unsigned matched = 0;  // bit i: row i has found at least one partner
    while (true) {
#pragma unroll
      for (int i = 0; i < kC32Rows; ++i) {
        if (!((pend >> i) & 1u)) continue;
        bool stop = false;
        unsigned c = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const unsigned long long w = bk[i].w[j];
          if (slot_empty(w)) stop = true;           // the chain of this key ends in this bucket
          else if ((uint32_t)w == key[i]) {
            if (first[i] < 0) first[i] = (int32_t)(uint32_t)(w >> 32);
            ++c;
          }
        }
        if (c) {
          if (LEFT_LIKE && !((matched >> i) & 1u)) cnt[i] = 0;  // real matches replace the provisional (row,-1) pair
          matched |= 1u << i;
          cnt[i] += c;
          if (UNIQUE) stop = true;
        }
        if (stop) pend &= ~(1u << i);
      }
      if (!__any_sync(0xffffffffu, pend != 0)) break;
#pragma unroll
      for (int i = 0; i < kC32Rows; ++i) {
        if (!((pend >> i) & 1u)) continue;
        const unsigned p = where[i] >> 24;
        const unsigned at = ((where[i] & 0xffffffu) + 4u) & sm.part_mask[p];
        where[i] = (p << 24) | at;
        bk[i] = ld_bucket32_hint(t.slots + sm.part_off[p] + at, pol_table);
      }
    }
    if (MODE == P32_COUNT) {
#pragma unroll
      for (int i = 0; i < kC32Rows; ++i) counted += cnt[i];
      continue;
    }
    if (MODE == P32_POSITIONAL) {  // LEFT_LIKE && UNIQUE: exactly one pair per row, at the row's own position
#pragma unroll
      for (int i = 0; i < kC32Rows; ++i) {
        const size_t j = row0 + i * 32 + lane;
        if (j < pr.n) {
          st_i32_stream(out.probe + j, tag[i] >= 0 ? tag[i] : ~tag[i], pol_stream);
          st_i32_stream(out.build + j, first[i], pol_stream);
        }
      }
      continue;
    }
    // ---- ranks inside the warp (row-major: row i of all lanes, then row i + 1) ----
    unsigned rank[kC32Rows], total = 0;
#pragma unroll
    for (int i = 0; i < kC32Rows; ++i) {
      if (UNIQUE) {  // counts are 0 / 1
        const unsigned bal = __ballot_sync(0xffffffffu, cnt[i] != 0);
        rank[i] = total + __popc(bal & lanemask_lt());
        total += __popc(bal);
      } else {
        unsigned inc = cnt[i];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          const unsigned o = __shfl_up_sync(0xffffffffu, inc, d);
          if (lane >= (unsigned)d) inc += o;
        }
        rank[i] = total + inc - cnt[i];
        total += __shfl_sync(0xffffffffu, inc, 31);
      }
    }
#ifdef B200_LAB
    if (lab & 2u) continue;   // ablation: no output stores
#endif
    if (MODE == P32_CURSOR) {
      unsigned long long base = 0;
      if (lane == 0 && total) base = atomicAdd(out.cursor, (unsigned long long)total);
      base = __shfl_sync(0xffffffffu, base, 0);
#pragma unroll
      for (int i = 0; i < kC32Rows; ++i) {
        size_t pos = (size_t)base + rank[i];
        const int32_t prow = tag[i] >= 0 ? tag[i] : ~tag[i];
        if (cnt[i] == 1) {
          out.probe[pos] = prow;
          out.build[pos] = first[i];
        } else if (cnt[i] > 1) {  // duplicate build keys: walk the chain again
          const uint32_t h = KeyBits<uint32_t>::hash(key[i]);
          const unsigned p = g.pid(h);
          const unsigned long long* tb = t.slots + sm.part_off[p];
          unsigned at = (slot_hash(h) & sm.part_mask[p]) & ~3u;
          bool stop = false;
          while (!stop) {
            const Bucket32 b = ld_bucket32_hint(tb + at, pol_table);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              if (slot_empty(b.w[j])) stop = true;
              else if ((uint32_t)b.w[j] == key[i]) {
                out.probe[pos] = prow;
                out.build[pos] = (int32_t)(uint32_t)(b.w[j] >> 32);
                ++pos;
              }
            }
            at = (at + 4u) & sm.part_mask[p];
          }
        }
      }
      continue;
    }
    // ---- MODE == P32_CHUNK (UNIQUE, INNER): dense fill of per-warp chunks ----
    unsigned in_old = total;                       // pairs of this tile that still fit the current chunk
    unsigned long long old_at = chunk_base + chunk_used, new_base = 0;
    if (chunk_used + total > kC32Chunk) {
      in_old = kC32Chunk - chunk_used;
      new_base = __shfl_sync(0xffffffffu, next_base, 0);   // requested at the top of an earlier or this iteration
      chunk_base = new_base;
      chunk_used = total - in_old;
      have_next = false;
    } else {
      chunk_used += total;
    }
#pragma unroll
    for (int i = 0; i < kC32Rows; ++i) {
      if (cnt[i] == 0) continue;
      const size_t pos = rank[i] < in_old ? (size_t)(old_at + rank[i]) : (size_t)(new_base + (rank[i] - in_old));
      st_i32_stream(out.probe + pos, tag[i], pol_stream);
      st_i32_stream(out.build + pos, first[i], pol_stream);
    }
  }
  if (MODE == P32_COUNT) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) counted += __shfl_xor_sync(0xffffffffu, counted, d);
    if (lane == 0 && counted) atomicAdd(out.cursor, counted);
  }
  if (MODE == P32_CHUNK && lane == 0) {  // what this warp reserved and did not fill
    out.hole_start[2 * gw] = chunk_base + chunk_used;
    out.hole_len[2 * gw] = kC32Chunk - chunk_used;
    out.hole_start[2 * gw + 1] = next_base;
    out.hole_len[2 * gw + 1] = have_next ? kC32Chunk : 0u;
  }
}

// ---- probe, UNIQUE build keys (the PK/FK case, C3): the lean kernel ----
// ncu on the first version of probe32_kernel at C3 (profiles/r02a_ncu_join_full.md): ~200 SASS instructions per row,
// 17 % of them branch / reconvergence (BSSY, BSYNC, BRA) from per-row `if`s, stall_wait dominant with 4 warps per
// scheduler - the kernel was bound by its own instruction stream.  With that fixed (8.7 ms) the ablations showed a
// latency chain instead: look-ups alone +3.1 ms (= the L1TEX rate of one divergent sector per cycle per SM), but
// look-ups AND stores +6.5 ms, because nearly every 256-row tile has a row whose home bucket is full of other keys
// (~2 % of the rows), and the whole warp waited a second dependent L2 round trip for it.  This version:
//   * a tile of 256 consecutive pairs lies inside ONE partition except at partition boundaries (the pairs are
//     partition-contiguous), so the table base and mask are warp-uniform values found from the tile's position
//     (part_start), not per-row shared-memory look-ups;
//   * the bucket test is branch-free: first = row word of the slot whose key word equals the key, scanning from the
//     last slot to the first (an EMPTY slot that happens to "match" key 0xffffffff yields row -1 = "no partner",
//     which is the right answer, because slots fill in order and nothing real can follow an EMPTY slot);
//     a row is settled when it found its partner or when the bucket's last slot is EMPTY;
//   * unsettled rows are NOT waited for: they go to the warp's STRAGGLER QUEUE in shared memory ({key, tag,
//     position, bucket number}), and up to 32 queued rows ride along with the NEXT tile as a ninth row - their next
//     bucket is fetched together with that tile's home buckets.  A tile therefore costs one L2 round trip.  (If a
//     tile ever produces more stragglers than the queue holds, that tile falls back to resolving in place.)
//   * INNER: output positions from per-warp chunks as described above, with a branch-free fast path for tiles that
//     do not cross a chunk boundary.  LEFT / FULL: one pair per row at the row's own position.
// PADDED layout (cap_tiles != 0; INNER joins with unique build keys): the probe side was partitioned WITHOUT a histogram
// pass - partition p owns the fixed region [p * cap, (p + 1) * cap) of pr.pairs (cap = cap_tiles * 256 pairs) and
// part_start[p] is the END of what the scatter wrote there.  A tile then never straddles partitions; tiles beyond a
// region's end are empty and skipped, the last tile of a region is partial.  Because tiles can now be skipped, the
// mbarrier phase of every ring stage is tracked explicitly (it flips only when a bulk copy was waited for).
constexpr unsigned kQCap = 128;   // straggler queue entries per warp (power of two)

struct Probe32USmem {
  uint64_t bar[kC32Warps][kC32Stages];
  unsigned long long part_start[kMaxParts + 1];   // first pair of partition p in pr.pairs; [nparts] = pr.n
  unsigned part_off[kMaxParts];                   // first slot of partition p (total slots < 2^32 on this path)
  unsigned part_mask[kMaxParts];                  // (slots_p - 1) & ~3: bucket-aligned slot mask
  uint4 queue[kC32Warps][kQCap];                  // .x key, .y tag, .z position in pr.pairs, .w bucket number to try next
};
constexpr size_t probe32u_smem_bytes() {
  return (size_t)kC32Warps * kC32Stages * kC32Tile * sizeof(uint2) + sizeof(Probe32USmem) + 128;
}

template <bool LEFT_LIKE, bool PADDED>
__global__ void __launch_bounds__(kC32Threads, 1)
probe32_unique_kernel(Pairs32 pr, PartGeom g, Tables32 t, Probe32Out out, const unsigned long long* __restrict__ part_start,
                      unsigned cap_tiles, unsigned lab) {
  extern __shared__ __align__(128) unsigned char probe32u_smem[];
  uint2* const ring_all = reinterpret_cast<uint2*>(probe32u_smem);
  Probe32USmem& sm = *reinterpret_cast<Probe32USmem*>(probe32u_smem + (size_t)kC32Warps * kC32Stages * kC32Tile * sizeof(uint2));
  const unsigned tid = threadIdx.x, warp = tid >> 5, lane = tid & 31u;
  // The kernel sits at the 128-register limit of a 512-thread CTA, and a spilled register costs far more than its
  // instructions here: local-memory traffic does not fit the L1 left beside 170 KB of shared memory, reaches L2 and
  // evicts the tables (ncu, first padded version: 24 bytes of spill -> +20 GB of L2 reads, table hit rate 83 % -> 52 %,
  // 8.4 -> 13.4 ms).  The padded variant therefore keeps tile indices and output positions in 32 bits (tiles < 2^32;
  // join outputs are int32-indexed, so positions < 2^32), the contiguous variant keeps the types it was tuned with.
  using idx_t = typename std::conditional<PADDED, unsigned, size_t>::type;
  using pos_t = typename std::conditional<PADDED, unsigned, unsigned long long>::type;
  const idx_t tiles = (idx_t)((pr.n + kC32Tile - 1) / kC32Tile);
  const idx_t gw = (idx_t)blockIdx.x * kC32Warps + warp;
  const idx_t W = (idx_t)gridDim.x * kC32Warps;
  uint2* const ring = ring_all + (size_t)warp * kC32Stages * kC32Tile;
  uint64_t* const bar = sm.bar[warp];
  uint4* const queue = sm.queue[warp];
  uint64_t pol_stream, pol_table, pol_store;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_stream));
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol_table));
  pol_store = pol_stream;
#ifdef B200_LAB   // L2 policy A/B (lab >> 8): bit 0 pairs normal, bit 1 tables normal, bit 2 stores normal, bit 3 stores evict_last
  {
    uint64_t pol_normal;
    asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol_normal));
    const unsigned hk = lab >> 8;
    if (hk & 1u) pol_stream = pol_normal;
    if (hk & 2u) pol_table = pol_normal;
    if (hk & 4u) pol_store = pol_normal;
    if (hk & 8u) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol_store));
  }
#endif
  for (unsigned p = tid; p < g.nparts; p += kC32Threads) {
    sm.part_off[p] = (unsigned)t.offset[p];
    sm.part_mask[p] = t.mask[p] & ~3u;
  }
  for (unsigned p = tid; p <= g.nparts; p += kC32Threads) sm.part_start[p] = part_start[p];
  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < kC32Stages; ++s) tma::mbar_init(&bar[s], 1);
    tma::fence_barrier_init();
  }
  __syncthreads();  // the only CTA-wide barrier

  // rows of tile tl that exist: contiguous layout - everything up to pr.n; padded layout - up to the end of what the
  // scatter wrote into the tile's partition region (32-bit arithmetic: tiles < 2^32)
  auto tile_rows = [&](idx_t tl) -> unsigned {
    if (tl >= tiles) return 0u;
    const size_t row0 = (size_t)tl * kC32Tile;
    const size_t end = PADDED ? (size_t)sm.part_start[(unsigned)tl / cap_tiles] : pr.n;
    return end <= row0 ? 0u : (end - row0 >= (size_t)kC32Tile ? (unsigned)kC32Tile : (unsigned)(end - row0));
  };
  auto issue = [&](idx_t it) {  // lane 0: bulk copy of this warp's tile of iteration `it` (full tiles only)
    const idx_t tl = it * W + gw;
    if (PADDED ? tile_rows(tl) == (unsigned)kC32Tile : (tl < tiles && ((size_t)tl + 1) * kC32Tile <= pr.n)) {
      const int s = (int)(it % kC32Stages);
      tma::mbar_expect_tx(&bar[s], kC32Tile * (uint32_t)sizeof(uint2));
      bulk_load_policy(ring + (size_t)s * kC32Tile, pr.pairs + (size_t)tl * kC32Tile, kC32Tile * (uint32_t)sizeof(uint2), &bar[s],
                       pol_stream);
    }
  };
  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < kC32Stages; ++s) issue((idx_t)s);
  }
  unsigned phase_bits = 0;  // bit s = parity the next bulk copy into stage s completes
  const unsigned lt_mask = lanemask_lt();
  pos_t chunk_base = 0, next_base = 0;
  unsigned chunk_used = kC32Chunk;  // "no chunk yet"
  bool have_next = false;
  unsigned cur_p = 0;               // partition of the current tile's first pair (tiles of a warp only move forward)
  unsigned q_head = 0, q_count = 0; // straggler queue (warp-uniform)
#ifdef B200_LAB
  const bool no_lookup = (lab & 1u) != 0;     // ablation: no table look-up
  const bool drop_stragglers = (lab & 4u) != 0;  // ablation: unsettled rows are forgotten
#else
  constexpr bool no_lookup = false, drop_stragglers = false;
#endif
  // absolute index of the first slot of bucket number `round` of key k (0 = home bucket), any partition
  auto bucket_at = [&](uint32_t k, unsigned round) -> unsigned {
    const uint32_t h = KeyBits<uint32_t>::hash(k);
    const unsigned p = h >> g.shift;
    return sm.part_off[p] + ((slot_hash(h) + 4u * round) & sm.part_mask[p]);
  };
  // first = row word of the slot holding k (or -1); settled = found, or the bucket's last slot is EMPTY
  auto examine = [&](const Bucket32& b, uint32_t k, int32_t& f) -> bool {
    f = -1;
#pragma unroll
    for (int j = 3; j >= 0; --j)
      if ((uint32_t)b.w[j] == k) f = (int32_t)(uint32_t)(b.w[j] >> 32);
    return f >= 0 || (uint32_t)(b.w[3] >> 32) == 0xffffffffu;
  };

  for (idx_t it = 0;; ++it) {
    const idx_t tl = it * W + gw;
    const unsigned rows_here = PADDED ? tile_rows(tl) : 0u;
    const bool has_tile = PADDED ? rows_here != 0 : tl < tiles;
    if (PADDED ? (tl >= tiles && q_count == 0) : (!has_tile && q_count == 0)) break;   // after the last tile: extra iterations drain the queue
    const size_t row0 = (size_t)tl * kC32Tile;
    const bool full = PADDED ? rows_here == (unsigned)kC32Tile : (has_tile && row0 + kC32Tile <= pr.n);
    const int s = (int)(it % kC32Stages);
    if (PADDED && !has_tile && q_count == 0) {  // an empty tile of the padded layout (warp-uniform): nothing was copied for it
      if (lane == 0) issue(it + kC32Stages);
      continue;
    }
    if (!LEFT_LIKE && !have_next && chunk_used + (kC32Tile + 32) > kC32Chunk) {
      // the tile (+ up to 32 queued rows) may overflow the current chunk: ask for the next one now, use it later
      if (lane == 0) next_base = (pos_t)atomicAdd(out.cursor, (unsigned long long)kC32Chunk);
      have_next = true;
    }
    bool one_part = false;
    unsigned off_u = 0, mask_u = 0;
    if (has_tile) {
      if (PADDED) {
        cur_p = (unsigned)tl / cap_tiles;
        one_part = true;
      } else {
        while (row0 >= sm.part_start[cur_p + 1]) ++cur_p;
        one_part = (full ? row0 + kC32Tile : pr.n) <= sm.part_start[cur_p + 1];
      }
      off_u = sm.part_off[cur_p];
      mask_u = sm.part_mask[cur_p];
    }
    // The tile's pairs stay in the ring stage for the whole iteration and are re-read (LDS.64) where needed: holding
    // 8 keys + 8 tags in registers across the bucket loads spilled (128-register budget at 512 threads).  The stage
    // is handed back to the producer at the end of the iteration; three other stages are in flight meanwhile.
    uint2* const stage = ring + (size_t)s * kC32Tile;
    if (full) {
      // contiguous layout: every earlier use of the stage was a bulk copy, so the phase follows from the iteration number
      tma::mbar_wait(&bar[s], PADDED ? ((phase_bits >> s) & 1u) : ((unsigned)(it / kC32Stages) & 1u));
      if (PADDED) phase_bits ^= 1u << s;
    } else {  // partial tile / queue-draining iteration: no bulk copy was issued for this stage
#pragma unroll
      for (int i = 0; i < kC32Rows; ++i) {
        const size_t j = row0 + i * 32 + lane;
        uint2 v = make_uint2(0u, 0x80000000u);  // no row here: tag INT_MIN
        if (PADDED ? ((unsigned)(i * 32 + lane) < rows_here) : (has_tile && j < pr.n)) v = pr.pairs[j];
        stage[i * 32 + lane] = v;
      }
      __syncwarp();
    }
    auto pair_at = [&](int i) -> uint2 { return stage[i * 32 + lane]; };
    // ---- all bucket loads of the iteration in flight: the tile's home buckets + one queued straggler per lane ----
    Bucket32 bk[kC32Rows], qbk;
    if (one_part) {  // warp-uniform: the common case gets its own straight-line code
#pragma unroll
      for (int i = 0; i < kC32Rows; ++i) {
        const uint2 v = pair_at(i);
        if ((int32_t)v.y >= 0 && !no_lookup)
          bk[i] = ld_bucket32_hint(t.slots + (off_u + (slot_hash(KeyBits<uint32_t>::hash(v.x)) & mask_u)), pol_table);
      }
    } else {
#pragma unroll
      for (int i = 0; i < kC32Rows; ++i) {
        const uint2 v = pair_at(i);
        if ((int32_t)v.y >= 0 && !no_lookup) bk[i] = ld_bucket32_hint(t.slots + bucket_at(v.x, 0), pol_table);
      }
    }
    const unsigned take = q_count < 32u ? q_count : 32u;
    const bool qlive = lane < take;
    uint4 qe = make_uint4(0u, 0u, 0u, 0u);
    if (qlive) {
      qe = queue[(q_head + lane) & (kQCap - 1)];
      qbk = ld_bucket32_hint(t.slots + bucket_at(qe.x, qe.w), pol_table);
    }
    q_head += take;
    q_count -= take;
    // ---- examine ----
    int32_t first[kC32Rows], qfirst = -1;
    unsigned pend = 0;   // bit i: row i is unsettled (goes to the queue)
    unsigned live = 0;   // bit i: row i exists (tag != INT_MIN)
#pragma unroll
    for (int i = 0; i < kC32Rows; ++i) {
      const uint2 v = pair_at(i);
      int32_t f;
      const bool settled = examine(bk[i], v.x, f);
      const bool look = (int32_t)v.y >= 0 && !no_lookup;
      first[i] = look ? f : (no_lookup && (int32_t)v.y >= 0 ? 0 : -1);
      if (look && !settled) pend |= 1u << i;
      if (v.y != 0x80000000u) live |= 1u << i;
    }
    bool qpend = false;
    if (qlive) qpend = !examine(qbk, qe.x, qfirst);
    if (drop_stragglers) { pend = 0; qpend = false; }
    // ---- unsettled rows go (back) to the queue; if they do not fit, this tile's rows are resolved in place ----
    {
      unsigned need = __popc(__ballot_sync(0xffffffffu, qpend));
#pragma unroll
      for (int i = 0; i < kC32Rows; ++i) need += __popc(__ballot_sync(0xffffffffu, (pend >> i) & 1u));
      if (need) {
        if (q_count + need > kQCap) {  // does not fit (pathological collisions): blocking rounds for the tile's rows
          for (unsigned round = 1; __any_sync(0xffffffffu, pend != 0); ++round) {
#pragma unroll
            for (int i = 0; i < kC32Rows; ++i)
              if ((pend >> i) & 1u) bk[i] = ld_bucket32_hint(t.slots + bucket_at(pair_at(i).x, round), pol_table);
#pragma unroll
            for (int i = 0; i < kC32Rows; ++i) {
              if (!((pend >> i) & 1u)) continue;
              if (examine(bk[i], pair_at(i).x, first[i])) pend &= ~(1u << i);
            }
          }
        }
        unsigned tail = q_head + q_count;
        {  // queued rows that are still unsettled: next bucket
          const unsigned b = __ballot_sync(0xffffffffu, qpend);
          if (qpend) queue[(tail + __popc(b & lt_mask)) & (kQCap - 1)] = make_uint4(qe.x, qe.y, qe.z, qe.w + 1u);
          tail += __popc(b);
        }
#pragma unroll
        for (int i = 0; i < kC32Rows; ++i) {
          const unsigned b = __ballot_sync(0xffffffffu, (pend >> i) & 1u);
          if (!b) continue;  // warp-uniform
          if ((pend >> i) & 1u) {
            const uint2 v = pair_at(i);
            queue[(tail + __popc(b & lt_mask)) & (kQCap - 1)] = make_uint4(v.x, v.y, (uint32_t)(row0 + i * 32 + lane), 1u);
          }
          tail += __popc(b);
        }
        q_count = tail - q_head;
        __syncwarp();
      }
    }
    const bool qdone = qlive && !qpend;   // a queued row that is settled now (found or proven absent)
    if (LEFT_LIKE) {  // exactly one pair per row, at the row's own position; queued rows are written when settled
#pragma unroll
      for (int i = 0; i < kC32Rows; ++i) {
        if (((live >> i) & 1u) && !((pend >> i) & 1u)) {
          const size_t j = row0 + i * 32 + lane;
          const int32_t tg = (int32_t)pair_at(i).y;
          st_i32_stream(out.probe + j, tg >= 0 ? tg : ~tg, pol_store);
          st_i32_stream(out.build + j, first[i], pol_store);
        }
      }
      if (qdone) {
        st_i32_stream(out.probe + qe.z, (int32_t)qe.y, pol_store);
        st_i32_stream(out.build + qe.z, qfirst, pol_store);
      }
    } else {
      // ---- INNER: ranks from ballots, positions from the warp's chunk ----
      unsigned rank[kC32Rows], qrank, total = 0;
#pragma unroll
      for (int i = 0; i < kC32Rows; ++i) {
        if ((pend >> i) & 1u) first[i] = -1;   // queued: emitted later
        const unsigned b = __ballot_sync(0xffffffffu, first[i] >= 0);
        rank[i] = total + __popc(b & lt_mask);
        total += __popc(b);
      }
      {
        const unsigned b = __ballot_sync(0xffffffffu, qdone && qfirst >= 0);
        qrank = total + __popc(b & lt_mask);
        total += __popc(b);
      }
#ifdef B200_LAB
      if (lab & 2u) total = 0;   // ablation: no output stores
#endif
      if (total == 0) {
      } else if (chunk_used + total <= kC32Chunk) {  // the iteration's pairs fit the current chunk
        int32_t* const po = out.probe + (chunk_base + chunk_used);
        int32_t* const bo = out.build + (chunk_base + chunk_used);
        chunk_used += total;
#pragma unroll
        for (int i = 0; i < kC32Rows; ++i) {
          if (first[i] >= 0) {
            st_i32_stream(po + rank[i], (int32_t)pair_at(i).y, pol_store);
            st_i32_stream(bo + rank[i], first[i], pol_store);
          }
        }
        if (qdone && qfirst >= 0) {
          st_i32_stream(po + qrank, (int32_t)qe.y, pol_store);
          st_i32_stream(bo + qrank, qfirst, pol_store);
        }
      } else {  // split: the first in_old pairs finish the current chunk, the rest start the next one
        const unsigned in_old = kC32Chunk - chunk_used;
        const pos_t old_at = chunk_base + chunk_used;
        const pos_t new_base = __shfl_sync(0xffffffffu, next_base, 0);
        chunk_base = new_base;
        chunk_used = total - in_old;
        have_next = false;
        auto place = [&](unsigned r) -> size_t { return r < in_old ? (size_t)(old_at + r) : (size_t)(new_base + (r - in_old)); };
#pragma unroll
        for (int i = 0; i < kC32Rows; ++i) {
          if (first[i] < 0) continue;
          const size_t pos = place(rank[i]);
          st_i32_stream(out.probe + pos, (int32_t)pair_at(i).y, pol_store);
          st_i32_stream(out.build + pos, first[i], pol_store);
        }
        if (qdone && qfirst >= 0) {
          const size_t pos = place(qrank);
          st_i32_stream(out.probe + pos, (int32_t)qe.y, pol_store);
          st_i32_stream(out.build + pos, qfirst, pol_store);
        }
      }
    }
    // ---- every lane is done with the stage: hand it back to the producer ----
    __syncwarp();
    if ((PADDED || full) && lane == 0) {  // padded: also after a partial tile - its stage was written with ordinary stores
      tma::fence_proxy_async();
      issue(it + kC32Stages);
    }
  }
  if (!LEFT_LIKE && lane == 0) {  // what this warp reserved and did not fill
    out.hole_start[2 * gw] = chunk_base + chunk_used;
    out.hole_len[2 * gw] = kC32Chunk - chunk_used;
    out.hole_start[2 * gw + 1] = next_base;
    out.hole_len[2 * gw + 1] = have_next ? kC32Chunk : 0u;
  }
}

// ---- fix-up of the chunked output: move the tail into the holes ----
// plan kernel (one CTA): sorts the holes by position, computes found = allocated - sum(hole lengths), and two
// segment lists with exclusive prefix sums: RECEIVERS = hole positions below `found`, DONORS = filled positions at or
// above `found`.  Both lists describe the same number of pairs (moved).  move kernel: pair m of the donors goes to
// position m of the receivers.
constexpr int kFixThreads = 1024;
constexpr int kFixCap = 8192;   // >= 2 * warps of the probe grid (148 SMs x 16 warps x 2 = 4736)
constexpr int kFixWindow = 2 * kFixCap;   // chunks covered by the direct ordering of the holes (fixup_plan_kernel)

struct FixPlan {          // device memory, filled by fixup_plan_kernel
  unsigned long long found, moved;
  unsigned n_recv, n_donor;
  unsigned long long recv_start[kFixCap + 1], donor_start[kFixCap + 1];
  unsigned recv_pre[kFixCap + 2], donor_pre[kFixCap + 2];   // exclusive prefix sums of the segment lengths (+ total)
};

// exclusive scan of v[0..kFixCap) in place (shared memory), returns the total; all kFixThreads threads call it
static __device__ unsigned fix_scan(unsigned* v, unsigned* warp_sums) {
  constexpr int PER = kFixCap / kFixThreads;
  const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  unsigned loc[PER], sum = 0;
#pragma unroll
  for (int j = 0; j < PER; ++j) {
    loc[j] = v[tid * PER + j];
    sum += loc[j];
  }
  unsigned inc = sum;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const unsigned o = __shfl_up_sync(0xffffffffu, inc, d);
    if (lane >= (unsigned)d) inc += o;
  }
  if (lane == 31) warp_sums[warp] = inc;
  __syncthreads();
  unsigned off = 0, tot = 0;
  for (int w = 0; w < kFixThreads / 32; ++w) {
    const unsigned ws = warp_sums[w];
    if ((unsigned)w < warp) off += ws;
    tot += ws;
  }
  unsigned run = off + inc - sum;
#pragma unroll
  for (int j = 0; j < PER; ++j) {
    v[tid * PER + j] = run;
    run += loc[j];
  }
  __syncthreads();
  return tot;
}

__global__ void __launch_bounds__(kFixThreads)
fixup_plan_kernel(const unsigned long long* __restrict__ hole_start, const unsigned* __restrict__ hole_len, unsigned n_holes,
                  const unsigned long long* __restrict__ cursor, FixPlan* __restrict__ plan) {
  extern __shared__ __align__(16) unsigned char fix_smem[];
  unsigned long long* hs = reinterpret_cast<unsigned long long*>(fix_smem);          // [kFixCap] start << 24 | idx? no: start
  unsigned* hl = reinterpret_cast<unsigned*>(hs + kFixCap);                          // [kFixCap] length
  unsigned* len_a = hl + kFixCap;                                                    // [kFixCap] receiver lengths
  unsigned* len_b = len_a + kFixCap;                                                 // [kFixCap] donor lengths
  __shared__ unsigned warp_sums[kFixThreads / 32];
  const unsigned tid = threadIdx.x;
  const unsigned long long allocated = *cursor;
  // sort key: start (40 bits are plenty) << 24 | length (<= 2048); empty entries sort last
  for (unsigned i = tid; i < kFixCap; i += kFixThreads) {
    unsigned long long k = ~0ull;
    if (i < n_holes && hole_len[i] != 0) k = (hole_start[i] << 24) | hole_len[i];
    hs[i] = k;
  }
  __syncthreads();
  // Ordering the holes.  Every hole is the unfilled tail of ONE output chunk (or a whole reserved chunk), so its chunk
  // number orders it, and the chunks with holes are the last ones every warp took - a window of a few thousand chunks
  // at the end of the allocation.  When that window fits kFixWindow entries the holes are ordered by direct placement
  // + one scan (~10 us); the 91-stage bitonic sort (0.3 ms, 7 % of the step at 8 GPUs) is the fallback.
  __shared__ unsigned long long s_cmin;
  __shared__ unsigned s_count;
  if (tid == 0) s_cmin = ~0ull;
  __syncthreads();
  {
    unsigned long long mine = ~0ull;
    for (unsigned i = tid; i < kFixCap; i += kFixThreads)
      if (hs[i] != ~0ull) {
        const unsigned long long c = (hs[i] >> 24) / kC32Chunk;
        mine = c < mine ? c : mine;
      }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      const unsigned long long o = __shfl_xor_sync(0xffffffffu, mine, d);
      mine = o < mine ? o : mine;
    }
    if ((tid & 31u) == 0 && mine != ~0ull) atomicMin(&s_cmin, mine);
  }
  __syncthreads();
  const unsigned long long cmin = s_cmin;
  const bool direct = cmin != ~0ull && allocated / kC32Chunk - cmin <= (unsigned long long)kFixWindow;
  if (cmin == ~0ull) {
    // no holes at all: hs is already "sorted" (all entries empty)
  } else if (direct) {
    unsigned* win = len_a;  // kFixWindow entries = len_a + len_b (both are written only later)
    for (unsigned i = tid; i < kFixWindow; i += kFixThreads) win[i] = 0;
    __syncthreads();
    for (unsigned i = tid; i < kFixCap; i += kFixThreads)
      if (hs[i] != ~0ull) win[(hs[i] >> 24) / kC32Chunk - cmin] = (unsigned)(hs[i] & 0xffffffu);
    __syncthreads();
    constexpr int PER = kFixWindow / kFixThreads;
    unsigned cnt = 0;
#pragma unroll
    for (int j = 0; j < PER; ++j) cnt += win[tid * PER + j] != 0;
    unsigned inc = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const unsigned o = __shfl_up_sync(0xffffffffu, inc, d);
      if ((tid & 31u) >= (unsigned)d) inc += o;
    }
    if ((tid & 31u) == 31u) warp_sums[tid >> 5] = inc;
    __syncthreads();
    unsigned off = 0, tot = 0;
    for (int w = 0; w < kFixThreads / 32; ++w) {
      const unsigned ws = warp_sums[w];
      if ((unsigned)w < (tid >> 5)) off += ws;
      tot += ws;
    }
    unsigned at = off + inc - cnt;
#pragma unroll
    for (int j = 0; j < PER; ++j) {
      const unsigned len = win[tid * PER + j];
      if (len) {  // a hole is the TAIL of its chunk: it ends where the chunk ends
        const unsigned long long start = (cmin + tid * PER + j + 1) * kC32Chunk - len;
        hs[at++] = (start << 24) | len;
      }
    }
    if (tid == 0) s_count = tot;
    __syncthreads();
    for (unsigned i = s_count + tid; i < kFixCap; i += kFixThreads) hs[i] = ~0ull;
    __syncthreads();
  } else {
  for (unsigned size = 2; size <= kFixCap; size <<= 1) {       // bitonic sort, ascending
    for (unsigned stride = size >> 1; stride > 0; stride >>= 1) {
      for (unsigned i = tid; i < kFixCap / 2; i += kFixThreads) {
        const unsigned lo = 2 * i - (i & (stride - 1));
        const unsigned hi = lo + stride;
        const bool up = (lo & size) == 0;
        const unsigned long long a = hs[lo], b = hs[hi];
        if ((a > b) == up) {
          hs[lo] = b;
          hs[hi] = a;
        }
      }
      __syncthreads();
    }
  }
  }
  __shared__ unsigned long long s_last_end;
  if (tid == 0) s_last_end = 0;
  __syncthreads();
  for (unsigned i = tid; i < kFixCap; i += kFixThreads) {
    const unsigned long long k = hs[i];
    hl[i] = k == ~0ull ? 0u : (unsigned)(k & 0xffffffu);
    // the entries are ordered with the empty ones last: exactly one real entry is followed by an empty one (or by the end)
    if (k != ~0ull && (i + 1 == kFixCap || hs[i + 1] == ~0ull)) s_last_end = (k >> 24) + (k & 0xffffffu);
  }
  __syncthreads();
  for (unsigned i = tid; i < kFixCap; i += kFixThreads) len_a[i] = hl[i];
  __syncthreads();
  const unsigned long long holes_total = fix_scan(len_a, warp_sums);
  const unsigned long long found = allocated - holes_total;
  // receivers: hole i clipped to [0, found).  donors: the filled gap BEFORE hole i clipped to [found, allocated),
  // entry kFixCap - 1 doubles as "the gap after the last hole" (there are at most 4736 real holes)
  for (unsigned i = tid; i < kFixCap; i += kFixThreads) {
    const unsigned long long k = hs[i];
    unsigned ra = 0, db = 0;
    if (k != ~0ull) {
      const unsigned long long st = k >> 24, en = st + hl[i];
      if (st < found) ra = (unsigned)((en < found ? en : found) - st);
      unsigned long long prev_end = 0;
      if (i > 0) prev_end = (hs[i - 1] >> 24) + hl[i - 1];
      const unsigned long long gs = prev_end > found ? prev_end : found;
      if (st > gs) db = (unsigned)(st - gs);
    } else if (i == kFixCap - 1) {
      const unsigned long long last_end = s_last_end;  // end of the last real hole (0 if there is none)
      const unsigned long long gs = last_end > found ? last_end : found;
      if (allocated > gs) db = (unsigned)(allocated - gs);
    }
    len_a[i] = ra;
    len_b[i] = db;
  }
  __syncthreads();
  // segment start positions (before the scans overwrite the lengths)
  for (unsigned i = tid; i < kFixCap; i += kFixThreads) {
    const unsigned long long k = hs[i];
    unsigned long long rs = 0, ds = 0;
    if (k != ~0ull) {
      rs = k >> 24;
      unsigned long long prev_end = 0;
      if (i > 0) prev_end = (hs[i - 1] >> 24) + hl[i - 1];
      ds = prev_end > found ? prev_end : found;
    } else if (i == kFixCap - 1) {
      const unsigned long long last_end = s_last_end;
      ds = last_end > found ? last_end : found;
    }
    plan->recv_start[i] = rs;
    plan->donor_start[i] = ds;
  }
  const unsigned moved_a = fix_scan(len_a, warp_sums);
  const unsigned moved_b = fix_scan(len_b, warp_sums);
  for (unsigned i = tid; i < kFixCap; i += kFixThreads) {
    plan->recv_pre[i] = len_a[i];
    plan->donor_pre[i] = len_b[i];
  }
  if (tid == 0) {
    plan->recv_pre[kFixCap] = moved_a;
    plan->donor_pre[kFixCap] = moved_b;
    plan->found = found;
    plan->moved = moved_a < moved_b ? moved_a : moved_b;  // equal by construction
    plan->n_recv = kFixCap;
    plan->n_donor = kFixCap;
  }
}

// last index i in [0, n) with pre[i] <= m  (pre is an exclusive prefix sum, non-decreasing; zero-length segments
// share a prefix with their successor, so "last" lands on the segment that really contains m)
static __device__ __forceinline__ unsigned seg_of(const unsigned* __restrict__ pre, unsigned n, unsigned m) {
  unsigned lo = 0, hi = n;  // invariant: pre[lo] <= m, pre[hi] > m  (pre[n] = total > m)
  while (hi - lo > 1) {
    const unsigned mid = (lo + hi) >> 1;
    if (pre[mid] <= m) lo = mid;
    else hi = mid;
  }
  return lo;
}

__global__ void __launch_bounds__(256)
fixup_move_kernel(const FixPlan* __restrict__ plan, int32_t* __restrict__ a, int32_t* __restrict__ b) {
  const unsigned moved = (unsigned)plan->moved;
  const unsigned stride = gridDim.x * blockDim.x;
  for (unsigned m = blockIdx.x * blockDim.x + threadIdx.x; m < moved; m += stride) {
    const unsigned r = seg_of(plan->recv_pre, kFixCap, m), d = seg_of(plan->donor_pre, kFixCap, m);
    const size_t to = (size_t)(plan->recv_start[r] + (m - plan->recv_pre[r]));
    const size_t from = (size_t)(plan->donor_start[d] + (m - plan->donor_pre[d]));
    a[to] = a[from];
    b[to] = b[from];
  }
}
