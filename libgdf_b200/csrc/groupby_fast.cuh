// Single-key group-by build kernel (v5).  Included by groupby.cu inside its anonymous namespace,
// after fold() / smem_add64() / kEmptyKey are defined.
//
// One CTA of 1024 threads per SM owns an 8192-slot, 2-way set-associative cache of {key, accumulator}
// in shared memory (a set's two keys are ONE 128-bit shared-memory load).  Under Zipf skew the cache absorbs
// ~65 % of the rows with shared-memory atomics; rows whose key is not cached fold into ONE L2-resident global
// table with red.global.  The input stream is staged one step ahead with cp.async (L2 evict_first).
//
// History.  v3 -> v4 (profiles/r01a, r01b): the instruction count per row was cut from 233 to 105 and the kernel
// went from 10.1 to 7.1 ms at C4.  What bounded v4 (profiles/r02_microbench.txt, r02_notes.md): a row that missed
// the cache walked the global table slot by slot ({key,acc} in 16 bytes, linear probing at load 0.24), one row
// after the other.  14 % of those walks need a second DEPENDENT L2 round trip, and with ~11 missing lanes per
// warp-row nearly every one of the four rows of a step paid it: +2.2 us per 128-row step, although the same
// B200 retires 1.9e11 spread-address red.add.u64 per second when nothing depends on a load.
// v5 removes the dependent walks:
//   * the global table is split into a key array and an accumulator array and is BUCKETISED: the home of a key
//     is an aligned group of BK keys (32 bytes = one L2 sector for BK = 4) fetched with one vector load; at load
//     0.24 a key lives outside its home bucket with probability 0.3 %, so the common miss is exactly one L2 round
//     trip + one fire-and-forget red, for all rows of the step at once;
//   * the step is reordered: cache sets of all four rows are examined first, the bucket loads of the misses are
//     issued (two rows at a time: 1024 threads leave 64 registers each, and four 32-byte buckets spilled), the
//     shared-memory atomics of the hits run while those loads are in flight, then the misses are resolved;
//   * everything else (claim of an empty slot, bucket overflow, the side slot for a key equal to the EMPTY
//     pattern) is one out-of-line slow path.
#pragma once

constexpr int kFastThreads4 = 1024;
constexpr unsigned kCacheSets4 = 4096;                 // x 2 ways
constexpr unsigned kCacheSlots4 = 2 * kCacheSets4;
constexpr unsigned kGrabRows4 = 32 * 4 * 32;           // rows per warp per work grab (4096)
constexpr unsigned kBucketKeys = 4;                    // keys per global bucket: 4 x 8 B = one 32-byte sector
constexpr int kMissBatch = 2;                          // rows whose bucket loads are in flight together (register budget: 64)

struct FastCache4 {
  unsigned long long key[kCacheSlots4];   // way pairs are adjacent: one LDS.128 per set
  unsigned long long acc[kCacheSlots4];
  // one-step-ahead staging of the input stream, 2 KB per warp: while a warp works on a step, the next
  // step's 128 keys + 128 values are already on their way (cp.async)
  uint4 stage[kFastThreads4 / 32][128];
};
// The input stream is read once: L2 evict_first, so that it does not push the 64 MB global table out of L2
// (ncu at full size before the hint: 21.0 GB of DRAM traffic for 16.0 GB of input).
static __device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src, uint64_t policy) {
  asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" ::"r"(
                   (uint32_t)__cvta_generic_to_shared(smem_dst)),
               "l"(gmem_src), "l"(policy)
               : "memory");
}

// Internal position hash.  h's top bits pick the cache set, its low bits the global bucket.
static __device__ __forceinline__ uint32_t mix_key(unsigned long long k) {
  uint32_t x = (uint32_t)k * 0x9E3779B1u ^ (uint32_t)(k >> 32) * 0x85EBCA77u;
  x ^= x >> 15;
  x *= 0x2C1B3C6Du;
  x ^= x >> 13;
  return x;
}

template <int FOLD>
static __device__ __forceinline__ void cache_fold(unsigned long long* acc, int64_t v) {
  if (FOLD == OP_SUM) smem_add64(acc, v);
  else if (FOLD == OP_MIN) atomicMin(reinterpret_cast<long long*>(acc), (long long)v);
  else atomicMax(reinterpret_cast<long long*>(acc), (long long)v);
}

// The L2-resident global table: keys and accumulators in separate arrays, `slots` a power of two (a multiple of
// kBucketKeys), plus one side slot at index `slots` for a real key equal to the EMPTY pattern.
struct GlobalTable {
  unsigned long long* keys;
  int64_t* acc;
  unsigned long long* cnt;   // per-slot row counts, AVG only
  unsigned mask;             // slots - 1
  unsigned probe_buckets;    // give up (overflow flag) after this many buckets
};

struct Bucket4 {
  unsigned long long k[kBucketKeys];
};
static __device__ __forceinline__ Bucket4 ld_bucket(const unsigned long long* p) {  // 32-byte aligned, L2 only
  Bucket4 b;
  asm volatile("ld.global.cg.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(b.k[0]), "=l"(b.k[1]), "=l"(b.k[2]), "=l"(b.k[3]) : "l"(p));
  return b;
}

// Everything that is not "the key sits in its home bucket": claim an EMPTY slot (first occurrence of a key),
// walk on when the bucket is full of other keys.  Restarts at the home bucket.  Out of line on purpose: it keeps
// the registers of the hot loop free (v3 inlined a loop like this one and spilled).
template <int FOLD, bool WITH_CNT>
static __device__ __noinline__ bool global_fold_slow(GlobalTable t, unsigned bucket, unsigned long long key, int64_t v,
                                                     unsigned long long c) {
  for (unsigned probe = 0; probe < t.probe_buckets; ++probe) {
#pragma unroll 1
    for (unsigned j = 0; j < kBucketKeys; ++j) {
      const unsigned s = bucket + j;
      unsigned long long k0 = *reinterpret_cast<volatile unsigned long long*>(&t.keys[s]);
      if (k0 == kEmptyKey) {
        const unsigned long long prev = atomicCAS(&t.keys[s], kEmptyKey, key);
        k0 = (prev == kEmptyKey) ? key : prev;
      }
      if (k0 == key) {
        fold(&t.acc[s], v, FOLD);
        if (WITH_CNT) atomicAdd(&t.cnt[s], c);
        return true;
      }
    }
    bucket = (bucket + kBucketKeys) & t.mask;
  }
  return false;
}

template <typename KT, typename IT, int FOLD, bool COUNT_ROWS, bool WITH_CNT>
__global__ void __launch_bounds__(kFastThreads4, 1)
build_fast_kernel_v5(const KT* __restrict__ key_col, const IT* __restrict__ values, size_t n, GlobalTable tab,
                     int* __restrict__ flags /*[0]=side slot used, [1]=overflow*/,
                     unsigned long long* __restrict__ work_counter, unsigned lab) {
  using UK = typename std::conditional<sizeof(KT) == 8, unsigned long long, unsigned>::type;
  constexpr bool VEC = sizeof(KT) == 8 && sizeof(IT) == 8;  // two rows per 128-bit load
  extern __shared__ __align__(16) unsigned char smem_raw[];
  FastCache4& cache = *reinterpret_cast<FastCache4*>(smem_raw);
  unsigned* ccnt = reinterpret_cast<unsigned*>(smem_raw + sizeof(FastCache4));  // only when WITH_CNT
  const int64_t identity = FOLD == OP_MIN ? (int64_t)std::numeric_limits<IT>::max()
                                          : (FOLD == OP_MAX ? (int64_t)std::numeric_limits<IT>::lowest() : 0);
  for (unsigned i = threadIdx.x; i < kCacheSlots4; i += kFastThreads4) {
    cache.key[i] = kEmptyKey;
    cache.acc[i] = (unsigned long long)identity;
    if (WITH_CNT) ccnt[i] = 0;
  }
  __syncthreads();
  const unsigned lane = lane_id();
  const bool vec_ok = VEC && aligned16(key_col) && aligned16(values);
  uint4* const stage = cache.stage[threadIdx.x >> 5];
  size_t staged_step = ~(size_t)0;  // row index of the step whose data is in (or on its way to) `stage`
  // lane's 16-byte chunks of a full step: keys of rows 2*lane(+1), 64+2*lane(+1), then the same for values
  uint64_t pol_stream;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_stream));
  auto prefetch_step = [&](size_t st) {
    cp_async16(&stage[lane], reinterpret_cast<const unsigned long long*>(key_col) + st + 2 * lane, pol_stream);
    cp_async16(&stage[32 + lane], reinterpret_cast<const unsigned long long*>(key_col) + st + 64 + 2 * lane, pol_stream);
    if (!COUNT_ROWS) {
      cp_async16(&stage[64 + lane], reinterpret_cast<const unsigned long long*>(values) + st + 2 * lane, pol_stream);
      cp_async16(&stage[96 + lane], reinterpret_cast<const unsigned long long*>(values) + st + 64 + 2 * lane, pol_stream);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    staged_step = st;
  };

  // Cache set of row (k, h): returns the slot index (set*2 + way) holding k, or -1.  A set with a free way is
  // claimed on the spot (cold start only; slots are never freed, so a key that was placed is always found again).
  auto cache_slot = [&](unsigned long long k, uint32_t h) -> int {
    const unsigned set2 = (h >> 20) * 2u;  // top 12 bits
    const ulonglong2 kk = *reinterpret_cast<const ulonglong2*>(&cache.key[set2]);
    if (kk.x == k) return (int)set2;
    if (kk.y == k) return (int)set2 + 1;
    if (kk.x == kEmptyKey) {
      const unsigned long long prev = atomicCAS(&cache.key[set2], kEmptyKey, k);
      if (prev == kEmptyKey || prev == k) return (int)set2;
    }
    if (kk.x == kEmptyKey || kk.y == kEmptyKey) {
      const unsigned long long prev = atomicCAS(&cache.key[set2 + 1], kEmptyKey, k);
      if (prev == kEmptyKey || prev == k) return (int)set2 + 1;
    }
    return -1;
  };

  while (true) {
    unsigned long long grab = 0;
    if (lane == 0) grab = atomicAdd(work_counter, 1ull);
    grab = __shfl_sync(0xffffffffu, grab, 0);
    const size_t base = (size_t)grab * kGrabRows4;
    if (base >= n) break;
    const size_t end = base + kGrabRows4 < n ? base + kGrabRows4 : n;
#pragma unroll 1
    for (size_t step = base; step < end; step += 128) {
      unsigned long long k[4];
      int64_t v[4];
      bool live[4];
      if (vec_ok && step + 128 <= end) {  // lane reads rows step + 2*lane (+1) and step + 64 + 2*lane (+1)
        if (staged_step != step) prefetch_step(step);  // first step of a grab: nothing was staged
        asm volatile("cp.async.wait_group 0;" ::: "memory");
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const uint4 kr = stage[q * 32 + lane];
          k[2 * q] = ((unsigned long long)kr.y << 32) | kr.x;
          k[2 * q + 1] = ((unsigned long long)kr.w << 32) | kr.z;
          if (!COUNT_ROWS) {
            const uint4 vr = stage[64 + q * 32 + lane];
            v[2 * q] = (int64_t)(((unsigned long long)vr.y << 32) | vr.x);
            v[2 * q + 1] = (int64_t)(((unsigned long long)vr.w << 32) | vr.z);
          } else {
            v[2 * q] = v[2 * q + 1] = 1;
          }
          live[2 * q] = live[2 * q + 1] = true;
        }
        // every lane reads back exactly the chunks it copied itself, so the buffer can be refilled at once
        if (step + 256 <= end) prefetch_step(step + 128);
      } else {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const size_t r = step + (size_t)u * 32 + lane;
          live[u] = r < end;
          k[u] = live[u] ? (unsigned long long)(UK)key_col[r] : 0ull;
          v[u] = (live[u] && !COUNT_ROWS) ? (int64_t)values[r] : 1;
        }
      }
      // (1) cache sets of all four rows.  loc[u] >= 0: cache slot holding the key; loc[u] == kDead: nothing left
      // to do for this row; otherwise ~loc[u] is the first slot of the key's home bucket in the global table.
      constexpr int kDead = INT_MIN;
      int loc[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        loc[u] = kDead;
        if (!live[u]) continue;
        if (k[u] == kEmptyKey) {  // a real key equal to the EMPTY pattern lives in the side slot
          fold(&tab.acc[tab.mask + 1], v[u], FOLD);
          if (WITH_CNT) atomicAdd(&tab.cnt[tab.mask + 1], 1ull);
          flags[0] = 1;
          continue;
        }
        const uint32_t h = mix_key(k[u]);
        const int s = cache_slot(k[u], h);
        loc[u] = s >= 0 ? s : ~(int)((h & tab.mask) & ~(kBucketKeys - 1));
      }
#ifdef B200_LAB   // timing ablation of profiles/r02_notes.md (wrong results by design): cache phase only
      if (lab & 4u) {
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (loc[u] >= 0) cache_fold<FOLD>(&cache.acc[loc[u]], v[u]);
        continue;
      }
#endif
      // (2) misses are resolved kMissBatch rows at a time: home buckets fetched together, then examined - the key
      // sits in its home bucket (one red) or the slow path takes over.  (3) The shared-memory folds of ALL hits are
      // issued right after the first batch of loads, so they run while those loads are in flight.
#pragma unroll
      for (int u0 = 0; u0 < 4; u0 += kMissBatch) {
        Bucket4 bk[kMissBatch];
#pragma unroll
        for (int q = 0; q < kMissBatch; ++q)
          if (loc[u0 + q] < 0 && loc[u0 + q] != kDead) bk[q] = ld_bucket(tab.keys + (unsigned)~loc[u0 + q]);
        if (u0 == 0) {
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            if (loc[u] < 0) continue;
            cache_fold<FOLD>(&cache.acc[loc[u]], v[u]);
            if (WITH_CNT) atomicAdd(&ccnt[loc[u]], 1u);
          }
        }
#pragma unroll
        for (int q = 0; q < kMissBatch; ++q) {
          const int u = u0 + q;
          if (loc[u] >= 0 || loc[u] == kDead) continue;
          const unsigned b = (unsigned)~loc[u];
          int j = -1;
#pragma unroll
          for (int w = (int)kBucketKeys - 1; w >= 0; --w)
            if (bk[q].k[w] == k[u]) j = w;
          if (j >= 0) {
            fold(&tab.acc[b + j], v[u], FOLD);
            if (WITH_CNT) atomicAdd(&tab.cnt[b + j], 1ull);
          } else if (!global_fold_slow<FOLD, WITH_CNT>(tab, b, k[u], v[u], 1ull)) {
            flags[1] = 1;
          }
        }
      }
    }
  }
  __syncthreads();
  for (unsigned i = threadIdx.x; i < kCacheSlots4; i += kFastThreads4) {
    const unsigned long long ck = cache.key[i];
    if (ck == kEmptyKey) continue;
    const unsigned b = (mix_key(ck) & tab.mask) & ~(kBucketKeys - 1);
    if (!global_fold_slow<FOLD, WITH_CNT>(tab, b, ck, (int64_t)cache.acc[i], WITH_CNT ? (unsigned long long)ccnt[i] : 0ull))
      flags[1] = 1;
  }
}
