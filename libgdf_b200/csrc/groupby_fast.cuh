// Single-key group-by build kernel (v4).  Included by groupby.cu inside its anonymous namespace,
// after FastSlot / fold() / smem_add64() are defined.
//
// History (profiles/r01a_ncu_full_summary.md): v3 ran C4 at 1.58 TB/s with the issue slots 70 % busy -
// about 230 SASS instructions per row (MurmurHash3 per row, a 4-probe cache loop, a non-inlined
// global fold, dynamically indexed register arrays that ptxas spilled to local memory).  The kernel was
// instruction-bound, not memory-bound.  v4 is the same algorithm with the instruction count cut:
//   * the table position comes from a 3-multiply mixer (the hash is internal: nothing observable
//     depends on it, unlike gdf_hash / gdf_hash_partition which keep MurmurHash3);
//   * the per-CTA cache is 2-way set associative and a set is ONE 128-bit shared-memory load;
//   * rows are read with 128-bit loads (two 8-byte rows per load), four rows per lane per step,
//     everything fully unrolled so that all per-row state stays in registers;
//   * the op (sum / min / max), "count rows" and "keep a row count for AVG" are template parameters.
// One CTA of 1024 threads per SM owns an 8192-slot cache (128 KB of shared memory); rows whose key is
// not cached fold straight into the L2-resident global table with red.global.  The input stream is
// staged one step ahead with cp.async (64 KB).  Variants that were measured and lost (miss queue, lock-step
// inline-PTX cache phase, deferred misses with 768 threads): profiles/r01b_experiments.md.
#pragma once

constexpr int kFastThreads4 = 1024;
constexpr unsigned kCacheSets4 = 4096;                 // x 2 ways
constexpr unsigned kCacheSlots4 = 2 * kCacheSets4;
constexpr unsigned kGrabRows4 = 32 * 4 * 32;           // rows per warp per work grab (4096)

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

// Internal position hash.  h's top bits pick the cache set, its low bits the global slot.
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

// Fold (v, c) for `key` into the global table starting at slot s whose key word `k0` is already loaded.
template <int FOLD, bool WITH_CNT>
static __device__ __forceinline__ bool global_fold4(FastSlot* __restrict__ tab, unsigned long long* __restrict__ cnt,
                                                    unsigned mask, unsigned probe_limit, unsigned s,
                                                    unsigned long long k0, unsigned long long key, int64_t v,
                                                    unsigned long long c) {
  for (unsigned probe = 0; probe < probe_limit; ++probe) {
    if (k0 == kEmptyKey) {
      const unsigned long long prev = atomicCAS(&tab[s].key, kEmptyKey, key);
      k0 = (prev == kEmptyKey) ? key : prev;
    }
    if (k0 == key) {
      fold(&tab[s].acc, v, FOLD);
      if (WITH_CNT) atomicAdd(&cnt[s], c);
      return true;
    }
    s = (s + 1) & mask;
    k0 = tab[s].key;
  }
  return false;
}

template <typename KT, typename IT, int FOLD, bool COUNT_ROWS, bool WITH_CNT, bool LEAN>
__global__ void __launch_bounds__(kFastThreads4, 1)
build_fast_kernel_v4(const KT* __restrict__ key_col, const IT* __restrict__ values, size_t n,
                     FastSlot* __restrict__ tab, unsigned long long* __restrict__ cnt, unsigned mask, unsigned slots,
                     unsigned probe_limit, int* __restrict__ flags /*[0]=side slot used, [1]=overflow*/,
                     unsigned long long* __restrict__ work_counter) {
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

  // LEAN (additive folds): 32-bit shared addresses computed once (no generic->shared conversion per access)
  // and the carry of the 64-bit shared-memory add examined at the END of the step, so the ATOMS round trip
  // of a row is not waited for before the next row starts.
  const uint32_t s_key = (uint32_t)__cvta_generic_to_shared(&cache.key[0]);
  const uint32_t s_acc = (uint32_t)__cvta_generic_to_shared(&cache.acc[0]);
  uint32_t lean_old[4], lean_slot[4];
  bool lean_hit[4];
  auto cache_try_lean = [&](int u, unsigned long long k, int64_t v, uint32_t h) -> bool {
    const unsigned set2 = (h >> 20) * 2u;
    ulonglong2 kk;
    asm volatile("ld.shared.v2.u64 {%0,%1}, [%2];" : "=l"(kk.x), "=l"(kk.y) : "r"(s_key + set2 * 8u));
    lean_hit[u] = false;
    int way = -1;
    if (kk.x == k) way = 0;
    else if (kk.y == k) way = 1;
    if (way >= 0) {
      lean_slot[u] = set2 + (unsigned)way;
      asm volatile("atom.shared.add.u32 %0, [%1], %2;"
                   : "=r"(lean_old[u])
                   : "r"(s_acc + lean_slot[u] * 8u), "r"((uint32_t)(unsigned long long)v));
      lean_hit[u] = true;
      if (WITH_CNT) atomicAdd(&ccnt[lean_slot[u]], 1u);
      return true;
    }
    if (kk.x == kEmptyKey || kk.y == kEmptyKey) {  // cold start only: claim a free way
      if (kk.x == kEmptyKey) {
        const unsigned long long prev = atomicCAS(&cache.key[set2], kEmptyKey, k);
        if (prev == kEmptyKey || prev == k) way = 0;
      }
      if (way < 0) {
        const unsigned long long prev = atomicCAS(&cache.key[set2 + 1], kEmptyKey, k);
        if (prev == kEmptyKey || prev == k) way = 1;
      }
      if (way >= 0) {
        cache_fold<FOLD>(&cache.acc[set2 + way], v);
        if (WITH_CNT) atomicAdd(&ccnt[set2 + way], 1u);
        return true;
      }
    }
    return false;
  };

  // one row: cache first, then the global table.  gk = key word of the row's first global slot,
  // loaded by the caller for all rows of a step before any of them is resolved.
  auto cache_try = [&](unsigned long long k, int64_t v, uint32_t h) -> bool {
    const unsigned set = h >> 20;  // top 12 bits
    const ulonglong2 kk = *reinterpret_cast<const ulonglong2*>(&cache.key[2 * set]);
    int way = -1;
    if (kk.x == k) way = 0;
    else if (kk.y == k) way = 1;
    else if (kk.x == kEmptyKey || kk.y == kEmptyKey) {  // cold start only: claim a free way
      if (kk.x == kEmptyKey) {
        const unsigned long long prev = atomicCAS(&cache.key[2 * set], kEmptyKey, k);
        if (prev == kEmptyKey || prev == k) way = 0;
      }
      if (way < 0) {
        const unsigned long long prev = atomicCAS(&cache.key[2 * set + 1], kEmptyKey, k);
        if (prev == kEmptyKey || prev == k) way = 1;
      }
    }
    if (way < 0) return false;
    cache_fold<FOLD>(&cache.acc[2 * set + way], v);
    if (WITH_CNT) atomicAdd(&ccnt[2 * set + way], 1u);
    return true;
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
      uint32_t h[4];
      bool miss[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        h[u] = mix_key(k[u]);
        miss[u] = live[u];
        if (LEAN) lean_hit[u] = false;
        if (live[u] && k[u] != kEmptyKey) miss[u] = LEAN ? !cache_try_lean(u, k[u], v[u], h[u]) : !cache_try(k[u], v[u], h[u]);
      }
#ifdef B200_LAB_GROUPBY   // timing experiments of profiles/r01b_experiments.md (wrong results): -DB200_LAB_GROUPBY
      if (probe_limit & 0x40000000u) continue;   // B200_LAB_GB=4: drop the global path entirely
#endif
      // rows that missed the cache: first global slots fetched together, then resolved
      unsigned long long gk[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) gk[u] = (miss[u] && k[u] != kEmptyKey) ? tab[h[u] & mask].key : kEmptyKey;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (!miss[u]) continue;
        if (k[u] == kEmptyKey) {  // a real key equal to the EMPTY pattern lives in the side slot
          fold(&tab[slots].acc, v[u], FOLD);
          if (WITH_CNT) atomicAdd(&cnt[slots], 1ull);
          flags[0] = 1;
#ifdef B200_LAB_GROUPBY
        } else if (probe_limit & 0x20000000u) {    // B200_LAB_GB=2: load the slot key, skip the fold
          if (gk[u] == 12345ull) flags[1] = 1;
#endif
        } else if (!global_fold4<FOLD, WITH_CNT>(tab, cnt, mask, probe_limit & 0x0fffffffu, h[u] & mask, gk[u], k[u], v[u], 1ull)) {
          flags[1] = 1;
        }
      }
      if (LEAN) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (!lean_hit[u]) continue;
          const uint32_t lo = (uint32_t)(unsigned long long)v[u];
          const uint32_t up = (uint32_t)((unsigned long long)v[u] >> 32) + (uint32_t)((uint32_t)(lean_old[u] + lo) < lean_old[u]);
          if (up) asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(s_acc + lean_slot[u] * 8u + 4u), "r"(up) : "memory");
        }
      }
    }
  }
  __syncthreads();
  for (unsigned i = threadIdx.x; i < kCacheSlots4; i += kFastThreads4) {
    const unsigned long long ck = cache.key[i];
    if (ck == kEmptyKey) continue;
    const unsigned s = mix_key(ck) & mask;
    if (!global_fold4<FOLD, WITH_CNT>(tab, cnt, mask, probe_limit, s, tab[s].key, ck, (int64_t)cache.acc[i],
                                      WITH_CNT ? (unsigned long long)ccnt[i] : 0ull))
      flags[1] = 1;
  }
}
