// Single-key group-by build kernel (v3).  Included by groupby.cu inside its anonymous namespace,
// after FastSlot / fold() / hash_key() / smem_add64() are defined.
//
// Changes against the first version, driven by profiles/r01_groupby_ncu.md:
//   * work is handed out dynamically, 2048 rows per warp per grab (one global atomic per grab):
//     with a static split 52 % of all warp samples sat at the final CTA barrier waiting for the
//     slowest warps;
//   * 2 CTAs x 768 threads per SM (48 warps) instead of 1 x 1024: the kernel is latency-bound
//     (0.34 eligible warps per scheduler), so more resident warps hide more L2 latency; the per-CTA
//     cache shrinks to 4096 slots (80 KB) to make room;
//   * the first global-table slot of every row that missed the cache is loaded for all U rows of the
//     step before any of them is resolved (U independent L2 requests in flight per thread).
#pragma once

constexpr int kFastThreads3 = 768;
constexpr unsigned kCacheSlots3 = 4096;
constexpr unsigned kCacheProbes3 = 4;
constexpr int kFastU = 4;                                   // rows per thread per step
constexpr unsigned kGrabRows = 32 * kFastU * 16;            // rows per warp per work grab

struct FastCache3 {
  unsigned long long key[kCacheSlots3];
  unsigned long long acc[kCacheSlots3];
  unsigned cnt[kCacheSlots3];
};

// Resolve one row against the global table, starting from an already loaded first slot key.
static __device__ __noinline__ bool global_fold_from(FastSlot* __restrict__ tab, unsigned long long* __restrict__ cnt,
                                                        unsigned mask, unsigned probe_limit, unsigned s,
                                                        unsigned long long k0, unsigned long long key, int64_t v,
                                                        unsigned long long c, int fold_op) {
  for (unsigned probe = 0; probe < probe_limit; ++probe) {
    if (k0 == kEmptyKey) {
      const unsigned long long prev = atomicCAS(&tab[s].key, kEmptyKey, key);
      k0 = (prev == kEmptyKey) ? key : prev;
    }
    if (k0 == key) {
      fold(&tab[s].acc, v, fold_op);
      if (cnt) atomicAdd(&cnt[s], c);
      return true;
    }
    s = (s + 1) & mask;
    k0 = tab[s].key;
  }
  return false;
}

template <typename KT, typename IT>
__global__ void __launch_bounds__(kFastThreads3, 2)
build_fast_kernel_v3(const KT* __restrict__ key_col, const IT* __restrict__ values, size_t n, int op,
                     FastSlot* __restrict__ tab, unsigned long long* __restrict__ cnt, unsigned mask, unsigned slots,
                     unsigned probe_limit, int* __restrict__ flags /*[0]=side slot used, [1]=overflow*/,
                     unsigned long long* __restrict__ work_counter) {
  using UK = typename std::conditional<sizeof(KT) == 8, unsigned long long, unsigned>::type;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  FastCache3& cache = *reinterpret_cast<FastCache3*>(smem_raw);
  const int fold_op = (op == OP_COUNT || op == OP_AVG) ? OP_SUM : op;
  const bool additive = fold_op == OP_SUM;
  const int64_t identity = op == OP_MIN ? (int64_t)std::numeric_limits<IT>::max()
                                        : (op == OP_MAX ? (int64_t)std::numeric_limits<IT>::lowest() : 0);
  for (unsigned i = threadIdx.x; i < kCacheSlots3; i += kFastThreads3) {
    cache.key[i] = kEmptyKey;
    cache.acc[i] = (unsigned long long)identity;
    cache.cnt[i] = 0;
  }
  __syncthreads();
  const unsigned lane = lane_id();
  while (true) {
    unsigned long long grab = 0;
    if (lane == 0) grab = atomicAdd(work_counter, 1ull);
    grab = __shfl_sync(0xffffffffu, grab, 0);
    const size_t base = (size_t)grab * kGrabRows;
    if (base >= n) break;
    const size_t end = base + kGrabRows < n ? base + kGrabRows : n;
#pragma unroll 1
    for (size_t step = base; step < end; step += 32 * kFastU) {
      unsigned long long k[kFastU];
      int64_t v[kFastU];
      uint32_t h[kFastU];
      unsigned gs[kFastU];
      unsigned long long gk[kFastU];
      bool miss[kFastU];
#pragma unroll
      for (int u = 0; u < kFastU; ++u) {
        const size_t r = step + (size_t)u * 32 + lane;
        const bool live = r < end;
        k[u] = live ? (unsigned long long)(UK)key_col[r] : 0ull;
        v[u] = (live && op != OP_COUNT) ? (int64_t)values[r] : 1;
        miss[u] = live;
        h[u] = 0;
      }
      // ---- phase 1: per-CTA shared-memory cache ----
#pragma unroll
      for (int u = 0; u < kFastU; ++u) {
        if (!miss[u]) continue;
        h[u] = hash_key<KT>(k[u]);
        if (k[u] == kEmptyKey) continue;  // the EMPTY-pattern key is handled by the global side slot
        unsigned s = (h[u] >> 7) & (kCacheSlots3 - 1);
#pragma unroll
        for (unsigned p = 0; p < kCacheProbes3; ++p, s = (s + 1) & (kCacheSlots3 - 1)) {
          unsigned long long ck = cache.key[s];
          if (ck == kEmptyKey) {
            const unsigned long long prev = atomicCAS(&cache.key[s], kEmptyKey, k[u]);
            ck = (prev == kEmptyKey) ? k[u] : prev;
          }
          if (ck == k[u]) {
            if (additive) smem_add64(&cache.acc[s], v[u]);
            else if (fold_op == OP_MIN) atomicMin(reinterpret_cast<long long*>(&cache.acc[s]), (long long)v[u]);
            else atomicMax(reinterpret_cast<long long*>(&cache.acc[s]), (long long)v[u]);
            if (cnt) atomicAdd(&cache.cnt[s], 1u);
            miss[u] = false;
            break;
          }
        }
      }
      // ---- phase 2: rows that missed go to the L2-resident table; first slots fetched together ----
#pragma unroll
      for (int u = 0; u < kFastU; ++u) {
        gs[u] = h[u] & mask;
        gk[u] = (miss[u] && k[u] != kEmptyKey) ? tab[gs[u]].key : kEmptyKey;
      }
#pragma unroll
      for (int u = 0; u < kFastU; ++u) {
        if (!miss[u]) continue;
        if (k[u] == kEmptyKey) {
          fold(&tab[slots].acc, v[u], fold_op);
          if (cnt) atomicAdd(&cnt[slots], 1ull);
          flags[0] = 1;
        } else if (!global_fold_from(tab, cnt, mask, probe_limit, gs[u], gk[u], k[u], v[u], 1ull, fold_op)) {
          flags[1] = 1;
        }
      }
    }
  }
  __syncthreads();
  for (unsigned i = threadIdx.x; i < kCacheSlots3; i += kFastThreads3) {
    const unsigned long long ck = cache.key[i];
    if (ck == kEmptyKey) continue;
    const unsigned s = hash_key<KT>(ck) & mask;
    if (!global_fold_from(tab, cnt, mask, probe_limit, s, tab[s].key, ck, (int64_t)cache.acc[i],
                          (unsigned long long)cache.cnt[i], fold_op))
      flags[1] = 1;
  }
}
