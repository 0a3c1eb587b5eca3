// Building-block throughput probes used to size the group-by / join kernels (results recorded in
// profiles/).  Not part of the product.  nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/bin/microbench tools/microbench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__device__ __forceinline__ uint32_t lcg(uint32_t& s) { s = s * 1664525u + 1013904223u; return s ^ (s >> 15); }

template <typename T>
__global__ void smem_atomic_kernel(int iters, unsigned mask, T* sink) {
  extern __shared__ unsigned char raw[];
  T* tab = reinterpret_cast<T*>(raw);
  for (unsigned i = threadIdx.x; i <= mask; i += blockDim.x) tab[i] = 0;
  __syncthreads();
  uint32_t s = blockIdx.x * 7919u + threadIdx.x * 104729u + 1u;
  for (int i = 0; i < iters; ++i) atomicAdd(&tab[lcg(s) & mask], (T)1);
  __syncthreads();
  if (threadIdx.x == 0) sink[blockIdx.x] = tab[0];
}

__global__ void smem_plain_rmw_kernel(int iters, unsigned mask, unsigned* sink) {  // non-atomic LDS+STS for comparison
  extern __shared__ unsigned char raw[];
  unsigned* tab = reinterpret_cast<unsigned*>(raw);
  for (unsigned i = threadIdx.x; i <= mask; i += blockDim.x) tab[i] = 0;
  __syncthreads();
  uint32_t s = blockIdx.x * 7919u + threadIdx.x * 104729u + 1u;
  for (int i = 0; i < iters; ++i) { unsigned a = lcg(s) & mask; tab[a] = tab[a] + 1; }
  __syncthreads();
  if (threadIdx.x == 0) sink[blockIdx.x] = tab[0];
}

__global__ void gmem_red_kernel(unsigned long long* tab, size_t mask, int iters) {
  uint32_t s = blockIdx.x * 7919u + threadIdx.x * 104729u + 1u;
  for (int i = 0; i < iters; ++i) {
    size_t a = (((size_t)lcg(s) << 16) ^ lcg(s)) & mask;
    atomicAdd(&tab[a], 1ull);
  }
}

__global__ void gmem_probe_red_kernel(ulonglong2* tab, size_t mask, int iters) {  // read key (16B slot) then red on value
  uint32_t s = blockIdx.x * 7919u + threadIdx.x * 104729u + 1u;
  unsigned long long acc = 0;
  for (int i = 0; i < iters; ++i) {
    size_t a = (((size_t)lcg(s) << 16) ^ lcg(s)) & mask;
    unsigned long long k = *reinterpret_cast<volatile unsigned long long*>(&tab[a].x);
    acc += k;
    atomicAdd(&tab[a].y, 1ull);
  }
  if (acc == 0x1234567) tab[0].x = acc;
}

__global__ void gmem_read16_kernel(const uint4* tab, size_t mask, int iters, unsigned* sink) {
  uint32_t s = blockIdx.x * 7919u + threadIdx.x * 104729u + 1u;
  unsigned acc = 0;
  for (int i = 0; i < iters; i += 4) {
    size_t a0 = (((size_t)lcg(s) << 16) ^ lcg(s)) & mask, a1 = (((size_t)lcg(s) << 16) ^ lcg(s)) & mask;
    size_t a2 = (((size_t)lcg(s) << 16) ^ lcg(s)) & mask, a3 = (((size_t)lcg(s) << 16) ^ lcg(s)) & mask;
    uint4 v0 = tab[a0], v1 = tab[a1], v2 = tab[a2], v3 = tab[a3];
    acc += v0.x + v1.y + v2.z + v3.w;
  }
  if (acc == 0x1234567) sink[0] = acc;
}

__global__ void match_kernel(int iters, unsigned mask, unsigned* sink) {
  uint32_t s = blockIdx.x * 7919u + threadIdx.x * 104729u + 1u;
  unsigned acc = 0;
  for (int i = 0; i < iters; ++i) acc += __match_any_sync(0xffffffffu, lcg(s) & mask);
  if (acc == 0x1234567) sink[0] = acc;
}
__global__ void match64_kernel(int iters, unsigned mask, unsigned* sink) {
  uint32_t s = blockIdx.x * 7919u + threadIdx.x * 104729u + 1u;
  unsigned acc = 0;
  for (int i = 0; i < iters; ++i) acc += __match_any_sync(0xffffffffu, (unsigned long long)(lcg(s) & mask) * 0x100000001ull);
  if (acc == 0x1234567) sink[0] = acc;
}


__global__ void gmem_red32_kernel(unsigned* tab, size_t mask, int iters) {
  uint32_t s = blockIdx.x * 7919u + threadIdx.x * 104729u + 1u;
  for (int i = 0; i < iters; ++i) {
    size_t a = (((size_t)lcg(s) << 16) ^ lcg(s)) & mask;
    atomicAdd(&tab[a], 1u);
  }
}
__global__ void gmem_atom32_kernel(unsigned* tab, size_t mask, int iters, unsigned* sink) {  // returning atomic, 4 in flight
  uint32_t s = blockIdx.x * 7919u + threadIdx.x * 104729u + 1u;
  unsigned acc = 0;
  for (int i = 0; i < iters; i += 4) {
    size_t a0 = (((size_t)lcg(s) << 16) ^ lcg(s)) & mask, a1 = (((size_t)lcg(s) << 16) ^ lcg(s)) & mask;
    size_t a2 = (((size_t)lcg(s) << 16) ^ lcg(s)) & mask, a3 = (((size_t)lcg(s) << 16) ^ lcg(s)) & mask;
    unsigned r0 = atomicAdd(&tab[a0], 1u), r1 = atomicAdd(&tab[a1], 1u), r2 = atomicAdd(&tab[a2], 1u), r3 = atomicAdd(&tab[a3], 1u);
    acc += r0 + r1 + r2 + r3;
  }
  if (acc == 0x1234567) sink[0] = acc;
}
// the "spill stream" pattern: every CTA appends 16-byte records to one of `streams` private regions, position
// from a shared-memory cursor, no global atomics
__global__ void append_kernel(uint4* region, unsigned streams, unsigned cap, int iters) {
  extern __shared__ unsigned cur[];
  for (unsigned i = threadIdx.x; i < streams; i += blockDim.x) cur[i] = 0;
  __syncthreads();
  uint32_t s = blockIdx.x * 7919u + threadIdx.x * 104729u + 1u;
  for (int i = 0; i < iters; ++i) {
    const unsigned p = lcg(s) % streams;
    const unsigned pos = atomicAdd(&cur[p], 1u);
    if (pos < cap) region[((size_t)p * gridDim.x + blockIdx.x) * cap + pos] = make_uint4(s, p, pos, i);
  }
}
__global__ void gmem_read32_kernel(const ulonglong4* tab, size_t mask, int iters, unsigned* sink) {  // 32-byte buckets
  uint32_t s = blockIdx.x * 7919u + threadIdx.x * 104729u + 1u;
  unsigned long long acc = 0;
  for (int i = 0; i < iters; i += 4) {
    size_t a[4];
    unsigned long long x[4], y[4], z[4], w[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) a[j] = (((size_t)lcg(s) << 16) ^ lcg(s)) & mask;
#pragma unroll
    for (int j = 0; j < 4; ++j)
      asm volatile("ld.global.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(x[j]), "=l"(y[j]), "=l"(z[j]), "=l"(w[j]) : "l"(tab + a[j]));
#pragma unroll
    for (int j = 0; j < 4; ++j) acc += x[j] ^ y[j] ^ z[j] ^ w[j];
  }
  if (acc == 0x1234567) sink[0] = (unsigned)acc;
}

template <typename F>
float time_ms(F f) {
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  f();  // warm-up
  cudaEventRecord(a);
  f();
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms = 0; cudaEventElapsedTime(&ms, a, b);
  return ms;
}

int main() {
  int sms = 0; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  const int blocks = sms * 8, threads = 256, iters = 4096;
  const double ops = (double)blocks * threads * iters;
  void* sink; CK(cudaMalloc(&sink, 1 << 20));
  printf("SMs=%d blocks=%d threads=%d iters=%d\n", sms, blocks, threads, iters);
  for (unsigned slots : {64u, 1024u, 4096u}) {
    float ms = time_ms([&] { smem_atomic_kernel<unsigned><<<blocks, threads, slots * 4>>>(iters, slots - 1, (unsigned*)sink); });
    printf("smem atomicAdd u32  %5u slots: %8.3f ms  %7.1f Gop/s  (%.2f cyc/lane/SM @1.9GHz)\n", slots, ms, ops / ms * 1e-6, ms * 1e-3 * 1.9e9 * sms / ops);
    ms = time_ms([&] { smem_atomic_kernel<unsigned long long><<<blocks, threads, slots * 8>>>(iters, slots - 1, (unsigned long long*)sink); });
    printf("smem atomicAdd u64  %5u slots: %8.3f ms  %7.1f Gop/s  (%.2f cyc/lane/SM)\n", slots, ms, ops / ms * 1e-6, ms * 1e-3 * 1.9e9 * sms / ops);
    ms = time_ms([&] { smem_plain_rmw_kernel<<<blocks, threads, slots * 4>>>(iters, slots - 1, (unsigned*)sink); });
    printf("smem plain LDS+STS  %5u slots: %8.3f ms  %7.1f Gop/s\n", slots, ms, ops / ms * 1e-6);
  }
  CK(cudaGetLastError());
  for (size_t mb : {32ull, 64ull, 256ull, 4096ull}) {
    size_t bytes = mb << 20;
    unsigned long long* tab; CK(cudaMalloc(&tab, bytes)); CK(cudaMemset(tab, 0, bytes));
    float ms = time_ms([&] { gmem_red_kernel<<<blocks, threads>>>(tab, bytes / 8 - 1, iters / 4); });
    printf("gmem red.add u64 random in %5zu MB: %8.3f ms  %7.1f Gop/s\n", mb, ms, ops / 4 / ms * 1e-6);
    ms = time_ms([&] { gmem_probe_red_kernel<<<blocks, threads>>>((ulonglong2*)tab, bytes / 16 - 1, iters / 4); });
    printf("gmem ld key + red  random in %5zu MB: %8.3f ms  %7.1f Grow/s\n", mb, ms, ops / 4 / ms * 1e-6);
    ms = time_ms([&] { gmem_read16_kernel<<<blocks, threads>>>((const uint4*)tab, bytes / 16 - 1, iters / 4, (unsigned*)sink); });
    printf("gmem ld.128 random         in %5zu MB: %8.3f ms  %7.1f Gld/s  (%.0f GB/s of 32B sectors)\n", mb, ms, ops / 4 / ms * 1e-6, ops / 4 / ms * 1e-6 * 32);
    CK(cudaFree(tab));
  }
  for (size_t mb : {16ull, 64ull, 256ull}) {
    size_t bytes = mb << 20;
    unsigned* tab; CK(cudaMalloc(&tab, bytes)); CK(cudaMemset(tab, 0, bytes));
    float ms = time_ms([&] { gmem_red32_kernel<<<blocks, threads>>>(tab, bytes / 4 - 1, iters / 4); });
    printf("gmem red.add u32 random in %5zu MB: %8.3f ms  %7.1f Gop/s\n", mb, ms, ops / 4 / ms * 1e-6);
    ms = time_ms([&] { gmem_atom32_kernel<<<blocks, threads>>>(tab, bytes / 4 - 1, iters / 4, (unsigned*)sink); });
    printf("gmem atom.add u32 (returning, 4 in flight) random in %5zu MB: %8.3f ms  %7.1f Gop/s\n", mb, ms, ops / 4 / ms * 1e-6);
    ms = time_ms([&] { gmem_read32_kernel<<<blocks, threads>>>((const ulonglong4*)tab, bytes / 32 - 1, iters / 4, (unsigned*)sink); });
    printf("gmem ld.v4.u64 (32 B bucket) random in %5zu MB: %8.3f ms  %7.1f Gld/s\n", mb, ms, ops / 4 / ms * 1e-6);
    CK(cudaFree(tab));
  }
  for (unsigned streams : {148u, 296u, 592u}) {
    const unsigned cap = 16384, ablocks = sms, athreads = 1024;
    const int aiters = (int)((size_t)cap * streams / athreads * 3 / 4);
    uint4* region; CK(cudaMalloc(&region, (size_t)streams * ablocks * cap * 16));
    float ms = time_ms([&] { append_kernel<<<ablocks, athreads, streams * 4>>>(region, streams, cap, aiters); });
    const double recs = (double)ablocks * athreads * aiters;
    printf("append 16 B records to %u streams/CTA (%d CTAs x 1024 thr): %8.3f ms  %7.1f Grec/s  %7.1f GB/s\n", streams, ablocks, ms, recs / ms * 1e-6, recs * 16 / ms * 1e-6);
    CK(cudaFree(region));
  }
  {
    unsigned long long* one; CK(cudaMalloc(&one, 8)); CK(cudaMemset(one, 0, 8));
    float ms = time_ms([&] { gmem_red_kernel<<<blocks, threads>>>(one, 0, iters / 16); });
    printf("gmem red.add u64 SAME address: %8.3f ms  %7.2f Gop/s\n", ms, ops / 16 / ms * 1e-6);
  }
  for (unsigned m : {0xffu, 0xfffffu}) {
    float ms = time_ms([&] { match_kernel<<<blocks, threads>>>(iters, m, (unsigned*)sink); });
    printf("match.any b32 mask %x: %8.3f ms  %7.1f Glane/s\n", m, ms, ops / ms * 1e-6);
    ms = time_ms([&] { match64_kernel<<<blocks, threads>>>(iters, m, (unsigned*)sink); });
    printf("match.any b64 mask %x: %8.3f ms  %7.1f Glane/s\n", m, ms, ops / ms * 1e-6);
  }
  CK(cudaDeviceSynchronize());
  return 0;
}
