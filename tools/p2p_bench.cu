// In-kernel peer-store bandwidth over NVLink as a function of store width, run length and local/remote mix
// (sizes the fused partition+exchange scatter, libgdf_b200/csrc/join_part.cu).  Not part of the product.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/bin/p2p_bench tools/p2p_bench.cu
// Needs two GPUs with peer access: single process, cudaDeviceEnablePeerAccess.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); return 1; } } while (0)

// Every warp writes RUN-byte runs; run r of the grid goes to slot perm(r) of the destination, so that consecutive runs of
// one warp land far apart (like partition runs do).  W = bytes per lane per store (8 or 16).
// remote_every: run r goes to `remote` when (r % remote_every) != 0 ... see `mix` below.
template <int W>
__global__ void store_runs(unsigned char* local, unsigned char* remote, size_t total_bytes, unsigned run_bytes, unsigned remote_num,
                           unsigned remote_den, unsigned misalign) {
  const size_t runs = total_bytes / run_bytes;
  const unsigned lane = threadIdx.x & 31u;
  const size_t warp = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const size_t warps = (size_t)gridDim.x * (blockDim.x >> 5);
  for (size_t r = warp; r < runs; r += warps) {
    const size_t slot = (r * 2654435761ull) % runs;  // scatter the runs
    const bool to_remote = (r % remote_den) < remote_num;
    unsigned char* base = (to_remote ? remote : local) + slot * run_bytes + misalign;
    for (unsigned off = lane * W; off + W <= run_bytes - (misalign ? 64 : 0); off += 32 * W) {
      if (W == 16) *reinterpret_cast<uint4*>(base + off) = make_uint4((unsigned)r, lane, off, 7u);
      else *reinterpret_cast<uint2*>(base + off) = make_uint2((unsigned)r, lane);
    }
  }
}

// "split": what the join's element-parallel copy-out does - a warp instruction stores 32 consecutive staged 8-byte
// elements, runs are L elements long (L odd, so run boundaries drift through the warp) and start at odd/even offsets.
__global__ void store_split(unsigned char* local, unsigned char* remote, size_t total_bytes, unsigned L, unsigned remote_num,
                            unsigned remote_den) {
  const size_t elems = total_bytes / 8, slots = elems / 64;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x; j < elems; j += stride) {
    const size_t run = j / L, within = j % L;
    const size_t slot = (run * 2654435761ull) % slots;
    const bool to_remote = (run % remote_den) < remote_num;
    uint2* dst = reinterpret_cast<uint2*>(to_remote ? remote : local) + slot * 64 + (run & 1u) + within;
    *dst = make_uint2((unsigned)j, (unsigned)run);
  }
}

// "bulk": the same runs written by the TMA engine - one cp.async.bulk shared -> global per run, issued by one lane.
__global__ void store_bulk(unsigned char* local, unsigned char* remote, size_t total_bytes, unsigned run_bytes, unsigned remote_num,
                           unsigned remote_den, unsigned extra_plain = 0, unsigned shift_bytes = 0, unsigned stride_bytes = 0) {
  extern __shared__ __align__(128) unsigned char stage[];
  for (unsigned i = threadIdx.x; i < 32768 / 4; i += blockDim.x) reinterpret_cast<unsigned*>(stage)[i] = i;
  __syncthreads();
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  const size_t runs = total_bytes / (stride_bytes ? stride_bytes : run_bytes);
  const unsigned lane = threadIdx.x & 31u, warp_in_cta = threadIdx.x >> 5;
  const size_t warp = (size_t)blockIdx.x * (blockDim.x >> 5) + warp_in_cta;
  const size_t warps = (size_t)gridDim.x * (blockDim.x >> 5);
  unsigned n = 0;
  for (size_t r = warp * 32 + lane; r < runs; r += warps * 32, ++n) {   // every lane issues its own run
    const size_t slot = (r * 2654435761ull) % runs;
    const bool to_remote = (r % remote_den) < remote_num;
    unsigned char* dst = (to_remote ? remote : local) + slot * (stride_bytes ? stride_bytes : run_bytes) + shift_bytes;
    const unsigned src = (unsigned)__cvta_generic_to_shared(stage + ((r * run_bytes) & 16383u & ~15u));
    if (extra_plain) {  // the join scatter's odd head / tail pairs: plain 8-byte stores next to the bulk store
      const unsigned body = run_bytes - 16;
      *reinterpret_cast<uint2*>(dst) = make_uint2((unsigned)r, 1u);
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + 16), "r"(src), "r"(body) : "memory");
      if (extra_plain > 1) *reinterpret_cast<uint2*>(dst + 8) = make_uint2((unsigned)r, 2u);
    } else
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(run_bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    if ((n & 7u) == 7u) asm volatile("cp.async.bulk.wait_group.read 4;" ::: "memory");
  }
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

template <int W>
float run(unsigned char* local, unsigned char* remote, size_t bytes, unsigned run_bytes, unsigned num, unsigned den, unsigned misalign,
          int blocks) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  store_runs<W><<<blocks, 256>>>(local, remote, bytes, run_bytes, num, den, misalign);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  for (int i = 0; i < 3; ++i) store_runs<W><<<blocks, 256>>>(local, remote, bytes, run_bytes, num, den, misalign);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  return ms / 3;
}

int main() {
  int n = 0;
  CK(cudaGetDeviceCount(&n));
  if (n < 2) { printf("need 2 GPUs\n"); return 0; }
  int can = 0;
  CK(cudaDeviceCanAccessPeer(&can, 0, 1));
  printf("peer access 0->1: %d\n", can);
  const size_t bytes = 2ull << 30;
  unsigned char *local = nullptr, *remote = nullptr;
  CK(cudaSetDevice(1));
  CK(cudaMalloc(&remote, bytes + 4096));
  CK(cudaSetDevice(0));
  CK(cudaMalloc(&local, bytes + 4096));
  CK(cudaDeviceEnablePeerAccess(1, 0));
  {  // copy engine reference
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    CK(cudaMemcpyPeer(remote, 1, local, 0, bytes));
    cudaEventRecord(e0);
    CK(cudaMemcpyPeerAsync(remote, 1, local, 0, bytes, 0));
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("cudaMemcpyPeer 2 GiB: %.3f ms = %.0f GB/s\n", ms, bytes / ms / 1e6);
  }
  const int blocks = 148 * 8;
  printf("%-8s %-6s %-10s %-8s %-9s %-9s %-12s\n", "width", "run_B", "remote", "misalign", "ms", "GB/s all", "GB/s remote");
  const unsigned run_sizes[] = {256, 1024};
  for (int w = 0; w < 2; ++w)
    for (unsigned rs : run_sizes)
      for (int mix = 0; mix < 4; ++mix)
        for (unsigned mis = 0; mis <= (w ? 16u : 8u); mis += (w ? 16u : 8u)) {
          const unsigned num[] = {1, 1, 7, 0}, den[] = {1, 2, 8, 1};
          const float ms = w ? run<16>(local, remote, bytes, rs, num[mix], den[mix], mis, blocks)
                             : run<8>(local, remote, bytes, rs, num[mix], den[mix], mis, blocks);
          const double all = bytes / ms / 1e6, rem = all * num[mix] / den[mix];
          printf("%-8d %-6u %u/%-8u %-8u %-9.3f %-9.0f %-12.0f\n", w ? 16 : 8, rs, num[mix], den[mix], mis, ms, all, rem);
        }
  printf("\nsplit (element-parallel copy-out, runs of L 8-byte elements)\n%-6s %-10s %-9s %-9s %-12s\n", "L", "remote", "ms", "GB/s all", "GB/s remote");
  for (unsigned L : {17u, 33u, 65u, 129u})
    for (int mix = 0; mix < 3; ++mix) {
      const unsigned num[] = {1, 1, 7}, den[] = {1, 2, 8};
      cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
      store_split<<<blocks, 256>>>(local, remote, bytes, L, num[mix], den[mix]);
      cudaDeviceSynchronize();
      cudaEventRecord(e0);
      for (int i = 0; i < 3; ++i) store_split<<<blocks, 256>>>(local, remote, bytes, L, num[mix], den[mix]);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 3;
      printf("%-6u %u/%-8u %-9.3f %-9.0f %-12.0f\n", L, num[mix], den[mix], ms, bytes / ms / 1e6, bytes / ms / 1e6 * num[mix] / den[mix]);
    }
  printf("\nbulk (cp.async.bulk shared -> global, one per run)\n%-6s %-10s %-9s %-9s %-12s\n", "run_B", "remote", "ms", "GB/s all", "GB/s remote");
  CK(cudaFuncSetAttribute(store_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768));
  for (unsigned rs : {128u, 256u, 512u, 1024u, 4096u})
    for (int mix = 0; mix < 3; ++mix) {
      const unsigned num[] = {1, 1, 7}, den[] = {1, 2, 8};
      cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
      store_bulk<<<148 * 4, 256, 32768>>>(local, remote, bytes, rs, num[mix], den[mix]);
      CK(cudaDeviceSynchronize());
      cudaEventRecord(e0);
      for (int i = 0; i < 3; ++i) store_bulk<<<148 * 4, 256, 32768>>>(local, remote, bytes, rs, num[mix], den[mix]);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 3;
      printf("%-6u %u/%-8u %-9.3f %-9.0f %-12.0f\n", rs, num[mix], den[mix], ms, bytes / ms / 1e6, bytes / ms / 1e6 * num[mix] / den[mix]);
    }
  printf("\nbulk + plain 8-byte head/tail stores per run (one direction)\n%-8s %-6s %-10s %-9s %-12s\n", "plain", "run_B", "remote", "ms", "GB/s remote");
  for (unsigned extra = 0; extra <= 2; ++extra)
    for (unsigned rs : {256u, 512u}) {
      cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
      store_bulk<<<148 * 4, 256, 32768>>>(local, remote, bytes, rs, 1, 2, extra);
      CK(cudaDeviceSynchronize());
      cudaEventRecord(e0);
      for (int i = 0; i < 3; ++i) store_bulk<<<148 * 4, 256, 32768>>>(local, remote, bytes, rs, 1, 2, extra);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 3;
      printf("%-8u %-6u 1/2        %-9.3f %-12.0f\n", extra, rs, ms, bytes / ms / 1e6 / 2);
    }
  printf("\nbulk stores whose runs are not line-aligned (slots 512 B apart, 1/2 remote)\n%-8s %-8s %-9s %-14s\n", "run_B", "shift_B", "ms", "GB/s remote (payload)");
  for (unsigned rs : {256u, 272u, 320u})
    for (unsigned sh : {0u, 16u, 32u, 64u}) {
      cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
      store_bulk<<<148 * 4, 256, 32768>>>(local, remote, bytes, rs, 1, 2, 0, sh, 512);
      CK(cudaDeviceSynchronize());
      cudaEventRecord(e0);
      for (int i = 0; i < 3; ++i) store_bulk<<<148 * 4, 256, 32768>>>(local, remote, bytes, rs, 1, 2, 0, sh, 512);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 3;
      printf("%-8u %-8u %-9.3f %-14.0f\n", rs, sh, ms, (double)(bytes / 512) * rs / ms / 1e6 / 2);
    }
  // ---- both directions at once: GPU 0 -> 1 and GPU 1 -> 0 (what an all-to-all exchange does) ----
  {
    unsigned char *local1 = nullptr, *remote0 = nullptr;   // buffers for the kernel running on GPU 1
    CK(cudaSetDevice(1));
    CK(cudaMalloc(&local1, bytes + 4096));
    CK(cudaDeviceEnablePeerAccess(0, 0));
    CK(cudaFuncSetAttribute(store_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768));
    CK(cudaSetDevice(0));
    CK(cudaMalloc(&remote0, bytes + 4096));                // on GPU 0, written by GPU 1
    printf("\nbidirectional (both GPUs store to each other at the same time)\n%-8s %-6s %-10s %-9s %-14s\n", "kind", "run_B", "remote", "ms", "GB/s remote/dir");
    for (int kind = 0; kind < 2; ++kind)
      for (unsigned rs : {256u, 1024u})
        for (int mix = 0; mix < 2; ++mix) {
          const unsigned num[] = {1, 1}, den[] = {1, 2};
          cudaEvent_t a0, a1, b0, b1;
          CK(cudaSetDevice(0)); cudaEventCreate(&a0); cudaEventCreate(&a1);
          CK(cudaSetDevice(1)); cudaEventCreate(&b0); cudaEventCreate(&b1);
          for (int rep = 0; rep < 2; ++rep) {   // rep 0 = warm-up
            CK(cudaSetDevice(0)); cudaEventRecord(a0);
            for (int i = 0; i < 3; ++i) {
              if (kind) store_bulk<<<148 * 4, 256, 32768>>>(local, remote, bytes, rs, num[mix], den[mix]);
              else store_split<<<blocks, 256>>>(local, remote, bytes, rs / 8 + 1, num[mix], den[mix]);
            }
            cudaEventRecord(a1);
            CK(cudaSetDevice(1)); cudaEventRecord(b0);
            for (int i = 0; i < 3; ++i) {
              if (kind) store_bulk<<<148 * 4, 256, 32768>>>(local1, remote0, bytes, rs, num[mix], den[mix]);
              else store_split<<<blocks, 256>>>(local1, remote0, bytes, rs / 8 + 1, num[mix], den[mix]);
            }
            cudaEventRecord(b1);
            CK(cudaSetDevice(0)); CK(cudaDeviceSynchronize());
            CK(cudaSetDevice(1)); CK(cudaDeviceSynchronize());
          }
          float ma, mb;
          cudaEventElapsedTime(&ma, a0, a1); cudaEventElapsedTime(&mb, b0, b1);
          const float ms = (ma > mb ? ma : mb) / 3;
          printf("%-8s %-6u %u/%-8u %-9.3f %-14.0f\n", kind ? "bulk" : "split", rs, num[mix], den[mix], ms, bytes / ms / 1e6 * num[mix] / den[mix]);
        }
  }
  return 0;
}
