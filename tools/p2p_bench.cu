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
  const unsigned run_sizes[] = {256, 512, 1024, 4096, 65536};
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
  return 0;
}
