// TMA bulk staging for streaming scans (sm_100a).
//
// Every hot kernel of this library is a scan over one or two contiguous column ranges.  Instead of
// each thread issuing its own 128-bit loads (2048 LDG instructions for a 32 KB tile, all of whose
// results sit in registers until consumed), ONE elected thread issues a 1-D bulk asynchronous copy
// (cp.async.bulk.shared::cluster.global, SASS UBLKCP) per tile into a shared-memory ring, and the
// copy engine signals an mbarrier with the byte count when the tile has landed.  The CTA computes on
// stage s while stages s+1 .. s+S-1 are in flight, so the memory pipe stays full across the
// barriers / look-backs / atomics of the consuming code.
//
// Requirements of the instruction: source, destination and size are multiples of 16 bytes.  Callers
// fall back to direct loads for unaligned columns and for the ragged last tile.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace b200 {
namespace tma {

static __device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}

static __device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
// make mbarrier initialisation visible to the async proxy (the copy engine)
static __device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
// order prior generic-proxy accesses to shared memory before later async-proxy (bulk copy) accesses
static __device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
static __device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
static __device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// global -> shared bulk copy, completion counted in bytes on `bar`.  L2 hint: streamed once.
static __device__ __forceinline__ void bulk_load(void* smem_dst, const void* gmem_src, uint32_t bytes,
                                                 uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(smem_dst)),
      "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
// shared -> global bulk store (bulk async-group completion)
static __device__ __forceinline__ void bulk_store(void* gmem_dst, const void* smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst),
               "r"(smem_u32(smem_src)), "r"(bytes)
               : "memory");
}
static __device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
static __device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}

// (the destination of bulk_store may be peer memory: the TMA engine writes whole lines over NVLink, tools/p2p_bench.cu)
static __device__ __forceinline__ void bulk_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
static __device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
}  // namespace tma
}  // namespace b200
