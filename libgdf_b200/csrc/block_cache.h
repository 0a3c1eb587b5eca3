// Caching device-memory allocator shared by librmm.so (PoolAllocation mode) and libgdf.so
// (library-internal scratch).
//
// Why not just cudaMallocAsync?  The driver pool was the first implementation.  It is fast in the
// steady state, but with the join's multi-GB blocks (12 GB of partitioned pairs, 4 GB of tables, 8 GB
// of output) it intermittently re-maps physical memory between virtual ranges when a request does
// not fit a cached range exactly: single calls jumped from 40 ms to 130-290 ms on the host side with
// identical GPU work (profiles/r01_notes.md).  The reference's pool (cnmem, ref
// src/memory/memory.cpp:120-158) never goes back to the driver once warm; this cache restores that
// property: a freed block is kept and handed back for the next request of (about) the same size, so
// repeated operator calls perform no driver allocation at all.
//
// Reuse is immediate, i.e. ordered only by the CUDA stream the work is issued on.  Everything in
// libgdf.so runs on the legacy default stream, which also orders against every blocking stream.
#pragma once
#include <cuda_runtime_api.h>

#include <map>
#include <mutex>
#include <unordered_map>

namespace b200 {

class BlockCache {
 public:
  cudaError_t alloc(void** out, size_t bytes) {
    if (bytes == 0) bytes = 1;
    bytes = (bytes + 511) & ~(size_t)511;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    {
      std::lock_guard<std::mutex> g(mu_);
      auto& free_list = free_[dev];
      auto it = free_list.lower_bound(bytes);
      // accept a cached block up to 12.5 % (+1 MB) larger than the request
      if (it != free_list.end() && it->first <= bytes + bytes / 8 + (1u << 20)) {
        *out = it->second;
        live_[it->second] = Live{it->first, dev};
        free_list.erase(it);
        return cudaSuccess;
      }
    }
    void* p = nullptr;
    e = cudaMalloc(&p, bytes);
    if (e == cudaErrorMemoryAllocation) {  // give cached blocks back and retry once
      cudaGetLastError();
      trim();
      e = cudaMalloc(&p, bytes);
    }
    if (e != cudaSuccess) return e;
    std::lock_guard<std::mutex> g(mu_);
    live_[p] = Live{bytes, dev};
    *out = p;
    return cudaSuccess;
  }

  // Returns false if the pointer was not handed out by this cache.
  bool release(void* p) {
    if (!p) return true;
    std::lock_guard<std::mutex> g(mu_);
    auto it = live_.find(p);
    if (it == live_.end()) return false;
    free_[it->second.device].emplace(it->second.bytes, p);
    live_.erase(it);
    return true;
  }

  // cudaFree every cached (not in-use) block.
  void trim() {
    std::map<int, std::multimap<size_t, void*>> victims;
    {
      std::lock_guard<std::mutex> g(mu_);
      victims.swap(free_);
    }
    int cur = 0;
    cudaGetDevice(&cur);
    for (auto& per_dev : victims) {
      if (per_dev.second.empty()) continue;
      cudaSetDevice(per_dev.first);
      for (auto& kv : per_dev.second) cudaFree(kv.second);
    }
    cudaSetDevice(cur);
    cudaGetLastError();
  }

  size_t cached_bytes() {
    std::lock_guard<std::mutex> g(mu_);
    size_t s = 0;
    for (auto& per_dev : free_)
      for (auto& kv : per_dev.second) s += kv.first;
    return s;
  }

 private:
  struct Live {
    size_t bytes;
    int device;
  };
  std::mutex mu_;
  std::map<int, std::multimap<size_t, void*>> free_;
  std::unordered_map<void*, Live> live_;
};

}  // namespace b200
