// Caching device-memory allocator shared by librmm.so (PoolAllocation mode) and libgdf.so
// (library-internal scratch).
//
// Why not just cudaMallocAsync?  The driver pool was the first implementation.  It is fast in the
// steady state, but with the join's multi-GB blocks (12 GB of partitioned pairs, 4 GB of tables, 8 GB
// of output) it intermittently re-maps physical memory between virtual ranges when a request does
// not fit a cached range exactly: single calls jumped from 40 ms to 130-290 ms on the host side with
// identical GPU work (profiles/r02_notes.md).  The reference's pool (cnmem, ref
// src/memory/memory.cpp:120-158) never goes back to the driver once warm; this cache restores that
// property: a freed block is kept and handed back for the next request of (about) the same size, so
// repeated operator calls perform no driver allocation at all.
//
// Stream safety (ref cnmem keeps one child pool per registered stream, memory.cpp:120-158): a block
// freed on stream A may still be in use by work queued on A.  With `track_streams` the cache records an
// event on A when the block is freed; a later request on the SAME stream gets the block at once (stream
// order protects it), a request on ANOTHER stream first makes that stream wait for the event.  libgdf's
// scratch cache issues everything on the legacy default stream and runs without tracking.
//
// Bounded: at most `limit` bytes stay cached per process (default: half of the device's memory, set on
// first use); blocks released beyond that go back to the driver, largest first.  trim() returns everything.
#pragma once
#include <cuda_runtime_api.h>

#include <map>
#include <mutex>
#include <unordered_map>
#include <vector>

namespace b200 {

class BlockCache {
 public:
  explicit BlockCache(bool track_streams = false) : track_(track_streams) {}

  cudaError_t alloc(void** out, size_t bytes, cudaStream_t stream = 0) {
    if (bytes == 0) bytes = 1;
    bytes = (bytes + 511) & ~(size_t)511;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    {
      std::unique_lock<std::mutex> g(mu_);
      auto& free_list = free_[dev];
      // accept a cached block up to 12.5 % (+1 MB) larger than the request; prefer one freed on `stream`
      const size_t hi = bytes + bytes / 8 + (1u << 20);
      auto pick = free_list.end();
      for (auto it = free_list.lower_bound(bytes); it != free_list.end() && it->first <= hi; ++it) {
        if (pick == free_list.end()) pick = it;
        if (!track_ || it->second.stream == stream) {
          pick = it;
          break;
        }
      }
      if (pick != free_list.end()) {
        const Free f = pick->second;
        const size_t got = pick->first;
        free_list.erase(pick);
        cached_ -= got;
        live_[f.ptr] = Live{got, dev};
        g.unlock();
        if (f.event) {
          if (f.stream != stream) cudaStreamWaitEvent(stream, f.event, 0);  // order the new user after the old one
          recycle_event(f.event);
        }
        *out = f.ptr;
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
  bool release(void* p, cudaStream_t stream = 0) {
    if (!p) return true;
    Live l;
    {
      std::lock_guard<std::mutex> g(mu_);
      auto it = live_.find(p);
      if (it == live_.end()) return false;
      l = it->second;
      live_.erase(it);
    }
    Free f{p, stream, nullptr};
    if (track_) {
      f.event = new_event();
      if (f.event && cudaEventRecord(f.event, stream) != cudaSuccess) {  // e.g. a destroyed stream: be conservative
        cudaGetLastError();
        cudaDeviceSynchronize();
        recycle_event(f.event);
        f.event = nullptr;
      }
    }
    std::vector<Free> victims;
    {
      std::lock_guard<std::mutex> g(mu_);
      free_[l.device].emplace(l.bytes, f);
      cached_ += l.bytes;
      if (limit_ == 0) limit_ = default_limit();
      auto& fl = free_[l.device];
      while (cached_ > limit_ && !fl.empty()) {  // over budget: largest blocks of this device go back to the driver
        auto big = std::prev(fl.end());
        cached_ -= big->first;
        victims.push_back(big->second);
        fl.erase(big);
      }
    }
    for (const Free& v : victims) destroy(v);
    return true;
  }

  // cudaFree every cached (not in-use) block.
  void trim() {
    std::map<int, std::multimap<size_t, Free>> victims;
    {
      std::lock_guard<std::mutex> g(mu_);
      victims.swap(free_);
      cached_ = 0;
    }
    int cur = 0;
    cudaGetDevice(&cur);
    for (auto& per_dev : victims) {
      if (per_dev.second.empty()) continue;
      cudaSetDevice(per_dev.first);
      for (auto& kv : per_dev.second) destroy(kv.second);
    }
    cudaSetDevice(cur);
    cudaGetLastError();
  }

  size_t cached_bytes() {
    std::lock_guard<std::mutex> g(mu_);
    return cached_;
  }
  void set_limit(size_t bytes) {
    std::lock_guard<std::mutex> g(mu_);
    limit_ = bytes ? bytes : 1;
  }

 private:
  struct Live {
    size_t bytes;
    int device;
  };
  struct Free {
    void* ptr;
    cudaStream_t stream;
    cudaEvent_t event;  // recorded on `stream` when the block was freed (track_streams only)
  };
  static size_t default_limit() {
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) {
      cudaGetLastError();
      return (size_t)64 << 30;
    }
    return total_b / 2;
  }
  void destroy(const Free& f) {
    if (f.event) {
      cudaEventSynchronize(f.event);  // the old user's work must be done before the memory goes back to the driver
      recycle_event(f.event);
    }
    cudaFree(f.ptr);
  }
  // Events are created per free and destroyed on reuse (about a microsecond each): an event belongs to the
  // device it was created on, so a process-wide pool would have to be keyed by device for no measurable gain.
  static cudaEvent_t new_event() {
    cudaEvent_t e = nullptr;
    if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) {
      cudaGetLastError();
      return nullptr;
    }
    return e;
  }
  static void recycle_event(cudaEvent_t e) { cudaEventDestroy(e); }

  const bool track_;
  std::mutex mu_;
  size_t cached_ = 0, limit_ = 0;
  std::map<int, std::multimap<size_t, Free>> free_;
  std::unordered_map<void*, Live> live_;
};

}  // namespace b200
