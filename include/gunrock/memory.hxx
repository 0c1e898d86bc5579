/**
 * @file memory.hxx
 * @brief memory_space_t and raw allocation helpers (reference: include/gunrock/memory.hxx:34,63-123),
 * plus device_array_t, the RAII device buffer the B200 operators and algorithm state use instead of
 * thrust::device_vector (no Thrust on the hot path).
 */
#pragma once

#include <cstddef>
#include <cstring>
#include <map>
#include <mutex>
#include <unordered_map>
#include <utility>
#include <vector>
#include <cuda_runtime_api.h>
#include <gunrock/error.hxx>

namespace gunrock {
namespace memory {

enum memory_space_t { device, host };

/**
 * @brief Process-wide caching allocator for device memory. Every enactor run needs the same handful of
 * buffers (two frontiers, work offsets, bitmaps, per-vertex state); cudaMalloc/cudaFree cost 0.1-1 ms each
 * and cudaFree synchronises the device, which is more than a whole BFS on B200. Blocks are returned to a
 * per-(device, rounded size) free list instead and handed out again on the next run, so after the first
 * run of a given shape no call reaches the driver.
 * Stream safety (what cudaFree's implicit device synchronisation used to give): every stream the library launches on
 * is registered (standard_context_t does it). release() asks each registered stream of the block's device whether it
 * still has work queued (cudaStreamQuery, ~1 us; the usual answer after an operator's end-of-level synchronisation is
 * "idle") and only for a busy stream records an event that travels with the cached block; acquire() waits for those
 * events on the host before it hands the block out, so a block can never reach a second user — another context,
 * another stream, a legacy-stream memcpy — while kernels enqueued before its release may still touch it.
 * The cache is intentionally never destroyed at exit (the CUDA runtime may already be gone).
 */
class device_pool_t {
 public:
  static device_pool_t& instance() {
    static device_pool_t* pool = new device_pool_t;
    return *pool;
  }
  void* acquire(std::size_t bytes) {
    if (!bytes) return nullptr;
    const std::size_t rounded = round_up(bytes);
    int dev = 0;
    cudaGetDevice(&dev);
    {
      std::lock_guard<std::mutex> lock(mu);
      // best fit: the smallest cached block of this device that holds the request without wasting more than
      // half of itself (frontier buffers grow by data-dependent amounts, so exact sizes rarely repeat)
      for (auto it = cached.lower_bound({dev, rounded}); it != cached.end() && it->first.first == dev; ++it) {
        if (it->first.second > 2 * rounded + (std::size_t(4) << 20)) break;
        if (it->second.empty()) continue;
        block_t blk = it->second.back();
        it->second.pop_back();
        cached_bytes -= it->first.second;
        live[blk.p] = it->first;
        for (cudaEvent_t ev : blk.pending) {  // work that was still queued when the block was released
          cudaEventSynchronize(ev);
          spare_events.push_back(ev);
        }
        return blk.p;
      }
    }
    void* p = nullptr;
    cudaError_t st = cudaMalloc(&p, rounded);
    if (st != cudaSuccess) {  // out of memory: give the cache back to the driver and retry once
      cudaGetLastError();
      trim();
      st = cudaMalloc(&p, rounded);
    }
    error::throw_if_exception(st, "memory::allocate");
    std::lock_guard<std::mutex> lock(mu);
    live[p] = {dev, rounded};
    return p;
  }
  void release(void* p) {
    if (!p) return;
    std::lock_guard<std::mutex> lock(mu);
    auto it = live.find(p);
    if (it == live.end()) {  // not ours (allocated before the pool existed): plain free
      cudaFree(p);
      return;
    }
    block_t blk;
    blk.p = p;
    for (auto& st : streams) {
      if (st.dev != it->second.first) continue;
      const cudaError_t state = cudaStreamQuery(st.stream);
      if (state == cudaSuccess) continue;  // idle: nothing queued can touch the block
      cudaGetLastError();                  // (cudaErrorNotReady is not an error)
      if (state != cudaErrorNotReady) continue;  // a caller-owned stream that no longer exists
      cudaEvent_t ev = nullptr;
      if (!spare_events.empty()) {
        ev = spare_events.back();
        spare_events.pop_back();
      } else if (cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) != cudaSuccess) {
        cudaGetLastError();
        cudaStreamSynchronize(st.stream);  // no event to be had: wait here instead
        continue;
      }
      cudaEventRecord(ev, st.stream);
      blk.pending.push_back(ev);
    }
    cached[it->second].push_back(std::move(blk));
    cached_bytes += it->second.second;
    live.erase(it);
  }
  /// Streams whose queued work release() must respect (see the class comment). The device is the current one.
  void register_stream(cudaStream_t stream) {
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lock(mu);
    for (auto& st : streams)
      if (st.dev == dev && st.stream == stream) {
        ++st.users;
        return;
      }
    streams.push_back({dev, stream, 1});
  }
  void unregister_stream(cudaStream_t stream) {
    std::lock_guard<std::mutex> lock(mu);
    for (std::size_t i = 0; i < streams.size(); ++i)
      if (streams[i].stream == stream) {
        if (--streams[i].users == 0) streams.erase(streams.begin() + i);
        break;
      }
  }
  /// Return every cached (unused) block to the driver.
  void trim() {
    std::lock_guard<std::mutex> lock(mu);
    int cur = 0;
    cudaGetDevice(&cur);
    for (auto& kv : cached) {
      cudaSetDevice(kv.first.first);
      for (auto& blk : kv.second) {
        cudaFree(blk.p);  // synchronises the device: the pending events have fired
        for (cudaEvent_t ev : blk.pending) spare_events.push_back(ev);
      }
      kv.second.clear();
    }
    cudaSetDevice(cur);
    cached_bytes = 0;
  }
  std::size_t bytes_cached() const { return cached_bytes; }

 private:
  static std::size_t round_up(std::size_t b) {
    if (b <= (std::size_t(1) << 20)) {
      std::size_t r = 256;
      while (r < b) r <<= 1;
      return r;
    }
    const std::size_t grain = std::size_t(2) << 20;
    return (b + grain - 1) / grain * grain;
  }
  struct block_t {
    void* p = nullptr;
    std::vector<cudaEvent_t> pending;
  };
  std::mutex mu;
  std::map<std::pair<int, std::size_t>, std::vector<block_t>> cached;
  struct stream_ref_t {
    int dev;
    cudaStream_t stream;
    int users;
  };
  std::vector<stream_ref_t> streams;
  std::vector<cudaEvent_t> spare_events;
  std::unordered_map<void*, std::pair<int, std::size_t>> live;
  std::size_t cached_bytes = 0;
};

template <typename type_t>
inline type_t* allocate(std::size_t bytes, memory_space_t space = memory_space_t::device) {
  void* p = nullptr;
  if (bytes) {
    if (space == device)
      p = device_pool_t::instance().acquire(bytes);
    else
      error::throw_if_exception(cudaMallocHost(&p, bytes), "memory::allocate");
  }
  return static_cast<type_t*>(p);
}

template <typename type_t>
inline void free(type_t* p, memory_space_t space = memory_space_t::device) {
  if (!p) return;
  if (space == device)
    device_pool_t::instance().release((void*)p);
  else
    error::throw_if_exception(cudaFreeHost((void*)p), "memory::free");
}

template <typename type_t>
__host__ __device__ inline type_t* raw_pointer_cast(type_t* p) {
  return p;
}
/// Fancy pointers (thrust::device_ptr and friends) expose the raw pointer through get().
template <typename pointer_t, typename = decltype(std::declval<pointer_t>().get())>
inline auto raw_pointer_cast(pointer_t p) {
  return p.get();
}

/**
 * @brief Owning, growable device buffer of trivially-copyable elements. Move-only.
 * `resize` keeps the old contents (device-to-device copy) when growing; capacity never shrinks.
 */
template <typename type_t>
class device_array_t {
 public:
  device_array_t() = default;
  explicit device_array_t(std::size_t count) { resize(count); }
  device_array_t(const device_array_t&) = delete;
  device_array_t& operator=(const device_array_t&) = delete;
  device_array_t(device_array_t&& o) noexcept { swap(o); }
  device_array_t& operator=(device_array_t&& o) noexcept {
    swap(o);
    return *this;
  }
  ~device_array_t() { memory::free(ptr); }

  void swap(device_array_t& o) noexcept {
    std::swap(ptr, o.ptr);
    std::swap(count, o.count);
    std::swap(cap, o.cap);
  }

  void reserve(std::size_t n, bool keep = true) {
    if (n <= cap)
      return;
    type_t* fresh = allocate<type_t>(n * sizeof(type_t));
    if (ptr) {
      if (keep && count)
        error::throw_if_exception(cudaMemcpy(fresh, ptr, count * sizeof(type_t), cudaMemcpyDeviceToDevice),
                                  "device_array_t::reserve");
      memory::free(ptr);
    }
    ptr = fresh;
    cap = n;
  }
  void resize(std::size_t n) {
    reserve(n);
    count = n;
  }
  type_t* data() const { return ptr; }
  type_t* get() const { return ptr; }
  std::size_t size() const { return count; }
  std::size_t capacity() const { return cap; }
  bool empty() const { return count == 0; }

 private:
  type_t* ptr = nullptr;
  std::size_t count = 0, cap = 0;
};

}  // namespace memory
}  // namespace gunrock
