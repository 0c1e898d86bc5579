/**
 * @file memory.hxx
 * @brief memory_space_t and raw allocation helpers (reference: include/gunrock/memory.hxx:34,63-123),
 * plus device_array_t, the RAII device buffer the B200 operators and algorithm state use instead of
 * thrust::device_vector (no Thrust on the hot path).
 */
#pragma once

#include <cstddef>
#include <cstring>
#include <utility>
#include <cuda_runtime_api.h>
#include <gunrock/error.hxx>

namespace gunrock {
namespace memory {

enum memory_space_t { device, host };

template <typename type_t>
inline type_t* allocate(std::size_t bytes, memory_space_t space = memory_space_t::device) {
  void* p = nullptr;
  if (bytes)
    error::throw_if_exception(space == device ? cudaMalloc(&p, bytes) : cudaMallocHost(&p, bytes),
                              "memory::allocate");
  return static_cast<type_t*>(p);
}

template <typename type_t>
inline void free(type_t* p, memory_space_t space = memory_space_t::device) {
  if (p)
    error::throw_if_exception(space == device ? cudaFree((void*)p) : cudaFreeHost((void*)p), "memory::free");
}

template <typename type_t>
__host__ __device__ inline type_t* raw_pointer_cast(type_t* p) {
  return p;
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
  ~device_array_t() {
    if (ptr)
      cudaFree(ptr);
  }

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
      cudaFree(ptr);
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
