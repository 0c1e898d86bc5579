/**
 * @file load_store.hxx
 * @brief thread::load / thread::store used by graph accessors and user lambdas (reference:
 * include/gunrock/util/load_store.hxx:20-43, which forwards to cub::ThreadLoad/ThreadStore). Here they are
 * plain PTX-level accesses selected by a small modifier enum — no CUB.
 */
#pragma once

namespace gunrock {
namespace thread {

enum class cache_t { standard, read_only, streaming, volatile_, global /* L2 only: coherent across SMs */ };

template <cache_t modifier = cache_t::standard, typename type_t>
__host__ __device__ __forceinline__ type_t load(type_t* ptr) {
#ifdef __CUDA_ARCH__
  if constexpr (modifier == cache_t::read_only)
    return __ldg(ptr);
  else if constexpr (modifier == cache_t::streaming)
    return __ldcs(ptr);
  else if constexpr (modifier == cache_t::volatile_)
    return *reinterpret_cast<volatile type_t*>(ptr);
  else if constexpr (modifier == cache_t::global)
    return __ldcg(ptr);
  else
    return *ptr;
#else
  return *ptr;
#endif
}

template <cache_t modifier = cache_t::standard, typename type_t>
__host__ __device__ __forceinline__ void store(type_t* ptr, const type_t& value) {
#ifdef __CUDA_ARCH__
  if constexpr (modifier == cache_t::streaming)
    __stcs(ptr, value);
  else
    *ptr = value;
#else
  *ptr = value;
#endif
}

}  // namespace thread
}  // namespace gunrock
