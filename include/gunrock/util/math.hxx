/**
 * @file math.hxx
 * @brief math::atomic::{add,min,max,cas,exch} — every one returns the OLD value, as user lambdas rely on
 * (reference: include/gunrock/util/math.hxx:77-129; float/double min/max in
 * include/gunrock/cuda/atomic_functions.hxx:36-123 are CAS loops).
 *
 * B200 notes. float min/max: for the non-negative, non-NaN values SSSP produces, IEEE-754 bit patterns
 * order like signed integers, so a single atom.min.s32 replaces the reference's CAS retry loop (hub
 * vertices no longer serialise on retries); a sign test keeps negative values correct by falling back to
 * the unsigned-max trick, and NaN inputs take the CAS loop. When the caller ignores the result the
 * compiler emits RED instead of ATOM.
 */
#pragma once

#include <algorithm>
#include <cmath>
#include <cstdlib>

namespace gunrock {
namespace gcuda {

template <typename type_t>
__device__ __forceinline__ type_t atomicMin(type_t* address, type_t value) {
  return ::atomicMin(address, value);
}
template <typename type_t>
__device__ __forceinline__ type_t atomicMax(type_t* address, type_t value) {
  return ::atomicMax(address, value);
}

// float: ordered-int trick. Non-negative floats compare like int32; negative ones like reversed uint32.
__device__ __forceinline__ float atomicMin(float* address, float value) {
  if (value != value) {  // NaN: keep the reference's fminf semantics through a CAS loop
    int* a = reinterpret_cast<int*>(address);
    int seen = *a, expect;
    do {
      expect = seen;
      seen = ::atomicCAS(a, expect, __float_as_int(::fminf(value, __int_as_float(expect))));
    } while (seen != expect);
    return __int_as_float(seen);
  }
  return (__float_as_int(value) >= 0)
             ? __int_as_float(::atomicMin(reinterpret_cast<int*>(address), __float_as_int(value)))
             : __uint_as_float(::atomicMax(reinterpret_cast<unsigned int*>(address), __float_as_uint(value)));
}
__device__ __forceinline__ float atomicMax(float* address, float value) {
  if (value != value) {
    int* a = reinterpret_cast<int*>(address);
    int seen = *a, expect;
    do {
      expect = seen;
      seen = ::atomicCAS(a, expect, __float_as_int(::fmaxf(value, __int_as_float(expect))));
    } while (seen != expect);
    return __int_as_float(seen);
  }
  return (__float_as_int(value) >= 0)
             ? __int_as_float(::atomicMax(reinterpret_cast<int*>(address), __float_as_int(value)))
             : __uint_as_float(::atomicMin(reinterpret_cast<unsigned int*>(address), __float_as_uint(value)));
}

__device__ __forceinline__ double atomicMin(double* address, double value) {
  if (value != value) {
    unsigned long long* a = reinterpret_cast<unsigned long long*>(address);
    unsigned long long seen = *a, expect;
    do {
      expect = seen;
      seen = ::atomicCAS(a, expect, (unsigned long long)__double_as_longlong(::fmin(value, __longlong_as_double(expect))));
    } while (seen != expect);
    return __longlong_as_double(seen);
  }
  return (__double_as_longlong(value) >= 0)
             ? __longlong_as_double(::atomicMin(reinterpret_cast<long long*>(address), __double_as_longlong(value)))
             : __longlong_as_double((long long)::atomicMax(reinterpret_cast<unsigned long long*>(address),
                                                            (unsigned long long)__double_as_longlong(value)));
}
__device__ __forceinline__ double atomicMax(double* address, double value) {
  if (value != value) {
    unsigned long long* a = reinterpret_cast<unsigned long long*>(address);
    unsigned long long seen = *a, expect;
    do {
      expect = seen;
      seen = ::atomicCAS(a, expect, (unsigned long long)__double_as_longlong(::fmax(value, __longlong_as_double(expect))));
    } while (seen != expect);
    return __longlong_as_double(seen);
  }
  return (__double_as_longlong(value) >= 0)
             ? __longlong_as_double(::atomicMax(reinterpret_cast<long long*>(address), __double_as_longlong(value)))
             : __longlong_as_double((long long)::atomicMin(reinterpret_cast<unsigned long long*>(address),
                                                            (unsigned long long)__double_as_longlong(value)));
}

}  // namespace gcuda

namespace math {

template <typename type_t>
__host__ __device__ __forceinline__ constexpr type_t divide_round_up(type_t const& a, type_t const& b) {
  return (a + b - 1) / b;
}

template <typename type_t>
constexpr type_t log2(const type_t& n) {
  return (n < 2) ? 0 : 1 + log2(n / 2);
}

template <typename type_t>
constexpr const type_t& max(const type_t& a, const type_t& b) {
  return std::max(a, b);
}
template <typename type_t>
constexpr const type_t& min(const type_t& a, const type_t& b) {
  return std::min(a, b);
}

namespace atomic {

namespace detail {
/// Host stand-in of a device read-modify-write: applies `update(old)` in place and hands back the old value,
/// which is what every wrapper below must return (user lambdas compare against it). Not thread safe, like the
/// reference's host branch.
template <typename type_t, typename update_t>
__host__ __device__ __forceinline__ type_t host_rmw(type_t* address, update_t update) {
  const type_t before = *address;
  *address = update(before);
  return before;
}
}  // namespace detail

#ifdef __CUDA_ARCH__
#define GUNROCK_RMW(device_expr, host_update) return device_expr
#else
#define GUNROCK_RMW(device_expr, host_update) \
  return detail::host_rmw(address, [&](type_t const& before) -> type_t { return host_update; })
#endif

template <typename type_t>
__host__ __device__ __forceinline__ type_t add(type_t* address, type_t value) {
  GUNROCK_RMW(atomicAdd(address, value), before + value);
}
template <typename type_t>
__host__ __device__ __forceinline__ type_t min(type_t* address, type_t value) {
  GUNROCK_RMW(gcuda::atomicMin(address, value), value < before ? value : before);
}
template <typename type_t>
__host__ __device__ __forceinline__ type_t max(type_t* address, type_t value) {
  GUNROCK_RMW(gcuda::atomicMax(address, value), before < value ? value : before);
}
template <typename type_t>
__host__ __device__ __forceinline__ type_t cas(type_t* address, type_t compare, type_t value) {
  GUNROCK_RMW(atomicCAS(address, compare, value), before == compare ? value : before);
}
template <typename type_t>
__host__ __device__ __forceinline__ type_t exch(type_t* address, type_t value) {
  GUNROCK_RMW(atomicExch(address, value), value);
}

#undef GUNROCK_RMW

}  // namespace atomic
}  // namespace math
}  // namespace gunrock
