/**
 * @file type_limits.hxx
 * @brief The "invalid element" convention of frontiers: -1 for signed integers, max() for unsigned,
 * NaN for floating point, and util::limits::is_valid. Same values as the reference
 * (include/gunrock/util/type_limits.hxx:20-51,59-71) because user lambdas and frontiers exchange them.
 */
#pragma once

#include <cmath>
#include <limits>
#include <type_traits>

namespace gunrock {

template <typename type_t, typename enable_t = void>
struct numeric_limits : std::numeric_limits<type_t> {};

template <typename type_t>
struct numeric_limits<type_t, std::enable_if_t<std::is_integral<type_t>::value>> : std::numeric_limits<type_t> {
  __host__ __device__ constexpr static type_t invalid() {
    if constexpr (std::is_signed<type_t>::value)
      return static_cast<type_t>(-1);
    else
      return std::numeric_limits<type_t>::max();
  }
};

template <typename type_t>
struct numeric_limits<type_t, std::enable_if_t<std::is_floating_point<type_t>::value>>
    : std::numeric_limits<type_t> {
  __host__ __device__ constexpr static type_t invalid() { return std::numeric_limits<type_t>::quiet_NaN(); }
};

namespace util {
namespace limits {

template <typename type_t>
__host__ __device__ __forceinline__ constexpr bool is_valid(type_t value) {
  static_assert(std::is_arithmetic<type_t>::value, "type_t must be an arithmetic type.");
  if constexpr (std::is_integral<type_t>::value)
    return value != gunrock::numeric_limits<type_t>::invalid();
  else
    return value == value;  // NaN is the only value that differs from itself
}

}  // namespace limits
}  // namespace util
}  // namespace gunrock
