/**
 * @file capi_frontier.hxx
 * @brief A caller-owned device buffer dressed as a frontier, for the C-ABI entry points that run an operator on
 * raw pointers (probes, multi-GPU level kernels).
 */
#pragma once

#include "capi_common.hxx"

namespace ess {

/// A borrowed device buffer dressed as a frontier: the operators only need data()/size/capacity.
template <typename edge_t>
struct borrowed_frontier_t {
  using type_t = int32_t;
  using offset_t = edge_t;
  int32_t* ptr = nullptr;
  std::size_t count = 0, cap = 0;
  memory::device_array_t<int32_t> own;  // used when the caller gave no (or too small a) buffer
  int32_t* data() { return ptr; }
  std::size_t get_number_of_elements() const { return count; }
  void set_number_of_elements(std::size_t c) { count = c; }
  std::size_t get_capacity() const { return cap; }
  void reserve(std::size_t n) {
    if (n <= cap) return;
    own.reserve(n, false);
    ptr = own.data();
    cap = n;
  }
};

}  // namespace ess
