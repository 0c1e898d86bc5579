/** @file capi_dispatch.hxx  Runtime enum -> compile-time template argument for the C ABI. */
#pragma once
#include <type_traits>
#include "capi_common.hxx"

namespace ess {
using gunrock::operators::load_balance_t;

template <load_balance_t v>
using lb_c = std::integral_constant<load_balance_t, v>;

/// Calls f(lb_c<...>{}) for the four implemented balancers; others are "not supported" like the reference.
template <typename F>
int with_load_balance(int lb, F&& f) {
  switch (lb) {
    case ESS_LB_THREAD_MAPPED: return f(lb_c<load_balance_t::thread_mapped>{});
    case ESS_LB_BLOCK_MAPPED: return f(lb_c<load_balance_t::block_mapped>{});
    case ESS_LB_BUCKETING: return f(lb_c<load_balance_t::bucketing>{});
    case ESS_LB_MERGE_PATH:
    case ESS_LB_MERGE_PATH_V2: return f(lb_c<load_balance_t::merge_path>{});
    default: return fail("Advance type not supported.");
  }
}
}  // namespace ess
