/**
 * @file directional.cuh
 * @brief The push/pull operator pair of direction-optimised advance (used by kernels.cuh and pull.cuh).
 */
#pragma once

#include <type_traits>
#include <utility>

namespace gunrock {
namespace operators {
namespace advance {
namespace kernels {

/**
 * @brief Operator pair for direction-optimised advance. `push` is the ordinary advance operator (several
 * threads may race on the same neighbour, so it needs atomics); `pull` is what a bottom-up level calls: the
 * destination vertex is owned by exactly one thread and is known to be outside the visited set, so the
 * same update can usually be a plain store (BFS: `depth[v] = level; return true`) and the thread does not
 * wait for an atomic's round trip. Built with advance::directional(push, pull); a plain lambda is used for
 * both directions. Top-down levels of a direction-optimised advance use the pull form too, AFTER they have
 * claimed the neighbour with an atomic test-and-set on the visited bitmap (exclusive ownership again); the
 * pull form is therefore expected to return true (a false return leaves the vertex marked visited).
 */
template <typename push_t, typename pull_t>
struct directional_operator_t {
  push_t push;
  pull_t pull;
  template <typename V, typename E, typename W>
  __host__ __device__ __forceinline__ bool operator()(V& src, V& dst, E& edge, W& weight) const {
    return push(src, dst, edge, weight);
  }
};

template <typename T, typename = void>
struct has_pull_operator : std::false_type {};
template <typename T>
struct has_pull_operator<T, std::void_t<decltype(std::declval<T>().pull)>> : std::true_type {};

template <typename operator_t, typename vertex_t, typename edge_t, typename weight_t>
__device__ __forceinline__ bool call_pull(operator_t& op, vertex_t src, vertex_t dst, edge_t edge, weight_t weight) {
  if constexpr (has_pull_operator<operator_t>::value)
    return op.pull(src, dst, edge, weight);
  else
    return op(src, dst, edge, weight);
}

/**
 * @brief Two-phase form of an advance operator for latency-bound traversals (advance::two_phase(issue, resolve)).
 * `issue(src, dst, e, w)` starts the update — typically the atomic — and returns a token (e.g. candidate and old
 * value); `resolve(token)` turns it into the operator's bool. Kernels that keep several edges per lane in flight
 * (near_far.cuh) call every issue of a batch before the first resolve, so the atomics' round trips overlap; with a
 * plain operator the compare that consumes one atomic's result sits in front of the next atomic and the in-order
 * warp pays one full L2 round trip per edge (measured: ~1 us each, four in a row per lane on the 2-D grid).
 * Used as a plain operator it is resolve(issue(...)).
 */
template <typename issue_t, typename resolve_t>
struct two_phase_operator_t {
  issue_t issue;
  resolve_t resolve;
  template <typename V, typename E, typename W>
  __host__ __device__ __forceinline__ bool operator()(V& src, V& dst, E& edge, W& weight) const {
    return resolve(issue(src, dst, edge, weight));
  }
};

/**
 * @brief Two-phase operator with a per-SOURCE prologue (advance::two_phase(prepare, issue, resolve)):
 * `prepare(src)` runs once per frontier vertex and its result (e.g. the source's tentative distance) is handed to
 * every `issue(state, src, dst, e, w)` of that vertex — one label read per vertex instead of one per edge, and no
 * load in front of each atomic.
 */
template <typename prepare_t, typename issue_t, typename resolve_t>
struct staged_operator_t {
  prepare_t prepare;
  issue_t issue;
  resolve_t resolve;
  template <typename V, typename E, typename W>
  __host__ __device__ __forceinline__ bool operator()(V& src, V& dst, E& edge, W& weight) const {
    return resolve(issue(prepare(src), src, dst, edge, weight));
  }
};

template <typename T, typename = void>
struct has_prepare : std::false_type {};
template <typename T>
struct has_prepare<T, std::void_t<decltype(std::declval<T>().prepare)>> : std::true_type {};

template <typename T, typename = void>
struct has_two_phase : std::false_type {};
template <typename T>
struct has_two_phase<T, std::void_t<decltype(std::declval<T>().issue), decltype(std::declval<T>().resolve)>>
    : std::true_type {};

}  // namespace kernels
}  // namespace advance
}  // namespace operators
}  // namespace gunrock
