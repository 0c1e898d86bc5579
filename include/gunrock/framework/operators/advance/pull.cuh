/**
 * @file pull.cuh
 * @brief Bottom-up ("pull") step and dense-state kernels of direction-optimised advance.
 *
 * The reference declares advance_direction_t::optimized but throws when it is requested
 * (advance/merge_path.hxx:41-43) and cannot even hold CSR and CSC together (graph/detail/build.hxx:85-89),
 * so the semantics are defined here (DESIGN.md "direction-optimised advance"):
 * for every vertex v outside the visited set, walk v's IN-edges (CSC); for the first in-neighbour u that is
 * in the current frontier and for which op(u, v, e, w) returns true, v joins the next frontier and the
 * visited set and the walk stops. Frontier, next frontier and visited set are 1-bit-per-vertex maps: at
 * scale-26 each is 8 MiB and stays in the 126 MB L2, so the per-edge membership probe never reaches HBM.
 *
 * Mapping: one warp per 32-vertex word. A fully visited word costs one 4-byte read; otherwise lanes read
 * their row bounds coalesced, walk their lists, and a ballot assembles the next-frontier word, so dense
 * outputs are written without atomics.
 */
#pragma once

#include <gunrock/b200/warp.cuh>
#include <gunrock/cuda/context.hxx>
#include <gunrock/graph/graph.hxx>

namespace gunrock {
namespace operators {
namespace advance {
namespace kernels {

using b200::counter_t;
using gcuda::scratch_t;

/// visited := {isolated vertices} ∪ frontier; also Σdeg(frontier) -> counters[aux2].
template <typename vertex_t, typename edge_t>
__global__ void __launch_bounds__(256)
    init_visited_kernel(const edge_t* __restrict__ offsets, vertex_t n, unsigned* __restrict__ visited) {
  const unsigned lane = b200::lane_id();
  const std::size_t n_words = (std::size_t(n) + 31) / 32;
  const std::size_t warps = (std::size_t(gridDim.x) * blockDim.x) >> 5;
  for (std::size_t w = (std::size_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5; w < n_words; w += warps) {
    const std::size_t v = w * 32 + lane;
    bool skip = true;  // padding bits count as visited so nobody ever walks them
    if (v < std::size_t(n)) skip = offsets[v + 1] == offsets[v];
    const unsigned word = __ballot_sync(b200::full_mask, skip);
    if (lane == 0) visited[w] = word;
  }
}

/// Frontier vertices enter the visited set; accumulates their degree sum (m_f) into counters[aux2].
template <typename vertex_t, typename edge_t>
__global__ void __launch_bounds__(256)
    mark_frontier_kernel(const edge_t* __restrict__ offsets, const vertex_t* __restrict__ list, std::size_t count,
                         const counter_t* count_ptr, unsigned* __restrict__ visited, counter_t* counters) {
  if (count_ptr) count = std::size_t(*count_ptr);
  counter_t mine = 0;
  for (std::size_t i = std::size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < count;
       i += std::size_t(gridDim.x) * blockDim.x) {
    const vertex_t v = list[i];
    if (!util::limits::is_valid(v)) continue;
    if (visited) atomicOr(&visited[unsigned(v) >> 5], 1u << (unsigned(v) & 31u));
    mine += counter_t(offsets[v + 1] - offsets[v]);
  }
  // lanes may have diverged on `continue`; reconverge before the shuffle reduction
  __syncwarp();
  mine = b200::warp_sum(mine);
  if (b200::lane_id() == 0 && mine) atomicAdd(counters + scratch_t::aux2, mine);
}

/**
 * @brief One bottom-up level. counters[out_count] += |next frontier|, counters[aux2] += Σdeg(next frontier),
 * counters[aux0] += unvisited vertices walked, counters[aux1] += in-edges read (work accounting).
 * `A` is the CSC adjacency (for a symmetric graph it aliases the CSR arrays).
 *
 * A warp owns one 32-vertex word per trip; a lane owns one unvisited vertex and walks its in-edges in three
 * stages chosen to keep the slowest lane from holding the other 31 (the first version spent most of its time
 * with 8/32 lanes active, profiles/r01_pull_step_full_raw.csv):
 *   1. batched probes: `pull_batch` neighbours and then their frontier bits are loaded back-to-back
 *      (independent loads), hits are offered to the operator in edge order; up to `pull_quick` edges.
 *   2. lanes still searching with >= `pull_coop` edges left hand their list to the whole warp: 32 lanes probe
 *      32 consecutive edges per step (coalesced), a ballot finds the hits.
 *   3. short remainders finish with stage-1 batches.
 */
constexpr int pull_batch = 4;
constexpr int pull_quick = 8;
constexpr int pull_coop = 32;

template <typename vertex_t, typename edge_t, typename weight_t, typename operator_t>
__global__ void __launch_bounds__(256)
    pull_step_staged_kernel(const graph::adjacency_t<vertex_t, edge_t, weight_t> A, operator_t op,
                            const unsigned* __restrict__ frontier_bits, unsigned* __restrict__ next_bits,
                            unsigned* __restrict__ visited, counter_t* counters) {
  const unsigned lane = b200::lane_id();
  const std::size_t n_words = (std::size_t(A.n) + 31) / 32;
  const std::size_t warps = (std::size_t(gridDim.x) * blockDim.x) >> 5;
  counter_t found_vertices = 0, found_edges = 0, scanned = 0, inspected = 0;

  // One batch of up to pull_batch edges of this lane's list. Returns true when the search is over.
  auto probe_batch = [&](vertex_t v, edge_t& cur, edge_t end, bool& found) -> bool {
    const int cnt = int(end - cur < edge_t(pull_batch) ? end - cur : edge_t(pull_batch));
    vertex_t u[pull_batch];
    unsigned hit = 0;
#pragma unroll
    for (int i = 0; i < pull_batch; ++i)
      if (i < cnt) u[i] = __ldg(A.indices + cur + i);
#pragma unroll
    for (int i = 0; i < pull_batch; ++i)
      if (i < cnt && ((__ldg(frontier_bits + (unsigned(u[i]) >> 5)) >> (unsigned(u[i]) & 31u)) & 1u)) hit |= 1u << i;
    inspected += counter_t(cnt);
#pragma unroll
    for (int i = 0; i < pull_batch; ++i)
      if (!found && (hit & (1u << i))) {
        weight_t weight = A.values ? __ldg(A.values + cur + i) : weight_t(1);
        vertex_t src = u[i], dst = v;
        edge_t edge = cur + i;
        found = op(src, dst, edge, weight);
      }
    cur += edge_t(cnt);
    return found || cur == end;
  };

  for (std::size_t w = (std::size_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5; w < n_words; w += warps) {
    const unsigned seen = visited[w];
    bool found = false, searching = false;
    edge_t cur = 0, end = 0, deg = 0;
    const vertex_t v = vertex_t(w * 32 + lane);
    if (seen != 0xffffffffu && !((seen >> lane) & 1u)) {
      cur = A.offsets[v];
      end = A.offsets[v + 1];
      deg = end - cur;
      ++scanned;
      searching = deg > 0;
    }
    // stage 1: quick batched probes
    for (int probed = 0; searching && probed < pull_quick; probed += pull_batch)
      searching = !probe_batch(v, cur, end, found);
    // stage 2: long remainders, one list at a time, whole warp
    unsigned long_lists = __ballot_sync(b200::full_mask, searching && end - cur >= edge_t(pull_coop));
    while (long_lists) {
      const int owner = __ffs(long_lists) - 1;
      long_lists &= long_lists - 1;
      const edge_t c0 = __shfl_sync(b200::full_mask, cur, owner), c1 = __shfl_sync(b200::full_mask, end, owner);
      const vertex_t target = vertex_t(w * 32 + owner);
      bool got = false;
      for (edge_t base = c0; base < c1 && !got; base += 32) {
        const edge_t e = base + edge_t(lane);
        vertex_t u = 0;
        bool hit = false;
        if (e < c1) {
          u = __ldg(A.indices + e);
          hit = (__ldg(frontier_bits + (unsigned(u) >> 5)) >> (unsigned(u) & 31u)) & 1u;
        }
        if (lane == 0) inspected += counter_t(c1 - base < 32 ? c1 - base : 32);
        unsigned hits = __ballot_sync(b200::full_mask, hit);
        while (hits && !got) {  // offer hits to the operator in edge order
          const int first = __ffs(hits) - 1;
          hits &= hits - 1;
          bool ok = false;
          if (int(lane) == first) {
            weight_t weight = A.values ? __ldg(A.values + e) : weight_t(1);
            vertex_t src = u, dst = target;
            edge_t edge = e;
            ok = op(src, dst, edge, weight);
          }
          got = __shfl_sync(b200::full_mask, ok, first);
        }
      }
      if (int(lane) == owner) {
        found = got;
        searching = false;
      }
    }
    // stage 3: short remainders
    while (searching) searching = !probe_batch(v, cur, end, found);

    const unsigned fresh = __ballot_sync(b200::full_mask, found);
    if (lane == 0) {
      next_bits[w] = fresh;
      if (fresh) visited[w] = seen | fresh;
      found_vertices += __popc(fresh);
    }
    if (found) found_edges += counter_t(deg);
  }
  found_edges = b200::warp_sum(found_edges);
  scanned = b200::warp_sum(scanned);
  inspected = b200::warp_sum(inspected);
  if (lane == 0) {
    if (found_vertices) atomicAdd(counters + scratch_t::out_count, found_vertices);
    if (found_edges) atomicAdd(counters + scratch_t::aux2, found_edges);
    if (scanned) atomicAdd(counters + scratch_t::aux0, scanned);      // unvisited vertices walked
    if (inspected) atomicAdd(counters + scratch_t::aux1, inspected);  // in-edges actually read
  }
}

/**
 * @brief Bottom-up level, latency-pipelined variant. Same contract as above. One lane per unvisited vertex,
 * plain serial walk, but the visited word and the row bounds of the warp's NEXT word are requested before
 * the current word is processed (`prefetch`), so two of the five dependent memory latencies of a trip
 * (visited word -> row bounds -> first neighbour -> frontier bit -> operator's atomic) overlap with the
 * previous trip. The kernel is latency-bound at full occupancy (64 warps/SM x 1 chain each), so chain
 * length and registers (<= 32 keeps 8 CTAs/SM resident) are what matter.
 */
template <bool prefetch, typename vertex_t, typename edge_t, typename weight_t, typename operator_t>
__global__ void __launch_bounds__(256, 8)
    pull_step_kernel(const graph::adjacency_t<vertex_t, edge_t, weight_t> A, operator_t op,
                     const unsigned* __restrict__ frontier_bits, unsigned* __restrict__ next_bits,
                     unsigned* __restrict__ visited, counter_t* counters) {
  const unsigned lane = b200::lane_id();
  const std::size_t n_words = (std::size_t(A.n) + 31) / 32;
  const std::size_t warps = (std::size_t(gridDim.x) * blockDim.x) >> 5;
  counter_t found_vertices = 0, found_edges = 0, scanned = 0, inspected = 0;
  std::size_t w = (std::size_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  unsigned seen = 0xffffffffu;
  edge_t beg = 0, end = 0;
  vertex_t head = -1;
  const bool hinted = A.head != nullptr;
  if (w < n_words) {
    seen = visited[w];
    const std::size_t v = w * 32 + lane;
    if (v < std::size_t(A.n)) {
      beg = A.offsets[v];
      end = A.offsets[v + 1];
      if (hinted) head = __ldg(A.head + v);
    }
  }
  while (w < n_words) {
    const std::size_t w_next = w + warps;
    unsigned seen_next = 0xffffffffu;
    edge_t beg_next = 0, end_next = 0;
    vertex_t head_next = -1;
    if (prefetch && w_next < n_words) {  // requests for the next trip overlap with this trip's walk
      seen_next = visited[w_next];
      const std::size_t v = w_next * 32 + lane;
      if (v < std::size_t(A.n)) {
        beg_next = A.offsets[v];
        end_next = A.offsets[v + 1];
        if (hinted) head_next = __ldg(A.head + v);
      }
    }
    bool found = false;
    if (!((seen >> lane) & 1u)) {
      ++scanned;
      const vertex_t v = vertex_t(w * 32 + lane);
      bool head_tried = false;
      if (head >= 0 && ((__ldg(frontier_bits + (unsigned(head) >> 5)) >> (unsigned(head) & 31u)) & 1u)) {
        // the hinted in-neighbour is in the frontier: no adjacency access at all for this vertex
        edge_t edge = __ldg(A.head_edge + v);
        weight_t weight = A.values ? __ldg(A.values + edge) : weight_t(1);
        vertex_t src = head, dst = v;
        found = op(src, dst, edge, weight);
        head_tried = true;
      }
      for (edge_t e = beg; e < end && !found; ++e) {
        ++inspected;
        const vertex_t u = __ldg(A.indices + e);
        if (head_tried && u == head) continue;  // already offered to the operator
        if ((__ldg(frontier_bits + (unsigned(u) >> 5)) >> (unsigned(u) & 31u)) & 1u) {
          weight_t weight = A.values ? __ldg(A.values + e) : weight_t(1);
          vertex_t src = u, dst = v;
          edge_t edge = e;
          if (op(src, dst, edge, weight)) {
            found = true;
            break;
          }
        }
      }
    }
    const unsigned fresh = __ballot_sync(b200::full_mask, found);
    if (lane == 0) {
      next_bits[w] = fresh;
      if (fresh) visited[w] = seen | fresh;
      found_vertices += __popc(fresh);
    }
    if (found) found_edges += counter_t(end - beg);
    if (!prefetch && w_next < n_words) {
      seen_next = visited[w_next];
      const std::size_t v = w_next * 32 + lane;
      if (seen_next != 0xffffffffu && v < std::size_t(A.n)) {
        beg_next = A.offsets[v];
        end_next = A.offsets[v + 1];
        if (hinted) head_next = __ldg(A.head + v);
      }
    }
    w = w_next;
    seen = seen_next;
    beg = beg_next;
    end = end_next;
    head = head_next;
  }
  found_edges = b200::warp_sum(found_edges);
  scanned = b200::warp_sum(scanned);
  inspected = b200::warp_sum(inspected);
  if (lane == 0) {
    if (found_vertices) atomicAdd(counters + scratch_t::out_count, found_vertices);
    if (found_edges) atomicAdd(counters + scratch_t::aux2, found_edges);
    if (scanned) atomicAdd(counters + scratch_t::aux0, scanned);
    if (inspected) atomicAdd(counters + scratch_t::aux1, inspected);
  }
}

/// Development knob: use the per-vertex head hints when the graph carries them (default on).
inline int& pull_hints_enabled() {
  static int enabled = 1;
  return enabled;
}

/// Development knob: which bottom-up kernel runs (0 serial walk, 1 + next-word prefetch, 2 staged/cooperative).
inline int& pull_variant() {
  static int variant = 1;
  return variant;
}

}  // namespace kernels
}  // namespace advance
}  // namespace operators
}  // namespace gunrock
