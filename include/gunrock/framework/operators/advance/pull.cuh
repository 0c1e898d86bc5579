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
 */
template <typename vertex_t, typename edge_t, typename weight_t, typename operator_t>
__global__ void __launch_bounds__(256)
    pull_step_kernel(const graph::adjacency_t<vertex_t, edge_t, weight_t> A, operator_t op,
                     const unsigned* __restrict__ frontier_bits, unsigned* __restrict__ next_bits,
                     unsigned* __restrict__ visited, counter_t* counters) {
  const unsigned lane = b200::lane_id();
  const std::size_t n_words = (std::size_t(A.n) + 31) / 32;
  const std::size_t warps = (std::size_t(gridDim.x) * blockDim.x) >> 5;
  counter_t found_vertices = 0, found_edges = 0, scanned = 0, inspected = 0;
  for (std::size_t w = (std::size_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5; w < n_words; w += warps) {
    const unsigned seen = visited[w];
    bool found = false;
    edge_t deg = 0;
    if (seen != 0xffffffffu && !((seen >> lane) & 1u)) {
      const vertex_t v = vertex_t(w * 32 + lane);
      const edge_t beg = A.offsets[v], end = A.offsets[v + 1];
      deg = end - beg;
      ++scanned;
      for (edge_t e = beg; e < end; ++e) {
        ++inspected;
        const vertex_t u = __ldg(A.indices + e);
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

}  // namespace kernels
}  // namespace advance
}  // namespace operators
}  // namespace gunrock
