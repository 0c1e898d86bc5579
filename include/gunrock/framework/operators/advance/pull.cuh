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

#include <type_traits>
#include <utility>
#include <gunrock/b200/warp.cuh>
#include <gunrock/framework/operators/advance/directional.cuh>
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
 * One warp per 32-vertex word, one lane per vertex. What ncu showed for the straightforward version
 * (profiles/r01_pull_step_full_raw.csv, r01_pull_sass_hotspots.txt): DRAM at 19 % of peak, 45 % issue
 * utilisation, and 70 % of the stall samples on two instructions — the consumer of the frontier-bit probe
 * (an L2 round trip) and the consumer of the operator's atomicMin. Neither bandwidth nor trip count is
 * the limit (a 4-words-per-warp version, a warp-cooperative version and a compacted-candidate-list version
 * were all slower or equal). So this version shortens the per-trip dependency chain:
 *   - software pipeline, depth 2: the visited word, row bounds and head hint of the warp's words t+1, t+2 and
 *     the frontier-bit probe of word t+1 are in flight while word t is processed;
 *   - head hint (graph::build::pull_hints): the highest-degree in-neighbour is read coalesced next to the row
 *     bounds and resolves ~85 % of the vertices of a Kronecker BFS without touching the adjacency list;
 *   - the operator's pull form (directional_operator_t) replaces the atomic round trip by a store;
 *   - 32-bit index arithmetic and counters (vertex ids are int32).
 * Semantics: the hinted in-neighbour is offered first, then the in-edges in order, until the operator
 * returns true.
 */
template <typename vertex_t, typename edge_t, typename weight_t, typename operator_t>
__global__ void __launch_bounds__(256, 6)
    pull_step_kernel(const graph::adjacency_t<vertex_t, edge_t, weight_t> A, operator_t op,
                     const unsigned* __restrict__ frontier_bits, unsigned* __restrict__ next_bits,
                     unsigned* __restrict__ visited, counter_t* counters) {
  // Multi-GPU use: A may hold only the rows of a contiguous range of global vertices; `visited` and `next_bits`
  // then point at the words of that range, frontier_bits and the neighbour ids stay global, and the operator
  // receives the LOCAL row index as destination.
  const unsigned lane = b200::lane_id();
  const unsigned n = unsigned(A.n);
  const unsigned n_words = (n + 31u) >> 5;
  const unsigned warps = (gridDim.x * blockDim.x) >> 5;
  const bool hinted = A.head != nullptr;
  unsigned found_vertices = 0, scanned = 0, inspected = 0;
  counter_t found_edges = 0;

  struct stage_t {
    unsigned seen;
    edge_t beg, end;
    vertex_t head;
  };
  auto load_seen = [&](unsigned w) -> unsigned { return w < n_words ? visited[w] : 0xffffffffu; };
  // row bounds + hint of this lane's vertex of word w — only for lanes whose visited bit is clear, so fully
  // or mostly visited words (isolated vertices, late levels) cost no row/hint traffic
  auto load_stage = [&](unsigned w, unsigned seen) {
    stage_t s{seen, 0, 0, vertex_t(-1)};
    if (!((seen >> lane) & 1u)) {  // padding lanes of the last word are marked visited by init_visited_kernel
      const unsigned v = (w << 5) + lane;
      s.beg = A.offsets[v];
      s.end = A.offsets[v + 1];
      if (hinted) s.head = __ldg(A.head + v);
    }
    return s;
  };
  auto probe_of = [&](const stage_t& s) -> unsigned {  // frontier word holding the head of this lane's vertex
    return s.head >= 0 ? __ldg(frontier_bits + (unsigned(s.head) >> 5)) : 0u;
  };

  // software pipeline over the warp's words w, w+warps, ...:  visited word 3 trips ahead, row bounds + hint
  // 2 trips ahead, frontier probe 1 trip ahead
  unsigned w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  stage_t cur = load_stage(w, load_seen(w));
  stage_t nxt = load_stage(w + warps, load_seen(w + warps));
  unsigned seen_far = load_seen(w + 2 * warps);
  unsigned probe = probe_of(cur);
  while (w < n_words) {
    const unsigned seen_farther = load_seen(w + 3 * warps);
    const stage_t far = load_stage(w + 2 * warps, seen_far);
    const unsigned probe_next = probe_of(nxt);
    bool found = false;
    if (!((cur.seen >> lane) & 1u)) {
      ++scanned;
      const vertex_t v = vertex_t((w << 5) + lane);
      bool head_tried = false;
      if (cur.head >= 0 && ((probe >> (unsigned(cur.head) & 31u)) & 1u)) {
        const edge_t edge = __ldg(A.head_edge + v);
        const weight_t weight = A.values ? __ldg(A.values + edge) : weight_t(1);
        found = call_pull(op, cur.head, v, edge, weight);
        head_tried = true;
      }
      for (edge_t e = cur.beg; e < cur.end && !found; ++e) {
        ++inspected;
        const vertex_t u = __ldg(A.indices + e);
        if (head_tried && u == cur.head) continue;  // already offered to the operator
        if ((__ldg(frontier_bits + (unsigned(u) >> 5)) >> (unsigned(u) & 31u)) & 1u) {
          const weight_t weight = A.values ? __ldg(A.values + e) : weight_t(1);
          found = call_pull(op, u, v, e, weight);
        }
      }
    }
    const unsigned fresh = __ballot_sync(b200::full_mask, found);
    if (lane == 0) {
      next_bits[w] = fresh;
      if (fresh) visited[w] = cur.seen | fresh;
      found_vertices += __popc(fresh);
    }
    if (found) found_edges += counter_t(cur.end - cur.beg);
    w += warps;
    cur = nxt;
    nxt = far;
    seen_far = seen_farther;
    probe = probe_next;
  }
  found_edges = b200::warp_sum(found_edges);
  scanned = b200::warp_sum(scanned);
  inspected = b200::warp_sum(inspected);
  if (lane == 0) {
    if (found_vertices) atomicAdd(counters + scratch_t::out_count, counter_t(found_vertices));
    if (found_edges) atomicAdd(counters + scratch_t::aux2, found_edges);
    if (scanned) atomicAdd(counters + scratch_t::aux0, counter_t(scanned));
    if (inspected) atomicAdd(counters + scratch_t::aux1, counter_t(inspected));
  }
}

/// Development knob: use the per-vertex head hints when the graph carries them (default on).
inline int& pull_hints_enabled() {
  static int enabled = 1;
  return enabled;
}

}  // namespace kernels
}  // namespace advance
}  // namespace operators
}  // namespace gunrock
