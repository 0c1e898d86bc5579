/**
 * @file enactor.hxx
 * @brief enactor_t: the bulk-synchronous driver. prepare_frontier() -> while(!is_converged()) { loop(); ++iteration }
 * -> finalize(), timed with the context's event timer; double-buffered frontiers swapped by selector.
 *
 * Contract kept from the reference (include/gunrock/framework/enactor.hxx:83-310): public members
 * properties, context, problem, frontiers, scanned_work_domain, active_frontier, inactive_frontier,
 * buffer_selector, iteration; the accessors; enact()'s exact call order (:243-254); pointer-stable
 * frontiers[] with swap-by-selector (:229-235) — k-core watches a fixed buffer pointer across swaps
 * (reference kcore.hxx:123,158).
 *
 * B200 changes: (1) frontier buffers are NOT pre-reserved at 1.5 x max(E,V) elements
 * (reference :181-191 — 25.8 GB at scale-26). Advance compacts its output, so buffers start at
 * n elements and grow on demand through the operators' overflow protocol. (2) `scanned_work_domain` is a
 * plain device array (merge-path work offsets). (3) `direction` holds the dense state of
 * direction-optimised advance (visited bitmap, dense frontiers, Beamer counters). (4) enact() brackets
 * the loop on the context's stream, not the legacy default stream.
 */
#pragma once

#include <memory>
#include <vector>

#include <gunrock/cuda/cuda.hxx>
#include <gunrock/framework/frontier/frontier.hxx>
#include <gunrock/framework/problem.hxx>

namespace gunrock {

struct enactor_properties_t {
  float frontier_sizing_factor{1.5f};
  std::size_t number_of_frontier_buffers{2};
  bool self_manage_frontiers{false};
  /// Beamer push->pull threshold: go bottom-up when frontier edges > unexplored edges / alpha.
  float direction_alpha{14.f};
  /// Beamer pull->push threshold: go top-down when frontier vertices < n / beta.
  float direction_beta{24.f};
  /// Initial elements reserved per frontier buffer (0 = number of vertices).
  std::size_t initial_frontier_capacity{0};
  enactor_properties_t() = default;
};

/**
 * @brief Dense side state of advance<..., advance_direction_t::optimized> (push/pull switching).
 * Owned by the enactor so it persists across loop() calls. All bitmaps are n bits.
 */
template <typename vertex_t, typename edge_t>
struct direction_state_t {
  using bits_t = frontier::frontier_t<vertex_t, edge_t, frontier::frontier_kind_t::vertex_frontier,
                                      frontier::frontier_view_t::bitmap>;
  bits_t visited;        ///< vertices that may no longer join a frontier
  bits_t dense[2];       ///< dense frontier, double buffered
  int dense_selector = 0;
  bool initialised = false;
  bool frontier_is_dense = false;  ///< the live frontier is dense[dense_selector]; the sparse buffer is stale
  bool pulling = false;
  long long frontier_edges = 0;    ///< sum of degrees of the live frontier (m_f)
  long long unexplored_edges = 0;  ///< sum of degrees of unvisited vertices (m_u)
  long long frontier_vertices = 0;
  long long previous_frontier_vertices = 0;
  int pull_steps = 0, push_steps = 0;  ///< statistics for the harness
  long long pull_vertices_scanned = 0;  ///< unvisited vertices walked by bottom-up levels
  long long pull_edges_inspected = 0;   ///< in-edges read by bottom-up levels (early exit counted)
  long long push_edges_expanded = 0;    ///< out-edges expanded by top-down levels
  long long push_vertices_expanded = 0; ///< frontier vertices expanded by top-down levels
  long long pull_hint_misses = 0;       ///< bottom-up: probed vertices whose hint missed (adjacency walked)
  long long pull_vertices_found = 0;    ///< vertices adopted by bottom-up levels
  long long push_vertices_found = 0;    ///< vertices claimed by top-down levels
  /// Size the three bitmaps for n vertices (idempotent; call before enact() to keep it out of the timed loop).
  void allocate(std::size_t n, gcuda::stream_t stream = 0) {
    if (visited.get_universe() == n) return;
    visited.resize(n, stream);
    dense[0].resize(n, stream);
    dense[1].resize(n, stream);
  }
  void reset() {
    initialised = false;
    frontier_is_dense = false;
    pulling = false;
    dense_selector = 0;
    frontier_edges = unexplored_edges = frontier_vertices = previous_frontier_vertices = 0;
    pull_steps = push_steps = 0;
    pull_vertices_scanned = pull_edges_inspected = push_edges_expanded = push_vertices_expanded = 0;
    pull_hint_misses = pull_vertices_found = push_vertices_found = 0;
  }
};

template <typename algorithm_problem_t,
          frontier::frontier_kind_t frontier_kind = frontier::frontier_kind_t::vertex_frontier,
          frontier::frontier_view_t frontier_view = frontier::frontier_view_t::vector>
struct enactor_t {
  using vertex_t = typename algorithm_problem_t::vertex_t;
  using edge_t = typename algorithm_problem_t::edge_t;
  using frontier_t = frontier::frontier_t<vertex_t, edge_t, frontier_kind>;

  enactor_properties_t properties;
  std::shared_ptr<gcuda::multi_context_t> context;
  algorithm_problem_t* problem;
  std::vector<frontier_t> frontiers;
  memory::device_array_t<edge_t> scanned_work_domain;
  frontier_t* active_frontier;
  frontier_t* inactive_frontier;
  int buffer_selector;
  int iteration;
  direction_state_t<vertex_t, edge_t> direction;
  /// "already emitted in this call" bitmap of operators::advance::execute_unique (all clear between calls).
  typename direction_state_t<vertex_t, edge_t>::bits_t unique_seen;
  /// 1-D partitioned runs (context->partition set): frontier size summed over all ranks, maintained by
  /// operators::exchange::execute and by enact() after prepare_frontier(); -1 = not partitioned.
  long long global_frontier_size = -1;
  memory::device_array_t<unsigned long long> exchange_counts;  ///< operators::exchange scratch
  memory::device_array_t<uint2> exchange_send, exchange_recv;

  enactor_t(const enactor_t&) = delete;
  enactor_t& operator=(const enactor_t&) = delete;

  enactor_t(algorithm_problem_t* _problem, std::shared_ptr<gcuda::multi_context_t> _context,
            enactor_properties_t _properties = enactor_properties_t())
      : properties(_properties),
        context(_context),
        problem(_problem),
        frontiers(_properties.number_of_frontier_buffers),
        active_frontier(&frontiers[0]),
        inactive_frontier(&frontiers[_properties.number_of_frontier_buffers > 1 ? 1 : 0]),
        buffer_selector(0),
        iteration(0) {
    if (!properties.self_manage_frontiers) {
      auto g = problem->get_graph();
      scanned_work_domain.resize(std::size_t(g.get_number_of_vertices()) + 1);  // as the reference (:168)
      std::size_t initial = properties.initial_frontier_capacity ? properties.initial_frontier_capacity
                                                                 : std::size_t(g.get_number_of_vertices());
      for (auto& buffer : frontiers) {
        buffer.set_resizing_factor(properties.frontier_sizing_factor);
        buffer.reserve(initial);
      }
    }
  }
  virtual ~enactor_t() = default;

  algorithm_problem_t* get_problem() { return problem; }
  frontier_t* get_input_frontier() { return active_frontier; }
  frontier_t* get_output_frontier() { return inactive_frontier; }
  enactor_t* get_enactor() { return this; }

  void swap_frontier_buffers() {
    buffer_selector ^= 1;
    active_frontier = &frontiers[buffer_selector];
    inactive_frontier = &frontiers[buffer_selector ^ 1];
  }

  /// Runs the algorithm to convergence; returns the milliseconds spent between the first is_converged()
  /// and the end of finalize() as measured by CUDA events on the context's stream.
  float enact() {
    auto single_context = context->get_context(0);
    prepare_frontier(get_input_frontier(), *context);
    if (context->partition && context->partition->world > 1) {  // partitioned: convergence is a global property
      memory::device_array_t<long long> size(1);
      long long mine = (long long)get_input_frontier()->get_number_of_elements();
      cudaMemcpyAsync(size.data(), &mine, sizeof(mine), cudaMemcpyHostToDevice, single_context->stream());
      context->partition->all_reduce_sum(size.data(), 1, single_context->stream());
      cudaMemcpyAsync(&mine, size.data(), sizeof(mine), cudaMemcpyDeviceToHost, single_context->stream());
      single_context->synchronize();
      global_frontier_size = mine;
    }
    auto& timer = single_context->timer();
    timer.begin();
    while (!is_converged(*context)) {
      loop(*context);
      ++iteration;
    }
    finalize(*context);
    return timer.end();
  }

  virtual void loop(gcuda::multi_context_t& context) = 0;
  virtual void prepare_frontier(frontier_t* f, gcuda::multi_context_t& context) {}
  virtual bool is_converged(gcuda::multi_context_t& context) {
    if (global_frontier_size >= 0) return global_frontier_size == 0;  // partitioned run: empty on every rank
    return active_frontier->is_empty();
  }
  virtual void finalize(gcuda::multi_context_t& context) {}
};

}  // namespace gunrock
