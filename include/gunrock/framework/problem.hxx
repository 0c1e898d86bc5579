/**
 * @file problem.hxx
 * @brief problem_t: algorithm state bound to a graph view and a context; init()/reset() are the two
 * virtuals an algorithm fills in. Member names and accessors are the reference's
 * (include/gunrock/framework/problem.hxx:29-59) because algorithm headers reach into them:
 * `graph_slice`, `context`, get_graph() (the view BY VALUE, so it can be captured into kernels),
 * get_multi_context(), get_single_context(device).
 */
#pragma once

#include <memory>
#include <gunrock/cuda/cuda.hxx>
#include <gunrock/graph/graph.hxx>
#include <gunrock/util/type_limits.hxx>

namespace gunrock {

template <typename graph_t>
struct problem_t {
  using graph_type = graph_t;
  using weight_t = typename graph_t::weight_type;
  using edge_t = typename graph_t::edge_type;
  using vertex_t = typename graph_t::vertex_type;
  using context_ptr_t = std::shared_ptr<gcuda::multi_context_t>;

  // a problem owns device state: it is neither copied nor sliced
  problem_t(const problem_t&) = delete;
  problem_t& operator=(const problem_t&) = delete;
  virtual ~problem_t() = default;

  problem_t() : graph_slice() {}
  problem_t(graph_t& G, context_ptr_t _context) : graph_slice(G), context(std::move(_context)) {}

  /// Algorithm hooks: allocate once / restore the start state before every enact().
  virtual void init() = 0;
  virtual void reset() = 0;

  auto get_graph() { return graph_slice; }
  /// Length of per-vertex label arrays: the whole graph's vertex count, also when this rank holds a row range only
  /// (1-D partitioned runs index labels by global id, see graph_properties_t::row_offset).
  std::size_t label_count() const {
    const auto& props = graph_slice.get_properties();
    return props.global_vertices ? std::size_t(props.global_vertices) : std::size_t(graph_slice.get_number_of_vertices());
  }
  /// Local row id of a global vertex on this rank, or invalid when another rank owns it (identity when not partitioned).
  vertex_t local_row(vertex_t v) const {
    const auto& props = graph_slice.get_properties();
    if (!props.global_vertices) return v;
    const long long local = (long long)v - props.row_offset;
    return local >= 0 && local < (long long)graph_slice.get_number_of_vertices()
               ? vertex_t(local) : gunrock::numeric_limits<vertex_t>::invalid();
  }
  auto get_multi_context() { return context; }
  auto get_single_context(gcuda::device_id_t device = 0) { return context->get_context(device); }
  /// Stream every kernel of this problem's algorithm is enqueued on.
  auto stream(gcuda::device_id_t device = 0) { return context->get_context(device)->stream(); }

  graph_t graph_slice;
  context_ptr_t context;
};

}  // namespace gunrock
