/**
 * @file problem.hxx
 * @brief problem_t: algorithm state bound to a graph view and a context; init()/reset() are the two
 * virtuals an algorithm fills in. Same members and accessors as the reference
 * (include/gunrock/framework/problem.hxx:29-59): get_graph() returns the view BY VALUE.
 */
#pragma once

#include <memory>
#include <gunrock/cuda/cuda.hxx>
#include <gunrock/graph/graph.hxx>

namespace gunrock {

template <typename graph_t>
struct problem_t {
  using vertex_t = typename graph_t::vertex_type;
  using edge_t = typename graph_t::edge_type;
  using weight_t = typename graph_t::weight_type;

  graph_t graph_slice;
  std::shared_ptr<gcuda::multi_context_t> context;

  problem_t() : graph_slice() {}
  problem_t(graph_t& G, std::shared_ptr<gcuda::multi_context_t> _context) : graph_slice(G), context(_context) {}
  virtual ~problem_t() = default;

  auto get_graph() { return graph_slice; }
  auto get_multi_context() { return context; }
  auto get_single_context(gcuda::device_id_t device = 0) { return context->get_context(device); }

  virtual void init() = 0;
  virtual void reset() = 0;

  problem_t(const problem_t&) = delete;
  problem_t& operator=(const problem_t&) = delete;
};

}  // namespace gunrock
