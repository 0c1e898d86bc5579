/**
 * @file uniquify.hxx
 * @brief operators::uniquify::execute — remove duplicate (and invalid) elements from a vertex frontier.
 * Enactor-form signature as the reference (framework/operators/uniquify/uniquify.hxx:44-72, which radix-sorts
 * the frontier and calls thrust::unique / unique_copy). A vertex frontier lives in [0, n), so the same
 * result — ascending, duplicate-free — is produced by a sparse -> dense -> sparse round trip over an n-bit
 * map (frontier::convert): O(|F| + n/8) bytes, no sort. `best_effort` / `percent` are accepted and ignored
 * (the exact answer is cheaper than the reference's approximate one).
 */
#pragma once

#include <gunrock/cuda/cuda.hxx>
#include <gunrock/error.hxx>
#include <gunrock/framework/frontier/frontier.hxx>
#include <gunrock/framework/operators/configs.hxx>

namespace gunrock {
namespace operators {
namespace uniquify {

/// Explicit form: `universe` = number of vertices (ids must lie in [0, universe)).
template <uniquify_algorithm_t type = uniquify_algorithm_t::unique, typename frontier_t>
void execute(frontier_t* input, frontier_t* output, std::size_t universe, gcuda::multi_context_t& context) {
  error::throw_if_exception(context.size() != 1, "`context.size() != 1` not supported");
  using vertex_t = typename frontier_t::vertex_type;
  using edge_t = typename frontier_t::edge_type;
  auto* ctx = context.get_context(0);
  // sized on the context's (non-blocking) stream: a clear on the legacy default stream would not be ordered with
  // the scatter convert() enqueues and could wipe bits that were just set
  frontier::frontier_t<vertex_t, edge_t, frontier::frontier_kind_t::vertex_frontier, frontier::frontier_view_t::bitmap>
      dense;
  dense.resize(universe, ctx->stream());
  frontier::convert(*input, dense, ctx->stream());
  frontier::convert(dense, *output, *ctx);
}

/// Enactor form (reference uniquify.hxx:44-72).
template <uniquify_algorithm_t type = uniquify_algorithm_t::unique, typename enactor_type>
void execute(enactor_type* E, gcuda::multi_context_t& context, bool best_effort_uniquification = false,
             const float uniquification_percent = 100, bool swap_buffers = true) {
  error::throw_if_exception(!best_effort_uniquification &&
                                (uniquification_percent < 0 || uniquification_percent > 100),
                            "Uniquification percentage must be a +ve float between 0 and 100.");
  const std::size_t n = std::size_t(E->get_problem()->get_graph().get_number_of_vertices());
  execute<type>(E->get_input_frontier(), E->get_output_frontier(), n, context);
  if (swap_buffers) E->swap_frontier_buffers();
}

}  // namespace uniquify
}  // namespace operators
}  // namespace gunrock
