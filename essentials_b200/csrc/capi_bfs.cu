/** @file capi_bfs.cu  C ABI: ess_bfs (gunrock::bfs::run, reference include/gunrock/algorithms/bfs.hxx:151-176). */
#include "capi_dispatch.hxx"
#include <gunrock/algorithms/bfs.hxx>

using namespace gunrock;

namespace {
template <operators::load_balance_t lb, operators::advance_direction_t dir, typename graph_t>
int run_bfs(ess_context_t ctx, graph_t& G, int32_t source, int32_t* d_depth, float alpha, float beta,
            ess_run_info* info) {
  enactor_properties_t props;
  if (alpha > 0) props.direction_alpha = alpha;
  if (beta > 0) props.direction_beta = beta;
  int pulls = 0, iters = 0;
  int32_t src = source;
  long long stats[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  float ms = bfs::run<lb, dir>(G, src, d_depth, (int32_t*)nullptr, ctx->ctx, props, &pulls, &iters, stats);
  ess::fill_info(info, ms, iters, pulls, iters - pulls);
  if (info)
    for (int i = 0; i < 8; ++i) info->reserved[i] = stats[i];
  return 0;
}
}  // namespace

extern "C" int ess_bfs(ess_context_t ctx, ess_graph_t g, int32_t source, int32_t* d_depth, int lb, int direction,
                       float alpha, float beta, ess_run_info* info) {
  ESS_TRY
  if (!ctx || !g || !d_depth) return ess::fail("ess_bfs: null argument");
  if (source < 0 || source >= g->n) return ess::fail("ess_bfs: source out of range");
  if (direction == ESS_DIR_OPTIMIZED && !g->has_csc)
    return ess::fail("CSR and CSC sparse-matrix representations required for direction-optimized advance.");
  if (direction != ESS_DIR_FORWARD && direction != ESS_DIR_OPTIMIZED)
    return ess::fail("ess_bfs: direction must be FORWARD or OPTIMIZED");
  return ess::with_load_balance(lb, [&](auto lbc) -> int {
    constexpr auto LB = decltype(lbc)::value;
    ESS_WITH_GRAPH(g, G, {
      if (direction == ESS_DIR_OPTIMIZED)
        return run_bfs<LB, operators::advance_direction_t::optimized>(ctx, G, source, d_depth, alpha, beta, info);
      return run_bfs<LB, operators::advance_direction_t::forward>(ctx, G, source, d_depth, alpha, beta, info);
    })
  });
  ESS_CATCH
}
