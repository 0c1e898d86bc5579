/** @file capi_pr.cu  C ABI: ess_pagerank, ess_ppr (reference include/gunrock/algorithms/pr.hxx:183-216, ppr.hxx:150-179). */
#include "capi_dispatch.hxx"
#include <gunrock/algorithms/pr.hxx>
#include <gunrock/algorithms/ppr.hxx>

using namespace gunrock;

extern "C" int ess_pagerank(ess_context_t ctx, ess_graph_t g, float alpha, float tol, int max_iterations, float* d_p,
                            int lb, int pull, ess_run_info* info) {
  ESS_TRY
  if (!ctx || !g || !d_p) return ess::fail("ess_pagerank: null argument");
  if (pull && !g->has_csc) return ess::fail("ess_pagerank: pull needs a CSC view");
  if (max_iterations <= 0) max_iterations = 1000;
  if (pull) {
    ESS_WITH_GRAPH(g, G, {
      int iters = 0;
      float ms = pr::run<operators::load_balance_t::block_mapped, true>(G, alpha, tol, d_p, ctx->ctx, &iters,
                                                                         max_iterations);
      ess::fill_info(info, ms, iters);
      return 0;
    })
  }
  return ess::with_load_balance(lb, [&](auto lbc) -> int {
    constexpr auto LB = decltype(lbc)::value;
    ESS_WITH_GRAPH(g, G, {
      int iters = 0;
      float ms = pr::run<LB, false>(G, alpha, tol, d_p, ctx->ctx, &iters, max_iterations);
      ess::fill_info(info, ms, iters);
      return 0;
    })
  });
  ESS_CATCH
}

extern "C" int ess_ppr(ess_context_t ctx, ess_graph_t g, int32_t seed, float alpha, float epsilon, float* d_p, int lb,
                       ess_run_info* info) {
  ESS_TRY
  if (!ctx || !g || !d_p) return ess::fail("ess_ppr: null argument");
  if (seed < 0 || seed >= g->n) return ess::fail("ess_ppr: seed out of range");
  return ess::with_load_balance(lb, [&](auto lbc) -> int {
    constexpr auto LB = decltype(lbc)::value;
    ESS_WITH_GRAPH(g, G, {
      int32_t s = seed;
      float a = alpha, e = epsilon;
      float ms = ppr::run<LB>(G, s, d_p, a, e, ctx->ctx);
      ess::fill_info(info, ms, 0);
      return 0;
    })
  });
  ESS_CATCH
}
