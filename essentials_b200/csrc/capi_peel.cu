/** @file capi_peel.cu  C ABI: ess_kcore, ess_color (reference include/gunrock/algorithms/kcore.hxx:202-222, color.hxx:155-180). */
#include "capi_dispatch.hxx"
#include <gunrock/algorithms/kcore.hxx>
#include <gunrock/algorithms/color.hxx>

using namespace gunrock;

extern "C" int ess_kcore(ess_context_t ctx, ess_graph_t g, int32_t* d_k_cores, int lb, ess_run_info* info) {
  ESS_TRY
  if (!ctx || !g || !d_k_cores) return ess::fail("ess_kcore: null argument");
  return ess::with_load_balance(lb, [&](auto lbc) -> int {
    constexpr auto LB = decltype(lbc)::value;
    ESS_WITH_GRAPH(g, G, {
      float ms = kcore::run<LB>(G, d_k_cores, ctx->ctx);
      ess::fill_info(info, ms, 0);
      return 0;
    })
  });
  ESS_CATCH
}

extern "C" int ess_color(ess_context_t ctx, ess_graph_t g, int32_t* d_colors, ess_run_info* info) {
  ESS_TRY
  if (!ctx || !g || !d_colors) return ess::fail("ess_color: null argument");
  ESS_WITH_GRAPH(g, G, {
    int iters = 0;
    float ms = color::run(G, d_colors, ctx->ctx, &iters);
    ess::fill_info(info, ms, iters);
    return 0;
  })
  ESS_CATCH
}
