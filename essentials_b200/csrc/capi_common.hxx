/**
 * @file capi_common.hxx
 * @brief Handle types and error plumbing shared by the C-ABI translation units (include/essentials_b200.h).
 * Each entry point instantiates the header-template operator API (include/gunrock/) for the fixed types the
 * reference's drivers use — vertex int32, weight float, edge int32 or int64 — and converts C++ exceptions
 * to return codes.
 */
#pragma once

#include <memory>
#include <string>

#include <essentials_b200.h>
#include <gunrock/algorithms/algorithms.hxx>

namespace ess {

using namespace gunrock;
using vertex_t = int32_t;
using weight_t = float;

template <typename edge_t>
using graph_of = graph::graph_t<memory::memory_space_t::device, vertex_t, edge_t, weight_t,
                                graph::graph_csr_t<vertex_t, edge_t, weight_t>,
                                graph::graph_csc_t<vertex_t, edge_t, weight_t>, graph::empty_coo_t>;

std::string& last_error();
int fail(const std::exception& e);
int fail(const char* message, int code = 999);
/// ess_tune("dist_peer_exchange"): 1 = the partitioned BFS exchanges its bitmaps with its own peer-memory
/// kernels when the IPC window could be mapped (default), 0 = NCCL collectives.
int& dist_peer_exchange();
int& dist_peer_timeout_ms();  ///< ess_tune("dist_peer_timeout_ms"): how long a rank waits for a peer's level data (4000)
int& dist_trace();  ///< ess_tune("dist_trace"): per-phase event timing of ess_dist_bfs on stderr (rank 0)

}  // namespace ess

struct ess_context_s {
  std::shared_ptr<gunrock::gcuda::multi_context_t> ctx;
  gunrock::gcuda::standard_context_t* single() { return ctx->get_context(0); }
};

struct ess_graph_s {
  int offset_bits = 32;
  bool has_csc = false;
  int64_t n = 0, m = 0;
  ess::graph_of<int32_t> g32;
  ess::graph_of<int64_t> g64;
  // bottom-up hints owned by the handle (graph::build::pull_hints)
  gunrock::memory::device_array_t<int32_t> hint_head;
  gunrock::memory::device_array_t<int32_t> hint_edge32;
  gunrock::memory::device_array_t<int64_t> hint_edge64;
  gunrock::memory::device_array_t<unsigned> hint_isolated;
  // device copies owned by the handle (ess_graph_create_from_host); empty when the caller owns the arrays
  gunrock::memory::device_array_t<unsigned char> own_offsets;
  gunrock::memory::device_array_t<int32_t> own_indices;
  gunrock::memory::device_array_t<float> own_values;
};

#define ESS_TRY try {
#define ESS_CATCH                                   \
  }                                                 \
  catch (const gunrock::error::exception_t& e) {    \
    ess::last_error() = e.what();                   \
    int code = int(e.status());                     \
    return code ? code : 999;                       \
  }                                                 \
  catch (const std::exception& e) {                 \
    return ess::fail(e);                            \
  }                                                 \
  catch (...) {                                     \
    return ess::fail("unknown exception");          \
  }

/// Calls `body(G)` with the graph_t of the handle's edge width.
#define ESS_WITH_GRAPH(handle, G, ...)  \
  if ((handle)->offset_bits == 64) {    \
    auto& G = (handle)->g64;            \
    __VA_ARGS__                         \
  } else {                              \
    auto& G = (handle)->g32;            \
    __VA_ARGS__                         \
  }

namespace ess {
inline void fill_info(ess_run_info* info, float ms, int iterations, int pull = 0, int push = 0) {
  if (!info) return;
  info->enact_ms = ms;
  info->iterations = iterations;
  info->pull_steps = pull;
  info->push_steps = push;
  for (auto& r : info->reserved) r = 0;
}
}  // namespace ess
