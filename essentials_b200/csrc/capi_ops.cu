/**
 * @file capi_ops.cu
 * @brief C ABI: operator-level probes (advance / filter with fixed test operators) and the per-level
 * kernels of the 1-D partitioned multi-GPU BFS.
 */
#include "capi_dispatch.hxx"

using namespace gunrock;
using gcuda::scratch_t;

namespace {

/// A borrowed device buffer dressed as a frontier: the operators only need data()/size/capacity.
template <typename edge_t>
struct borrowed_frontier_t {
  using type_t = int32_t;
  using offset_t = edge_t;
  int32_t* ptr = nullptr;
  std::size_t count = 0, cap = 0;
  memory::device_array_t<int32_t> own;  // used when the caller gave no (or too small a) buffer
  int32_t* data() { return ptr; }
  std::size_t get_number_of_elements() const { return count; }
  void set_number_of_elements(std::size_t c) { count = c; }
  std::size_t get_capacity() const { return cap; }
  void reserve(std::size_t n) {
    if (n <= cap) return;
    own.reserve(n, false);
    ptr = own.data();
    cap = n;
  }
};

template <operators::load_balance_t lb, typename graph_t>
int advance_probe(ess_context_t ctx, graph_t& G, int direction, const int32_t* d_frontier, int64_t frontier_size,
                  int32_t* d_out, int64_t out_capacity, int64_t* out_count, int32_t* d_edge_calls, int32_t modulus) {
  using edge_t = typename graph_t::edge_type;
  borrowed_frontier_t<edge_t> in, out;
  in.ptr = const_cast<int32_t*>(d_frontier);
  in.count = in.cap = std::size_t(frontier_size);
  out.ptr = d_out;
  out.cap = d_out ? std::size_t(out_capacity) : 0;
  memory::device_array_t<edge_t> segments;
  auto op = [d_edge_calls, modulus] __host__ __device__(int32_t const& src, int32_t const& nbr, edge_t const& e,
                                                        float const& w) -> bool {
    if (d_edge_calls) math::atomic::add(d_edge_calls + e, 1);
    return (long long)(src + nbr + (long long)e) % modulus != 0;
  };
  using namespace operators;
  if (direction == ESS_DIR_BACKWARD)
    advance::execute<lb, advance_direction_t::backward, advance_io_type_t::vertices, advance_io_type_t::vertices>(
        G, op, &in, &out, segments, *ctx->ctx);
  else
    advance::execute<lb, advance_direction_t::forward, advance_io_type_t::vertices, advance_io_type_t::vertices>(
        G, op, &in, &out, segments, *ctx->ctx);
  if (out_count) *out_count = int64_t(out.count);
  if (d_out && out.ptr != d_out) {  // the caller's buffer was too small: report the count, copy what fits
    std::size_t fit = out.count < std::size_t(out_capacity) ? out.count : std::size_t(out_capacity);
    cudaMemcpy(d_out, out.ptr, fit * sizeof(int32_t), cudaMemcpyDeviceToDevice);
  }
  return 0;
}

template <operators::filter_algorithm_t alg, typename graph_t>
int filter_probe(ess_context_t ctx, graph_t& G, const int32_t* d_in, int64_t size, int32_t* d_out, int64_t* out_count,
                 int32_t* d_calls, int32_t modulus) {
  using edge_t = typename graph_t::edge_type;
  borrowed_frontier_t<edge_t> in, out;
  in.ptr = const_cast<int32_t*>(d_in);
  in.count = in.cap = std::size_t(size);
  out.ptr = d_out;
  out.cap = std::size_t(size);
  auto op = [d_calls, modulus] __host__ __device__(int32_t const& v) -> bool {
    if (d_calls) math::atomic::add(d_calls + v, 1);
    return v % modulus != 0;
  };
  operators::filter::execute<alg>(G, op, &in, &out, *ctx->ctx);
  if (out_count) *out_count = int64_t(out.count);
  return 0;
}

// ---- multi-GPU BFS level kernels (bitmaps are global, rows are local) -------------------------------

/// Bottom-up over owned rows: owned vertex v (global id row_begin+v) not yet visited joins the candidates
/// when one of its in-neighbours is in the global frontier bitmap.
template <typename edge_t>
__global__ void __launch_bounds__(256)
    partition_pull_kernel(const graph::adjacency_t<int32_t, edge_t, float> A, long long row_begin,
                          const unsigned* __restrict__ frontier_bits, const unsigned* __restrict__ visited_bits,
                          unsigned* __restrict__ candidate_bits) {
  const unsigned lane = b200::lane_id();
  const std::size_t n_words = (std::size_t(A.n) + 31) / 32;
  const std::size_t first_word = std::size_t(row_begin) >> 5;
  const std::size_t warps = (std::size_t(gridDim.x) * blockDim.x) >> 5;
  for (std::size_t w = (std::size_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5; w < n_words; w += warps) {
    const unsigned seen = visited_bits[first_word + w];
    bool found = false;
    const std::size_t v = w * 32 + lane;
    if (seen != 0xffffffffu && !((seen >> lane) & 1u) && v < std::size_t(A.n)) {
      for (edge_t e = A.offsets[v], end = A.offsets[v + 1]; e < end; ++e) {
        const unsigned u = unsigned(__ldg(A.indices + e));
        if ((__ldg(frontier_bits + (u >> 5)) >> (u & 31u)) & 1u) {
          found = true;
          break;
        }
      }
    }
    const unsigned fresh = __ballot_sync(b200::full_mask, found);
    if (lane == 0) candidate_bits[first_word + w] = fresh;
  }
}

/// Owner side: fresh = candidate & ~visited over the owned words; depth, visited, next frontier, counters.
template <typename edge_t>
__global__ void __launch_bounds__(256)
    absorb_kernel(const edge_t* __restrict__ offsets, long long n_local, long long row_begin, int level,
                  const unsigned* __restrict__ candidate_bits, unsigned* __restrict__ visited_bits,
                  unsigned* __restrict__ next_bits, int* __restrict__ depth_local, b200::counter_t* counters) {
  const unsigned lane = b200::lane_id();
  const std::size_t n_words = (std::size_t(n_local) + 31) / 32;
  const std::size_t first_word = std::size_t(row_begin) >> 5;
  const std::size_t warps = (std::size_t(gridDim.x) * blockDim.x) >> 5;
  b200::counter_t vertices = 0, edges = 0;
  for (std::size_t w = (std::size_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5; w < n_words; w += warps) {
    const unsigned seen = visited_bits[first_word + w];
    const unsigned fresh = candidate_bits[first_word + w] & ~seen;
    if (lane == 0) {
      next_bits[first_word + w] = fresh;
      if (fresh) visited_bits[first_word + w] = seen | fresh;
      vertices += __popc(fresh);
    }
    const std::size_t v = w * 32 + lane;
    if (((fresh >> lane) & 1u) && v < std::size_t(n_local)) {
      depth_local[v] = level;
      edges += b200::counter_t(offsets[v + 1] - offsets[v]);
    }
  }
  edges = b200::warp_sum(edges);
  if (lane == 0) {
    if (vertices) atomicAdd(counters + scratch_t::out_count, vertices);
    if (edges) atomicAdd(counters + scratch_t::aux2, edges);
  }
}

template <typename graph_t>
int partition_step(ess_context_t ctx, graph_t& G, int64_t row_begin, int64_t n_global, int pull,
                   const uint32_t* d_frontier_bits, const uint32_t* d_visited_bits, uint32_t* d_candidate_bits) {
  using edge_t = typename graph_t::edge_type;
  auto* c = ctx->single();
  auto stream = c->stream();
  const auto A = graph::adjacency_of<false>(G);
  const std::size_t n_local = std::size_t(A.n);
  if (pull) {
    partition_pull_kernel<edge_t><<<gcuda::persistent_grid(*c, (n_local + 255) / 256, 8), 256, 0, stream>>>(
        A, (long long)row_begin, d_frontier_bits, d_visited_bits, d_candidate_bits);
    error::check_last("partition pull");
    return 0;
  }
  // push: owned slice of the frontier bitmap -> list of local row ids -> balanced advance whose operator
  // ORs unvisited neighbours into the (global) candidate bitmap.
  auto& scratch = c->scratch();
  borrowed_frontier_t<edge_t> in, out;
  in.reserve(n_local);
  scratch.zero(stream);
  const std::size_t words = (n_local + 31) / 32;
  frontier::kernels::gather_bits_kernel<<<gcuda::persistent_grid(*c, (words + 255) / 256, 8), 256, 0, stream>>>(
      d_frontier_bits + (std::size_t(row_begin) >> 5), words, in.data(), scratch.d + scratch_t::out_count);
  scratch.fetch(stream);
  in.count = std::size_t(scratch.h[scratch_t::out_count]);
  cudaMemsetAsync(d_candidate_bits, 0, ((std::size_t(n_global) + 31) / 32) * sizeof(uint32_t), stream);
  if (!in.count) return 0;
  memory::device_array_t<edge_t> segments;
  const unsigned* visited = d_visited_bits;
  unsigned* candidate = d_candidate_bits;
  auto op = [visited, candidate] __device__(int32_t const& src, int32_t const& nbr, edge_t const& e,
                                            float const& w) -> bool {
    const unsigned u = unsigned(nbr), bit = 1u << (u & 31u);
    if (!(visited[u >> 5] & bit) && !(candidate[u >> 5] & bit)) atomicOr(candidate + (u >> 5), bit);
    return false;
  };
  using namespace operators;
  advance::execute<load_balance_t::bucketing, advance_direction_t::forward, advance_io_type_t::vertices,
                   advance_io_type_t::none>(G, op, &in, &out, segments, *ctx->ctx);
  return 0;
}

}  // namespace

extern "C" {

int ess_advance_probe(ess_context_t ctx, ess_graph_t g, int lb, int direction, const int32_t* d_frontier,
                      int64_t frontier_size, int32_t* d_out, int64_t out_capacity, int64_t* out_count,
                      int32_t* d_edge_calls, int32_t modulus) {
  ESS_TRY
  if (!ctx || !g) return ess::fail("ess_advance_probe: null argument");
  if (direction == ESS_DIR_OPTIMIZED)
    return ess::fail("direction-optimized advance needs the enactor form (it keeps dense state in E)");
  if (direction == ESS_DIR_BACKWARD && !g->has_csc) return ess::fail("backward advance needs a CSC view");
  if (modulus <= 0) modulus = 1 << 30;
  return ess::with_load_balance(lb, [&](auto lbc) -> int {
    constexpr auto LB = decltype(lbc)::value;
    ESS_WITH_GRAPH(g, G, {
      return advance_probe<LB>(ctx, G, direction, d_frontier, frontier_size, d_out, out_capacity, out_count,
                               d_edge_calls, modulus);
    })
  });
  ESS_CATCH
}

int ess_filter_probe(ess_context_t ctx, ess_graph_t g, int alg, const int32_t* d_in, int64_t size, int32_t* d_out,
                     int64_t* out_count, int32_t* d_calls, int32_t modulus) {
  ESS_TRY
  if (!ctx || !g || !d_out) return ess::fail("ess_filter_probe: null argument");
  if (modulus <= 0) modulus = 1 << 30;
  using operators::filter_algorithm_t;
  ESS_WITH_GRAPH(g, G, {
    switch (alg) {
      case ESS_FILTER_REMOVE:
        return filter_probe<filter_algorithm_t::remove>(ctx, G, d_in, size, d_out, out_count, d_calls, modulus);
      case ESS_FILTER_PREDICATED:
        return filter_probe<filter_algorithm_t::predicated>(ctx, G, d_in, size, d_out, out_count, d_calls, modulus);
      case ESS_FILTER_COMPACT:
        return filter_probe<filter_algorithm_t::compact>(ctx, G, d_in, size, d_out, out_count, d_calls, modulus);
      case ESS_FILTER_BYPASS:
        return filter_probe<filter_algorithm_t::bypass>(ctx, G, d_in, size, d_out, out_count, d_calls, modulus);
      default:
        return ess::fail("Filter type not supported.");
    }
  })
  ESS_CATCH
}

int ess_bfs_partition_step(ess_context_t ctx, ess_graph_t g, int64_t row_begin, int64_t n_global, int pull,
                           const uint32_t* d_frontier_bits, const uint32_t* d_visited_bits,
                           uint32_t* d_candidate_bits) {
  ESS_TRY
  if (!ctx || !g) return ess::fail("ess_bfs_partition_step: null argument");
  if (row_begin % 32) return ess::fail("ess_bfs_partition_step: row_begin must be a multiple of 32");
  ESS_WITH_GRAPH(g, G, {
    return partition_step(ctx, G, row_begin, n_global, pull, d_frontier_bits, d_visited_bits, d_candidate_bits);
  })
  ESS_CATCH
}

int ess_bfs_absorb(ess_context_t ctx, ess_graph_t g, int64_t row_begin, int64_t n_global, int32_t level,
                   const uint32_t* d_candidate_bits, uint32_t* d_visited_bits, uint32_t* d_next_bits,
                   int32_t* d_depth_local, int64_t* fresh_vertices, int64_t* fresh_edges) {
  ESS_TRY
  if (!ctx || !g) return ess::fail("ess_bfs_absorb: null argument");
  auto* c = ctx->single();
  auto stream = c->stream();
  auto& scratch = c->scratch();
  scratch.zero(stream);
  const std::size_t n_local = std::size_t(g->n);
  const unsigned grid = gcuda::persistent_grid(*c, (n_local + 255) / 256, 8);
  if (g->offset_bits == 64)
    absorb_kernel<int64_t><<<grid, 256, 0, stream>>>(g->g64.get_row_offsets(), (long long)n_local, (long long)row_begin,
                                                      level, d_candidate_bits, d_visited_bits, d_next_bits,
                                                      d_depth_local, scratch.d);
  else
    absorb_kernel<int32_t><<<grid, 256, 0, stream>>>(g->g32.get_row_offsets(), (long long)n_local, (long long)row_begin,
                                                      level, d_candidate_bits, d_visited_bits, d_next_bits,
                                                      d_depth_local, scratch.d);
  error::check_last("absorb");
  scratch.fetch(stream);
  if (fresh_vertices) *fresh_vertices = int64_t(scratch.h[scratch_t::out_count]);
  if (fresh_edges) *fresh_edges = int64_t(scratch.h[scratch_t::aux2]);
  (void)n_global;
  return 0;
  ESS_CATCH
}

}  // extern "C"
