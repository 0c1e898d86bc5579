/**
 * @file capi_ops.cu
 * @brief C ABI: operator-level probes (advance / filter / uniquify with fixed test operators).
 */
#include "capi_dispatch.hxx"
#include "capi_frontier.hxx"

using namespace gunrock;
using gcuda::scratch_t;
using ess::borrowed_frontier_t;

namespace {

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

/// As advance_probe through operators::advance::execute_unique (fused advance + uniquify).
template <operators::load_balance_t lb, typename graph_t>
int advance_unique_probe(ess_context_t ctx, graph_t& G, const int32_t* d_frontier, int64_t frontier_size,
                         int32_t* d_out, int64_t out_capacity, int64_t* out_count, int32_t* d_edge_calls,
                         int32_t modulus) {
  using edge_t = typename graph_t::edge_type;
  borrowed_frontier_t<edge_t> in, out;
  in.ptr = const_cast<int32_t*>(d_frontier);
  in.count = in.cap = std::size_t(frontier_size);
  out.ptr = d_out;
  out.cap = d_out ? std::size_t(out_capacity) : 0;
  memory::device_array_t<edge_t> segments;
  frontier::frontier_t<int32_t, edge_t, frontier::frontier_kind_t::vertex_frontier, frontier::frontier_view_t::bitmap> seen;
  auto op = [d_edge_calls, modulus] __host__ __device__(int32_t const& src, int32_t const& nbr, edge_t const& e,
                                                        float const& w) -> bool {
    if (d_edge_calls) math::atomic::add(d_edge_calls + e, 1);
    return (long long)(src + nbr + (long long)e) % modulus != 0;
  };
  using namespace operators;
  // twice: the second pass must see an all-clear map again and reproduce the first one's set
  for (int pass = 0; pass < 2; ++pass)
    advance::execute_unique<lb, advance_direction_t::forward, advance_io_type_t::vertices>(G, op, &in, &out, segments,
                                                                                          seen, *ctx->ctx);
  if (out_count) *out_count = int64_t(out.count);
  if (d_out && out.ptr != d_out) {
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

/// math::atomic::{add,min,max,exch} probe: old[i] = atomic::<op>(cell, values[i]). serial != 0 runs all updates in
/// one thread (deterministic order: old[i] is the running result before values[i]); otherwise one thread per value.
template <int OP, typename T>
__global__ void atomic_probe_kernel(T* cell, const T* __restrict__ values, int64_t n, T* __restrict__ old, int serial) {
  auto apply = [&](int64_t i) {
    T before;
    if constexpr (OP == 0) before = math::atomic::add(cell, values[i]);
    if constexpr (OP == 1) before = math::atomic::min(cell, values[i]);
    if constexpr (OP == 2) before = math::atomic::max(cell, values[i]);
    if constexpr (OP == 3) before = math::atomic::exch(cell, values[i]);
    old[i] = before;
  };
  if (serial) {
    if (blockIdx.x == 0 && threadIdx.x == 0)
      for (int64_t i = 0; i < n; ++i) apply(i);
  } else {
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x) apply(i);
  }
}

template <typename T>
int atomic_probe(ess_context_t ctx, int op, T* cell, const T* values, int64_t n, T* old, int serial) {
  auto stream = ctx->single()->stream();
  const unsigned grid = serial ? 1u : unsigned(n < 256 * 592 ? (n + 255) / 256 : 592);
  switch (op) {
    case 0: atomic_probe_kernel<0><<<grid ? grid : 1, 256, 0, stream>>>(cell, values, n, old, serial); break;
    case 1: atomic_probe_kernel<1><<<grid ? grid : 1, 256, 0, stream>>>(cell, values, n, old, serial); break;
    case 2: atomic_probe_kernel<2><<<grid ? grid : 1, 256, 0, stream>>>(cell, values, n, old, serial); break;
    case 3: atomic_probe_kernel<3><<<grid ? grid : 1, 256, 0, stream>>>(cell, values, n, old, serial); break;
    default: return ess::fail("ess_atomic_probe: op must be 0 add, 1 min, 2 max, 3 exch");
  }
  error::check_last("ess_atomic_probe");
  ctx->single()->synchronize();
  return 0;
}

}  // namespace

extern "C" {

int ess_atomic_probe(ess_context_t ctx, int op, int is_float, void* d_cell, const void* d_values, int64_t n,
                     void* d_old, int serial) {
  ESS_TRY
  if (!ctx || !d_cell || (n > 0 && (!d_values || !d_old))) return ess::fail("ess_atomic_probe: null argument");
  if (is_float) return atomic_probe<float>(ctx, op, (float*)d_cell, (const float*)d_values, n, (float*)d_old, serial);
  return atomic_probe<int32_t>(ctx, op, (int32_t*)d_cell, (const int32_t*)d_values, n, (int32_t*)d_old, serial);
  ESS_CATCH
}


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

int ess_advance_unique_probe(ess_context_t ctx, ess_graph_t g, int lb, const int32_t* d_frontier, int64_t frontier_size,
                             int32_t* d_out, int64_t out_capacity, int64_t* out_count, int32_t* d_edge_calls,
                             int32_t modulus) {
  ESS_TRY
  if (!ctx || !g) return ess::fail("ess_advance_unique_probe: null argument");
  if (modulus <= 0) modulus = 1 << 30;
  return ess::with_load_balance(lb, [&](auto lbc) -> int {
    constexpr auto LB = decltype(lbc)::value;
    ESS_WITH_GRAPH(g, G, {
      return advance_unique_probe<LB>(ctx, G, d_frontier, frontier_size, d_out, out_capacity, out_count, d_edge_calls,
                                      modulus);
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

int ess_uniquify_probe(ess_context_t ctx, ess_graph_t g, const int32_t* d_in, int64_t size, int32_t* d_out,
                       int64_t* out_count) {
  ESS_TRY
  if (!ctx || !g || !d_out) return ess::fail("ess_uniquify_probe: null argument");
  using frontier_type = frontier::frontier_t<int32_t, int32_t>;
  frontier_type in{std::size_t(size), 1.0f};
  frontier_type out;
  if (size) cudaMemcpy(in.data(), d_in, std::size_t(size) * sizeof(int32_t), cudaMemcpyDeviceToDevice);
  operators::uniquify::execute<operators::uniquify_algorithm_t::unique>(&in, &out, std::size_t(g->n), *ctx->ctx);
  const std::size_t k = out.get_number_of_elements();
  if (k) cudaMemcpy(d_out, out.data(), k * sizeof(int32_t), cudaMemcpyDeviceToDevice);
  if (out_count) *out_count = int64_t(k);
  return 0;
  ESS_CATCH
}

}  // extern "C"
