/**
 * @file pr.hxx
 * @brief PageRank client of the frontier operators (whole-graph advance, no output frontier).
 *
 * Follows the reference's iteration (include/gunrock/algorithms/pr.hxx): reset :64-92 (p = 1/n, plast = 0,
 * iweights[v] = alpha / Σw(v), 0 when v has no out-edges); loop :106-153 (plast = p; dangling mass
 * dsum = Σ_{iweights==0} alpha·p; p = (1-alpha+dsum)/n; advance<graph, none> with
 * atomic::add(p+dst, plast[src]·iweights[src]·w)); is_converged :155-178 (after the first iteration, stop
 * when max|p - plast| < tol). Enactor runs with self_manage_frontiers like the reference (:210-211).
 *
 * The n-length passes are fused: one kernel copies p→plast and reduces the dangling mass into a device
 * scalar (double), the next one reads that scalar to refill p — no host round trip between them
 * (the reference blocks on a thrust::transform_reduce, :128-131). `pull = true` replaces the atomic
 * scatter by a CSC gather with a per-vertex double accumulator (deterministic, no atomics): needs
 * a CSC view.
 */
#pragma once

#include <gunrock/algorithms/algorithms.hxx>

namespace gunrock {
namespace pr {

template <typename weight_t>
struct param_t {
  weight_t alpha;
  weight_t tol;
  int max_iterations;
  param_t(weight_t _alpha, weight_t _tol, int _max_iterations = 1000)
      : alpha(_alpha), tol(_tol), max_iterations(_max_iterations) {}
};

template <typename weight_t>
struct result_t {
  weight_t* p;
  result_t(weight_t* _p) : p(_p) {}
};

namespace kernels {
using b200::counter_t;

template <typename graph_t, typename weight_t>
__global__ void __launch_bounds__(256) inverse_weights_kernel(graph_t G, weight_t alpha, weight_t* iweights) {
  using edge_t = typename graph_t::edge_type;
  using vertex_t = typename graph_t::vertex_type;
  const std::size_t n = std::size_t(G.get_number_of_vertices());
  for (std::size_t v = std::size_t(blockIdx.x) * blockDim.x + threadIdx.x; v < n;
       v += std::size_t(gridDim.x) * blockDim.x) {
    const edge_t beg = G.get_starting_edge(vertex_t(v)), end = beg + G.get_number_of_neighbors(vertex_t(v));
    weight_t sum = 0;
    for (edge_t e = beg; e < end; ++e) sum += G.get_edge_weight(e);
    iweights[v] = sum != 0 ? alpha / sum : weight_t(0);
  }
}

/// plast = p and dangling += Σ_{iweights==0} alpha·p (double accumulator, one atomic per warp).
template <typename weight_t>
__global__ void __launch_bounds__(256)
    snapshot_kernel(const weight_t* __restrict__ p, weight_t* __restrict__ plast, const weight_t* __restrict__ iweights,
                    std::size_t n, weight_t alpha, double* dangling) {
  double mine = 0;
  for (std::size_t v = std::size_t(blockIdx.x) * blockDim.x + threadIdx.x; v < n;
       v += std::size_t(gridDim.x) * blockDim.x) {
    const weight_t x = p[v];
    plast[v] = x;
    if (iweights[v] == 0) mine += double(alpha) * double(x);
  }
  mine = b200::warp_sum(mine);
  if (b200::lane_id() == 0 && mine != 0) atomicAdd(dangling, mine);
}

template <typename weight_t>
__global__ void __launch_bounds__(256)
    teleport_kernel(weight_t* __restrict__ p, std::size_t n, weight_t alpha, const double* dangling) {
  const weight_t base = weight_t((1.0 - double(alpha) + *dangling) / double(n));
  for (std::size_t v = std::size_t(blockIdx.x) * blockDim.x + threadIdx.x; v < n;
       v += std::size_t(gridDim.x) * blockDim.x)
    p[v] = base;
}

/// Pull iteration: p[v] = base + Σ_{u→v} plast[u]·iweights[u]·w, one warp per vertex, double accumulation.
template <typename vertex_t, typename edge_t, typename weight_t>
__global__ void __launch_bounds__(256)
    gather_kernel(const graph::adjacency_t<vertex_t, edge_t, weight_t> A, const weight_t* __restrict__ plast,
                  const weight_t* __restrict__ iweights, weight_t* __restrict__ p, weight_t alpha,
                  const double* dangling) {
  const double base = (1.0 - double(alpha) + *dangling) / double(A.n);
  const unsigned lane = b200::lane_id();
  const std::size_t warps = (std::size_t(gridDim.x) * blockDim.x) >> 5;
  for (std::size_t v = (std::size_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5; v < std::size_t(A.n); v += warps) {
    const edge_t beg = A.offsets[v], end = A.offsets[v + 1];
    double acc = 0;
    for (edge_t e = beg + lane; e < end; e += 32) {
      const vertex_t u = __ldg(A.indices + e);
      const weight_t w = A.values ? __ldg(A.values + e) : weight_t(1);
      acc += double(plast[u] * iweights[u] * w);
    }
    acc = b200::warp_sum(acc);
    if (lane == 0) p[v] = weight_t(base + acc);
  }
}
}  // namespace kernels

template <typename graph_t, typename param_type, typename result_type>
struct problem_t : gunrock::problem_t<graph_t> {
  param_type param;
  result_type result;
  using vertex_t = typename graph_t::vertex_type;
  using edge_t = typename graph_t::edge_type;
  using weight_t = typename graph_t::weight_type;

  memory::device_array_t<weight_t> plast;
  memory::device_array_t<weight_t> iweights;
  memory::device_array_t<double> scalars;  ///< [0] dangling mass, [1] (as unsigned) max-abs-diff bits

  problem_t(graph_t& G, param_type& _param, result_type& _result, std::shared_ptr<gcuda::multi_context_t> _context)
      : gunrock::problem_t<graph_t>(G, _context), param(_param), result(_result) {}

  void init() override {
    const std::size_t n = std::size_t(this->get_graph().get_number_of_vertices());
    plast.resize(n);
    iweights.resize(n);
    scalars.resize(2);
  }
  void reset() override {
    auto* ctx = this->get_single_context();
    auto g = this->get_graph();
    const std::size_t n = std::size_t(g.get_number_of_vertices());
    b200::fill(*ctx, result.p, n, weight_t(1.0 / double(n)));
    b200::fill(*ctx, plast.data(), n, weight_t(0));
    kernels::inverse_weights_kernel<<<b200::stream_grid(*ctx, n), 256, 0, ctx->stream()>>>(g, param.alpha,
                                                                                          iweights.data());
  }
};

template <typename problem_t, operators::load_balance_t lb, bool pull>
struct enactor_t : gunrock::enactor_t<problem_t> {
  using base_t = gunrock::enactor_t<problem_t>;
  using vertex_t = typename problem_t::vertex_t;
  using edge_t = typename problem_t::edge_t;
  using weight_t = typename problem_t::weight_t;

  enactor_t(problem_t* _problem, std::shared_ptr<gcuda::multi_context_t> _context, enactor_properties_t _properties)
      : base_t(_problem, _context, _properties) {}

  void loop(gcuda::multi_context_t& context) override {
    auto E = this->get_enactor();
    auto P = this->get_problem();
    auto G = P->get_graph();
    auto* ctx = context.get_context(0);
    const std::size_t n = std::size_t(G.get_number_of_vertices());
    weight_t* p = P->result.p;
    weight_t* plast = P->plast.data();
    weight_t* iweights = P->iweights.data();
    const weight_t alpha = P->param.alpha;
    double* dangling = P->scalars.data();

    cudaMemsetAsync(dangling, 0, sizeof(double), ctx->stream());
    kernels::snapshot_kernel<<<b200::stream_grid(*ctx, n), 256, 0, ctx->stream()>>>(p, plast, iweights, n, alpha,
                                                                                    dangling);
    // (extended lambdas may not be defined inside an if-constexpr block, so it is built unconditionally)
    auto spread = [p, plast, iweights] __host__ __device__(vertex_t const& src, vertex_t const& dst,
                                                           edge_t const& edge, weight_t const& weight) -> bool {
      math::atomic::add(p + dst, plast[src] * iweights[src] * weight);
      return false;
    };
    if constexpr (pull) {
      const auto A = graph::adjacency_of<true>(G);
      kernels::gather_kernel<<<gcuda::persistent_grid(*ctx, (n + 7) / 8, 8), 256, 0, ctx->stream()>>>(
          A, plast, iweights, p, alpha, dangling);
      error::check_last("pr gather");
      (void)spread;
      (void)E;
    } else {
      kernels::teleport_kernel<<<b200::stream_grid(*ctx, n), 256, 0, ctx->stream()>>>(p, n, alpha, dangling);
      operators::advance::execute<lb, operators::advance_direction_t::forward, operators::advance_io_type_t::graph,
                                  operators::advance_io_type_t::none>(G, E, spread, context);
    }
  }

  bool is_converged(gcuda::multi_context_t& context) override {
    if (this->iteration == 0) return false;
    auto P = this->get_problem();
    if (this->iteration >= P->param.max_iterations) return true;
    auto* ctx = context.get_context(0);
    const std::size_t n = std::size_t(P->get_graph().get_number_of_vertices());
    unsigned* err_bits = reinterpret_cast<unsigned*>(P->scalars.data() + 1);
    cudaMemsetAsync(err_bits, 0, sizeof(unsigned), ctx->stream());
    b200::kernels::max_abs_diff_kernel<<<b200::stream_grid(*ctx, n), 256, 0, ctx->stream()>>>(
        P->result.p, P->plast.data(), n, err_bits);
    unsigned h_bits = 0;
    cudaMemcpyAsync(&h_bits, err_bits, sizeof(unsigned), cudaMemcpyDeviceToHost, ctx->stream());
    ctx->synchronize();
    float err;
    std::memcpy(&err, &h_bits, sizeof(float));
    return err < P->param.tol;
  }
};

template <operators::load_balance_t lb = operators::load_balance_t::block_mapped, bool pull = false, typename graph_t>
float run(graph_t& G, typename graph_t::weight_type alpha, typename graph_t::weight_type tol,
          typename graph_t::weight_type* p,
          std::shared_ptr<gcuda::multi_context_t> context =
              std::shared_ptr<gcuda::multi_context_t>(new gcuda::multi_context_t(0)),
          int* iterations = nullptr, int max_iterations = 1000) {
  using weight_t = typename graph_t::weight_type;
  using param_type = param_t<weight_t>;
  using result_type = result_t<weight_t>;
  using problem_type = problem_t<graph_t, param_type, result_type>;
  using enactor_type = enactor_t<problem_type, lb, pull>;

  param_type param(alpha, tol, max_iterations);
  result_type result(p);
  problem_type problem(G, param, result, context);
  problem.init();
  problem.reset();
  enactor_properties_t props;
  props.self_manage_frontiers = true;
  enactor_type enactor(&problem, context, props);
  float ms = enactor.enact();
  if (iterations) *iterations = enactor.iteration;
  return ms;
}

}  // namespace pr
}  // namespace gunrock
