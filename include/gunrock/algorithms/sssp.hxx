/**
 * @file sssp.hxx
 * @brief Single-source shortest paths (label-correcting) client of the frontier operators.
 *
 * Same operators-level recipe as the reference (include/gunrock/algorithms/sssp.hxx:98-151): advance with
 * `d = dist[src] + w; old = atomic::min(&dist[nbr], d); keep iff d < old`, then a bypass filter that drops
 * a vertex already seen in this iteration (`visited[v] == iteration`, benign race). reset (:61-78): FLT_MAX
 * everywhere, 0 at the source, visited = -1. The single float add per relaxation is rounded exactly like
 * the CPU reference's, and min over monotone relaxations has a unique fixed point, so distances are
 * bit-identical to sssp_cpu for non-negative weights. run() keeps the reference signature (:155-163); the
 * leading template parameter picks the balancer (default block_mapped as in the reference).
 *
 * `near_far = true` (additive): the same relaxation operator driven by operators::advance::execute_near_far —
 * Davidson et al.'s near/far ordering inside one persistent kernel (advance/near_far.cuh) — for high-diameter
 * graphs where the level-per-launch loop above is latency- and rework-bound. `delta` is the bucket width
 * (<= 0: the mean edge weight). Same distances, bit for bit.
 */
#pragma once

#include <gunrock/algorithms/algorithms.hxx>

namespace gunrock {
namespace sssp {

template <typename vertex_t>
struct param_t {
  vertex_t single_source;
  param_t(vertex_t _single_source) : single_source(_single_source) {}
};

template <typename vertex_t, typename weight_t>
struct result_t {
  weight_t* distances;
  vertex_t* predecessors;
  result_t(weight_t* _distances, vertex_t* _predecessors, vertex_t) : distances(_distances), predecessors(_predecessors) {}
};

template <typename graph_t, typename param_type, typename result_type>
struct problem_t : gunrock::problem_t<graph_t> {
  param_type param;
  result_type result;
  using vertex_t = typename graph_t::vertex_type;
  using edge_t = typename graph_t::edge_type;
  using weight_t = typename graph_t::weight_type;

  memory::device_array_t<vertex_t> visited;

  problem_t(graph_t& G, param_type& _param, result_type& _result, std::shared_ptr<gcuda::multi_context_t> _context)
      : gunrock::problem_t<graph_t>(G, _context), param(_param), result(_result) {}

  void init() override { visited.resize(std::size_t(this->get_graph().get_number_of_vertices())); }
  void reset() override {
    auto* ctx = this->get_single_context();
    const std::size_t n = std::size_t(this->get_graph().get_number_of_vertices());
    b200::fill(*ctx, result.distances, n, std::numeric_limits<weight_t>::max());
    b200::set_one(*ctx, result.distances + param.single_source, weight_t(0));
    b200::fill(*ctx, visited.data(), n, vertex_t(-1));
  }
};

template <typename problem_t, operators::load_balance_t lb, bool near_far = false>
struct enactor_t : gunrock::enactor_t<problem_t> {
  using base_t = gunrock::enactor_t<problem_t>;
  using vertex_t = typename problem_t::vertex_t;
  using edge_t = typename problem_t::edge_t;
  using weight_t = typename problem_t::weight_t;
  using frontier_t = typename base_t::frontier_t;

  float delta = 0.f;                               ///< near-far bucket width
  operators::advance::near_far_result_t near_far_stats;  ///< filled by the near-far path

  enactor_t(problem_t* _problem, std::shared_ptr<gcuda::multi_context_t> _context) : base_t(_problem, _context) {}

  void prepare_frontier(frontier_t* f, gcuda::multi_context_t& context) override {
    f->push_back(this->get_problem()->param.single_source);
  }

  void loop(gcuda::multi_context_t& context) override {
    auto E = this->get_enactor();
    auto P = this->get_problem();
    auto G = P->get_graph();
    auto distances = P->result.distances;
    auto visited = P->visited.data();
    const vertex_t iteration = vertex_t(this->iteration);

    // near-far form of the same relaxation: inside the persistent kernel the labels are read through L2
    auto relax = [distances] __host__ __device__(vertex_t const& source, vertex_t const& neighbor,
                                                 edge_t const& edge, weight_t const& weight) -> bool {
      const weight_t candidate = thread::load<thread::cache_t::global>(&distances[source]) + weight;
      return candidate < math::atomic::min(&distances[neighbor], candidate);
    };
    auto tentative = [distances] __host__ __device__(vertex_t const& v) -> float {
      return float(thread::load<thread::cache_t::global>(&distances[v]));
    };
    if constexpr (near_far) {
      near_far_stats = operators::advance::execute_near_far(G, E, relax, tentative, delta, context);
      this->iteration += near_far_stats.levels > 0 ? near_far_stats.levels - 1 : 0;  // enact() adds the last one
      (void)visited;
      (void)iteration;
      return;
    }

    auto shortest_path = [distances] __host__ __device__(vertex_t const& source, vertex_t const& neighbor,
                                                         edge_t const& edge, weight_t const& weight) -> bool {
      const weight_t candidate = thread::load(&distances[source]) + weight;
      return candidate < math::atomic::min(&distances[neighbor], candidate);
    };
    auto drop_repeats = [visited, iteration] __host__ __device__(vertex_t const& vertex) -> bool {
      if (visited[vertex] == iteration) return false;
      visited[vertex] = iteration;
      return true;
    };
    operators::advance::execute<lb>(G, E, shortest_path, context);
    operators::filter::execute<operators::filter_algorithm_t::bypass>(G, E, drop_repeats, context);
  }
};

template <operators::load_balance_t lb = operators::load_balance_t::block_mapped, bool near_far = false,
          typename graph_t>
float run(graph_t& G, typename graph_t::vertex_type& single_source, typename graph_t::weight_type* distances,
          typename graph_t::vertex_type* predecessors,
          std::shared_ptr<gcuda::multi_context_t> context =
              std::shared_ptr<gcuda::multi_context_t>(new gcuda::multi_context_t(0)),
          int* iterations = nullptr, float delta = 0.f, long long* work_stats = nullptr) {
  using vertex_t = typename graph_t::vertex_type;
  using weight_t = typename graph_t::weight_type;
  using param_type = param_t<vertex_t>;
  using result_type = result_t<vertex_t, weight_t>;
  using problem_type = problem_t<graph_t, param_type, result_type>;
  using enactor_type = enactor_t<problem_type, lb, near_far>;

  param_type param(single_source);
  result_type result(distances, predecessors, G.get_number_of_vertices());
  problem_type problem(G, param, result, context);
  problem.init();
  problem.reset();
  enactor_type enactor(&problem, context);
  if constexpr (near_far) {
    if (!(delta > 0.f)) {  // default bucket width: 64 x mean edge weight / mean degree (Davidson et al.'s rule of thumb)
      auto* ctx = context->get_context(0);
      const auto weights = graph::adjacency_of<false>(G).values;
      const auto m = G.get_number_of_edges();
      delta = 1.f;
      if (weights && m > 0) {
        memory::device_array_t<double> sum(1);
        cudaMemsetAsync(sum.data(), 0, sizeof(double), ctx->stream());
        b200::kernels::sum_kernel<<<b200::stream_grid(*ctx, std::size_t(m)), 256, 0, ctx->stream()>>>(
            weights, std::size_t(m), sum.data());
        double h = 0;
        cudaMemcpyAsync(&h, sum.data(), sizeof(double), cudaMemcpyDeviceToHost, ctx->stream());
        ctx->synchronize();
        if (h > 0) delta = float(h / double(m));
      }
      const double mean_degree = double(m) / double(G.get_number_of_vertices() > 0 ? G.get_number_of_vertices() : 1);
      delta = float(double(delta) * 64.0 / (mean_degree > 1.0 ? mean_degree : 1.0));
    }
    enactor.delta = delta;
  }
  float ms = enactor.enact();
  if (iterations) *iterations = enactor.iteration;
  if (work_stats) {  // [0] near levels, [1] far splits, [2] relaxations (near-far path only)
    work_stats[0] = enactor.near_far_stats.levels;
    work_stats[1] = enactor.near_far_stats.splits;
    work_stats[2] = (long long)enactor.near_far_stats.relaxations;
  }
  return ms;
}

}  // namespace sssp
}  // namespace gunrock
