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

template <typename problem_t, operators::load_balance_t lb>
struct enactor_t : gunrock::enactor_t<problem_t> {
  using base_t = gunrock::enactor_t<problem_t>;
  using vertex_t = typename problem_t::vertex_t;
  using edge_t = typename problem_t::edge_t;
  using weight_t = typename problem_t::weight_t;
  using frontier_t = typename base_t::frontier_t;

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

template <operators::load_balance_t lb = operators::load_balance_t::block_mapped, typename graph_t>
float run(graph_t& G, typename graph_t::vertex_type& single_source, typename graph_t::weight_type* distances,
          typename graph_t::vertex_type* predecessors,
          std::shared_ptr<gcuda::multi_context_t> context =
              std::shared_ptr<gcuda::multi_context_t>(new gcuda::multi_context_t(0)),
          int* iterations = nullptr) {
  using vertex_t = typename graph_t::vertex_type;
  using weight_t = typename graph_t::weight_type;
  using param_type = param_t<vertex_t>;
  using result_type = result_t<vertex_t, weight_t>;
  using problem_type = problem_t<graph_t, param_type, result_type>;
  using enactor_type = enactor_t<problem_type, lb>;

  param_type param(single_source);
  result_type result(distances, predecessors, G.get_number_of_vertices());
  problem_type problem(G, param, result, context);
  problem.init();
  problem.reset();
  enactor_type enactor(&problem, context);
  float ms = enactor.enact();
  if (iterations) *iterations = enactor.iteration;
  return ms;
}

}  // namespace sssp
}  // namespace gunrock
