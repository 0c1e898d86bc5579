/**
 * @file ppr.hxx
 * @brief Personalised PageRank (push / residual propagation) client of the frontier operators.
 *
 * Same two-operator iteration as the reference (include/gunrock/algorithms/ppr.hxx:105-146): a predicated
 * filter whose operator has side effects (p[v] += 2a/(1+a)·r[v]; r'[v] = 0; keep), then an advance whose
 * operator USES THE OLD VALUE returned by atomic::add to detect the residual crossing deg(dst)·epsilon, then
 * r = r'. reset :71-88. run / run_batch signatures as :150-158, :170-203 (batch = one host thread per seed,
 * each with its own context, reference batch.hxx:61-79).
 */
#pragma once

#include <thread>
#include <vector>
#include <gunrock/algorithms/algorithms.hxx>

namespace gunrock {
namespace ppr {

template <typename vertex_t, typename weight_t>
struct param_t {
  vertex_t seed;
  weight_t alpha;
  weight_t epsilon;
  param_t(vertex_t _seed, weight_t _alpha, weight_t _epsilon) : seed(_seed), alpha(_alpha), epsilon(_epsilon) {}
};

template <typename weight_t>
struct result_t {
  weight_t* p;
  result_t(weight_t* _p) : p(_p) {}
};

template <typename graph_t, typename param_type, typename result_type>
struct problem_t : gunrock::problem_t<graph_t> {
  param_type param;
  result_type result;
  using vertex_t = typename graph_t::vertex_type;
  using edge_t = typename graph_t::edge_type;
  using weight_t = typename graph_t::weight_type;

  memory::device_array_t<weight_t> r, r_prime;
  weight_t _2a1a, _1a1a;

  problem_t(graph_t& G, param_type& _param, result_type& _result, std::shared_ptr<gcuda::multi_context_t> _context)
      : gunrock::problem_t<graph_t>(G, _context), param(_param), result(_result) {}

  void init() override {
    const std::size_t n = std::size_t(this->get_graph().get_number_of_vertices());
    r.resize(n);
    r_prime.resize(n);
    const weight_t alpha = param.alpha;
    _2a1a = (2 * alpha) / (1 + alpha);
    _1a1a = (1 - alpha) / (1 + alpha);
  }
  void reset() override {
    auto* ctx = this->get_single_context();
    const std::size_t n = std::size_t(this->get_graph().get_number_of_vertices());
    b200::fill(*ctx, result.p, n, weight_t(0));
    b200::fill(*ctx, r.data(), n, weight_t(0));
    b200::fill(*ctx, r_prime.data(), n, weight_t(0));
    b200::set_one(*ctx, r.data() + param.seed, weight_t(1));
    b200::set_one(*ctx, r_prime.data() + param.seed, weight_t(1));
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
    f->push_back(this->get_problem()->param.seed);
  }

  void loop(gcuda::multi_context_t& context) override {
    auto E = this->get_enactor();
    auto P = this->get_problem();
    auto G = P->get_graph();
    weight_t* p = P->result.p;
    weight_t* r = P->r.data();
    weight_t* r_prime = P->r_prime.data();
    const weight_t epsilon = P->param.epsilon, _2a1a = P->_2a1a, _1a1a = P->_1a1a;

    auto settle = [p, r, r_prime, _2a1a] __host__ __device__(vertex_t const& vertex) -> bool {
      p[vertex] += _2a1a * r[vertex];
      r_prime[vertex] = 0;
      return true;
    };
    operators::filter::execute<operators::filter_algorithm_t::predicated>(G, E, settle, context);

    auto push_residual = [G, r, r_prime, _1a1a, epsilon] __host__ __device__(
                             vertex_t const& src, vertex_t const& dst, edge_t const& edge,
                             weight_t const& weight) -> bool {
      const weight_t update = _1a1a * r[src] / weight_t(G.get_number_of_neighbors(src));
      const weight_t before = math::atomic::add(r_prime + dst, update);
      const weight_t after = before + update;
      const weight_t threshold = weight_t(G.get_number_of_neighbors(dst)) * epsilon;
      return before < threshold && after >= threshold;
    };
    operators::advance::execute<lb>(G, E, push_residual, context);

    auto* ctx = context.get_context(0);
    b200::copy(*ctx, r_prime, r, std::size_t(G.get_number_of_vertices()));
  }
};

template <operators::load_balance_t lb = operators::load_balance_t::block_mapped, typename graph_t>
float run(graph_t& G, typename graph_t::vertex_type& seed, typename graph_t::weight_type* p,
          typename graph_t::weight_type& alpha, typename graph_t::weight_type& epsilon,
          std::shared_ptr<gcuda::multi_context_t> context =
              std::shared_ptr<gcuda::multi_context_t>(new gcuda::multi_context_t(0))) {
  using vertex_t = typename graph_t::vertex_type;
  using weight_t = typename graph_t::weight_type;
  using param_type = param_t<vertex_t, weight_t>;
  using result_type = result_t<weight_t>;
  using problem_type = problem_t<graph_t, param_type, result_type>;
  using enactor_type = enactor_t<problem_type, lb>;

  param_type param(seed, alpha, epsilon);
  result_type result(p);
  problem_type problem(G, param, result, context);
  problem.init();
  problem.reset();
  enactor_type enactor(&problem, context);
  return enactor.enact();
}

/// Seeds 0..n_seeds-1, one host thread and one fresh context each; p is n_seeds x n. Returns summed ms.
template <operators::load_balance_t lb = operators::load_balance_t::block_mapped, typename graph_t>
float run_batch(graph_t& G, typename graph_t::vertex_type& n_seeds, typename graph_t::weight_type* p,
                typename graph_t::weight_type& alpha, typename graph_t::weight_type& epsilon) {
  using vertex_t = typename graph_t::vertex_type;
  const std::size_t n = std::size_t(G.get_number_of_vertices());
  std::vector<float> elapsed(std::size_t(n_seeds), 0.f);
  std::vector<std::thread> workers;
  int device = 0;
  cudaGetDevice(&device);
  for (vertex_t job = 0; job < n_seeds; ++job)
    workers.emplace_back([&, job]() {
      cudaSetDevice(device);
      vertex_t seed = job;
      elapsed[std::size_t(job)] = ppr::run<lb>(G, seed, p + n * std::size_t(job), alpha, epsilon);
    });
  float total = 0.f;
  for (std::size_t j = 0; j < workers.size(); ++j) {
    workers[j].join();
    total += elapsed[j];
  }
  return total;
}

}  // namespace ppr
}  // namespace gunrock
