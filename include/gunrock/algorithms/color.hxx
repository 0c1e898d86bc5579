/**
 * @file color.hxx
 * @brief Graph colouring (Jones-Plassmann style, two colours per round) client of the filter operator.
 *
 * Same predicate as the reference (include/gunrock/algorithms/color.hxx:99-146): in round `it` an uncoloured
 * vertex compares its random number with every neighbour that is not yet finally coloured (a neighbour is
 * skipped only when it holds a colour other than 2·it and 2·it+1, i.e. from an EARLIER round); a local
 * maximum takes colour 2·it, a local minimum 2·it+1, ties break on vertex id; coloured vertices leave the
 * frontier (filter<predicated>). Note that a neighbour coloured in the CURRENT round still takes part in
 * the comparison exactly as if it were uncoloured, so reading colours while other threads write them gives
 * the same decisions as a per-round snapshot: the result is deterministic and equals the Jacobi oracle.
 * Randoms: generate::random::uniform_distribution over (0, n) (color.hxx:65).
 */
#pragma once

#include <gunrock/algorithms/algorithms.hxx>
#include <gunrock/algorithms/generate/random.hxx>

namespace gunrock {
namespace color {

struct param_t {};

template <typename vertex_t>
struct result_t {
  vertex_t* colors;
  result_t(vertex_t* colors_) : colors(colors_) {}
};

template <typename graph_t, typename param_type, typename result_type>
struct problem_t : gunrock::problem_t<graph_t> {
  param_type param;
  result_type result;
  using vertex_t = typename graph_t::vertex_type;
  using edge_t = typename graph_t::edge_type;
  using weight_t = typename graph_t::weight_type;

  memory::device_array_t<float> randoms;

  problem_t(graph_t& G, param_type& _param, result_type& _result, std::shared_ptr<gcuda::multi_context_t> _context)
      : gunrock::problem_t<graph_t>(G, _context), param(_param), result(_result) {}

  void init() override { randoms.resize(std::size_t(this->get_graph().get_number_of_vertices())); }
  void reset() override {
    auto* ctx = this->get_single_context();
    const std::size_t n = std::size_t(this->get_graph().get_number_of_vertices());
    b200::fill(*ctx, result.colors, n, gunrock::numeric_limits<vertex_t>::invalid());
    generate::random::uniform_distribution(randoms.data(), n, 0.0f, float(n), ctx->stream());
  }
};

template <typename problem_t>
struct enactor_t : gunrock::enactor_t<problem_t> {
  using base_t = gunrock::enactor_t<problem_t>;
  using vertex_t = typename problem_t::vertex_t;
  using edge_t = typename problem_t::edge_t;
  using weight_t = typename problem_t::weight_t;
  using frontier_t = typename base_t::frontier_t;

  enactor_t(problem_t* _problem, std::shared_ptr<gcuda::multi_context_t> _context) : base_t(_problem, _context) {}

  void prepare_frontier(frontier_t* f, gcuda::multi_context_t& context) override {
    const std::size_t n = std::size_t(this->get_problem()->get_graph().get_number_of_vertices());
    f->sequence(vertex_t(0), n, context.get_context(0)->stream());
  }

  void loop(gcuda::multi_context_t& context) override {
    auto E = this->get_enactor();
    auto P = this->get_problem();
    auto G = P->get_graph();
    vertex_t* colors = P->result.colors;
    const float* randoms = P->randoms.data();
    const vertex_t first_color = vertex_t(2 * this->iteration);

    auto try_color = [G, colors, randoms, first_color] __host__ __device__(vertex_t const& vertex) -> bool {
      const edge_t degree = G.get_number_of_neighbors(vertex);
      if (degree == 0) {
        colors[vertex] = first_color;
        return false;
      }
      bool is_max = true, is_min = true;
      const edge_t begin = G.get_starting_edge(vertex);
      const float mine = randoms[vertex];
      for (edge_t e = begin; e < begin + degree; ++e) {
        const vertex_t u = G.get_destination_vertex(e);
        const vertex_t cu = colors[u];
        if ((util::limits::is_valid(cu) && cu != first_color && cu != first_color + 1) || u == vertex) continue;
        const float theirs = randoms[u];
        if (mine < theirs || (mine == theirs && vertex < u)) is_max = false;
        if (mine > theirs || (mine == theirs && vertex > u)) is_min = false;
      }
      if (is_max) {
        colors[vertex] = first_color;
        return false;
      }
      if (is_min) {
        colors[vertex] = first_color + 1;
        return false;
      }
      return true;
    };
    operators::filter::execute<operators::filter_algorithm_t::predicated>(G, E, try_color, context);
  }
};

template <typename graph_t>
float run(graph_t& G, typename graph_t::vertex_type* colors,
          std::shared_ptr<gcuda::multi_context_t> context =
              std::shared_ptr<gcuda::multi_context_t>(new gcuda::multi_context_t(0)),
          int* iterations = nullptr) {
  using vertex_t = typename graph_t::vertex_type;
  using param_type = param_t;
  using result_type = result_t<vertex_t>;
  using problem_type = problem_t<graph_t, param_type, result_type>;
  using enactor_type = enactor_t<problem_type>;
  param_type param;
  result_type result(colors);
  problem_type problem(G, param, result, context);
  problem.init();
  problem.reset();
  enactor_type enactor(&problem, context);
  float ms = enactor.enact();
  if (iterations) *iterations = enactor.iteration;
  return ms;
}

}  // namespace color
}  // namespace gunrock
