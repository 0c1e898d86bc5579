/**
 * @file kcore.hxx
 * @brief k-core decomposition client of the frontier operators.
 *
 * Same peel as the reference (include/gunrock/algorithms/kcore.hxx:112-199): for k = iteration+1, repeat
 * { advance: a live source with degree <= k gets k_core = k, is marked to-be-deleted, and emits its live
 * neighbours; parallel_for<vertex>: deleted |= to_be_deleted; filter<predicated>: every emitted neighbour
 * decrements its degree atomically and survives iff the OLD degree was k+1 } until the frontier empties;
 * is_converged refills the frontier with all vertices and stops when every vertex is deleted.
 * The inner while watches the enactor's CURRENT input frontier (the reference captures the buffer pointer
 * once, :123, and relies on advance+filter swapping twice; here the pointer is re-read every trip, which
 * is the same buffer after an even number of swaps). Flag arrays are bytes, not thrust::device_vector<bool>.
 */
#pragma once

#include <gunrock/algorithms/algorithms.hxx>

namespace gunrock {
namespace kcore {

template <typename vertex_t>
struct result_t {
  int* k_cores;
  result_t(int* _k_cores) : k_cores(_k_cores) {}
};

namespace kernels {
template <typename graph_t>
__global__ void __launch_bounds__(256) reset_kernel(graph_t G, int* degrees, bool* deleted, bool* to_be_deleted,
                                                    int* k_cores) {
  using vertex_t = typename graph_t::vertex_type;
  const std::size_t n = std::size_t(G.get_number_of_vertices());
  for (std::size_t v = std::size_t(blockIdx.x) * blockDim.x + threadIdx.x; v < n;
       v += std::size_t(gridDim.x) * blockDim.x) {
    const int d = int(G.get_number_of_neighbors(vertex_t(v)));
    degrees[v] = d;
    deleted[v] = d == 0;
    to_be_deleted[v] = false;
    k_cores[v] = 0;
  }
}
}  // namespace kernels

template <typename graph_t, typename result_type>
struct problem_t : gunrock::problem_t<graph_t> {
  result_type result;
  using vertex_t = typename graph_t::vertex_type;
  using edge_t = typename graph_t::edge_type;
  using weight_t = typename graph_t::weight_type;

  memory::device_array_t<int> degrees;
  memory::device_array_t<bool> deleted, to_be_deleted;
  memory::device_array_t<b200::counter_t> alive_counter;

  problem_t(graph_t& G, result_type& _result, std::shared_ptr<gcuda::multi_context_t> _context)
      : gunrock::problem_t<graph_t>(G, _context), result(_result) {}

  void init() override {
    const std::size_t n = std::size_t(this->get_graph().get_number_of_vertices());
    degrees.resize(n);
    deleted.resize(n);
    to_be_deleted.resize(n);
    alive_counter.resize(1);
  }
  void reset() override {
    auto* ctx = this->get_single_context();
    auto g = this->get_graph();
    const std::size_t n = std::size_t(g.get_number_of_vertices());
    kernels::reset_kernel<<<b200::stream_grid(*ctx, n), 256, 0, ctx->stream()>>>(
        g, degrees.data(), deleted.data(), to_be_deleted.data(), result.k_cores);
  }
};

template <typename problem_t, operators::load_balance_t lb>
struct enactor_t : gunrock::enactor_t<problem_t> {
  using base_t = gunrock::enactor_t<problem_t>;
  using vertex_t = typename problem_t::vertex_t;
  using edge_t = typename problem_t::edge_t;
  using weight_t = typename problem_t::weight_t;
  using frontier_t = typename base_t::frontier_t;
  bool verbose = false;

  enactor_t(problem_t* _problem, std::shared_ptr<gcuda::multi_context_t> _context) : base_t(_problem, _context) {}

  void prepare_frontier(frontier_t* f, gcuda::multi_context_t& context) override {
    const std::size_t n = std::size_t(this->get_problem()->get_graph().get_number_of_vertices());
    f->sequence(vertex_t(0), n, context.get_context(0)->stream());
  }

  void loop(gcuda::multi_context_t& context) override {
    auto E = this->get_enactor();
    auto P = this->get_problem();
    auto G = P->get_graph();
    int* k_cores = P->result.k_cores;
    int* degrees = P->degrees.data();
    bool* deleted = P->deleted.data();
    bool* to_be_deleted = P->to_be_deleted.data();
    const int k = this->iteration + 1;

    auto peel = [=] __host__ __device__(vertex_t const& source, vertex_t const& neighbor, edge_t const& edge,
                                        weight_t const& weight) -> bool {
      if (deleted[source] || degrees[source] > k) return false;
      k_cores[source] = k;
      to_be_deleted[source] = true;
      return !deleted[neighbor];
    };
    auto commit = [=] __device__(vertex_t const& v) { deleted[v] = deleted[v] | to_be_deleted[v]; };
    auto decrement = [=] __host__ __device__(vertex_t const& vertex) -> bool {
      if (deleted[vertex]) return false;
      return math::atomic::add(&degrees[vertex], -1) == k + 1;
    };

    while (!this->get_input_frontier()->is_empty()) {
      operators::advance::execute<lb>(G, E, peel, context);
      operators::parallel_for::execute<operators::parallel_for_each_t::vertex>(G, commit, context);
      operators::filter::execute<operators::filter_algorithm_t::predicated>(G, E, decrement, context);
    }
  }

  bool is_converged(gcuda::multi_context_t& context) override {
    auto P = this->get_problem();
    auto* ctx = context.get_context(0);
    const std::size_t n = std::size_t(P->get_graph().get_number_of_vertices());
    b200::counter_t* alive = P->alive_counter.data();
    cudaMemsetAsync(alive, 0, sizeof(b200::counter_t), ctx->stream());
    b200::kernels::count_false_kernel<<<b200::stream_grid(*ctx, n), 256, 0, ctx->stream()>>>(P->deleted.data(), n,
                                                                                             alive);
    b200::counter_t h_alive = 0;
    cudaMemcpyAsync(&h_alive, alive, sizeof(h_alive), cudaMemcpyDeviceToHost, ctx->stream());
    ctx->synchronize();
    const bool graph_empty = h_alive == 0;
    if (graph_empty && verbose) std::printf("degeneracy = %d\n", this->iteration);
    this->get_input_frontier()->sequence(vertex_t(0), n, ctx->stream());
    return graph_empty;
  }
};

template <operators::load_balance_t lb = operators::load_balance_t::block_mapped, typename graph_t>
float run(graph_t& G, int* k_cores,
          std::shared_ptr<gcuda::multi_context_t> context =
              std::shared_ptr<gcuda::multi_context_t>(new gcuda::multi_context_t(0))) {
  using result_type = result_t<int>;
  using problem_type = problem_t<graph_t, result_type>;
  using enactor_type = enactor_t<problem_type, lb>;
  result_type result(k_cores);
  problem_type problem(G, result, context);
  problem.init();
  problem.reset();
  enactor_type enactor(&problem, context);
  return enactor.enact();
}

}  // namespace kcore
}  // namespace gunrock
