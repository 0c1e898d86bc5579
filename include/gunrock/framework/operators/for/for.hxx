/**
 * @file for.hxx
 * @brief operators::parallel_for::execute — apply a functor to every vertex / edge / weight of a graph or
 * every valid element of a frontier. Signatures as the reference (framework/operators/for/for.hxx:28-30,
 * 60-62, which call thrust::for_each); here a grid-stride kernel on the context's stream.
 */
#pragma once

#include <type_traits>
#include <gunrock/cuda/cuda.hxx>
#include <gunrock/framework/operators/configs.hxx>
#include <gunrock/util/type_limits.hxx>

namespace gunrock {
namespace operators {
namespace parallel_for {

namespace kernels {
template <typename index_t, typename func_t>
__global__ void __launch_bounds__(256) for_each_index_kernel(std::size_t count, func_t op) {
  for (std::size_t i = std::size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < count;
       i += std::size_t(gridDim.x) * blockDim.x) {
    index_t x = index_t(i);
    op(x);
  }
}
template <typename graph_t, typename func_t>
__global__ void __launch_bounds__(256) for_each_weight_kernel(graph_t G, std::size_t count, func_t op) {
  using edge_t = typename graph_t::edge_type;
  for (std::size_t i = std::size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < count;
       i += std::size_t(gridDim.x) * blockDim.x)
    op(G.get_edge_weight(edge_t(i)));
}
template <typename type_t, typename func_t>
__global__ void __launch_bounds__(256) for_each_element_kernel(const type_t* items, std::size_t count, func_t op) {
  for (std::size_t i = std::size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < count;
       i += std::size_t(gridDim.x) * blockDim.x) {
    type_t x = items[i];
    if (gunrock::util::limits::is_valid(x)) op(x);
  }
}
}  // namespace kernels

/// element: every valid element of a frontier.
template <parallel_for_each_t type, typename func_t, typename frontier_t>
std::enable_if_t<type == parallel_for_each_t::element> execute(frontier_t& f, func_t op,
                                                               gcuda::multi_context_t& context) {
  using type_t = typename frontier_t::type_t;
  auto* ctx = context.get_context(0);
  const std::size_t count = f.get_number_of_elements();
  if (!count) return;
  kernels::for_each_element_kernel<type_t>
      <<<gcuda::persistent_grid(*ctx, (count + 255) / 256, 8), 256, 0, ctx->stream()>>>(f.data(), count, op);
  error::check_last("parallel_for element");
  ctx->synchronize();  // as the reference (blocking thrust::cuda::par, for.hxx:33-41): callers read results on return
}

/// vertex / edge / weight: every vertex id, edge id or edge weight of the graph.
template <parallel_for_each_t type, typename func_t, typename graph_t>
std::enable_if_t<type != parallel_for_each_t::element> execute(graph_t& G, func_t op,
                                                               gcuda::multi_context_t& context) {
  using index_t = std::conditional_t<type == parallel_for_each_t::vertex, typename graph_t::vertex_type,
                                     typename graph_t::edge_type>;
  auto* ctx = context.get_context(0);
  const std::size_t count = type == parallel_for_each_t::vertex ? std::size_t(G.get_number_of_vertices())
                                                                : std::size_t(G.get_number_of_edges());
  if (!count) return;
  const unsigned grid = gcuda::persistent_grid(*ctx, (count + 255) / 256, 8);
  if constexpr (type == parallel_for_each_t::weight)
    kernels::for_each_weight_kernel<<<grid, 256, 0, ctx->stream()>>>(G, count, op);
  else
    kernels::for_each_index_kernel<index_t><<<grid, 256, 0, ctx->stream()>>>(count, op);
  error::check_last("parallel_for");
  ctx->synchronize();  // reference for.hxx:70-100 blocks too; mst.hxx:236-244 copies a flag to the host right after
}

}  // namespace parallel_for
}  // namespace operators
}  // namespace gunrock
