/**
 * @file neighborreduce.hxx
 * @brief operators::neighborreduce::execute — for every vertex v, output[v] = reduce over v's edges e of op(e)
 * with a user binary operator and initial value (a segmented reduction over the CSR rows).
 * Signature as the reference (framework/operators/neighborreduce/neighborreduce.hxx:55-67, which forwards to
 * mgpu::transform_segreduce); used by SpMV's pull form (algorithms/spmv.hxx:107-127). Here: one warp per row,
 * lanes stride the row (coalesced edge ids), a shuffle tree applies the user's operator — no atomics, and the
 * result does not depend on scheduling.
 */
#pragma once

#include <gunrock/cuda/cuda.hxx>
#include <gunrock/error.hxx>
#include <gunrock/b200/warp.cuh>
#include <gunrock/framework/operators/configs.hxx>

namespace gunrock {
namespace operators {
namespace neighborreduce {

namespace kernels {
template <typename vertex_t, typename edge_t, typename output_t, typename operator_t, typename arithmetic_t>
__global__ void __launch_bounds__(256)
    row_reduce_kernel(const edge_t* __restrict__ offsets, vertex_t n, output_t* __restrict__ output, operator_t op,
                      arithmetic_t arithmetic_op, output_t init_value) {
  const unsigned lane = b200::lane_id();
  const std::size_t warps = (std::size_t(gridDim.x) * blockDim.x) >> 5;
  for (std::size_t v = (std::size_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5; v < std::size_t(n); v += warps) {
    const edge_t beg = offsets[v], end = offsets[v + 1];
    output_t acc = init_value;
    bool have = false;  // lanes without an edge must not inject init_value into a non-idempotent operator
    for (edge_t e = beg + edge_t(lane); e < end; e += 32) {
      const output_t x = op(e);
      acc = have ? arithmetic_op(acc, x) : x;
      have = true;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      const output_t other = __shfl_down_sync(b200::full_mask, acc, d);
      const bool other_has = __shfl_down_sync(b200::full_mask, have, d);
      if (lane + d < 32 && other_has) {
        acc = have ? arithmetic_op(acc, other) : other;
        have = true;
      }
    }
    if (lane == 0) output[v] = have ? arithmetic_op(init_value, acc) : init_value;
  }
}
}  // namespace kernels

template <advance_io_type_t input_t = advance_io_type_t::graph, typename graph_t, typename enactor_t,
          typename output_t, typename operator_t, typename arithmetic_t>
void execute(graph_t& G, enactor_t* E, output_t* output, operator_t op, arithmetic_t arithmetic_op,
             output_t init_value, gcuda::multi_context_t& context) {
  error::throw_if_exception(context.size() != 1, "`context.size() != 1` not supported");
  static_assert(input_t == advance_io_type_t::graph, "neighborreduce takes the whole graph as input");
  using csr_v = typename graph_t::graph_csr_view_t;
  static_assert(graph_t::template contains_representation<csr_v>(),
                "CSR sparse-matrix representation required for neighborreduce operator.");
  auto* ctx = context.get_context(0);
  const auto n = G.get_number_of_vertices();
  if (n <= 0) return;
  kernels::row_reduce_kernel<<<gcuda::persistent_grid(*ctx, (std::size_t(n) + 7) / 8, 8), 256, 0, ctx->stream()>>>(
      G.get_row_offsets(), n, output, op, arithmetic_op, init_value);
  error::check_last("neighborreduce");
  ctx->synchronize();
  (void)E;
}

}  // namespace neighborreduce
}  // namespace operators
}  // namespace gunrock
