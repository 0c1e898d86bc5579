/**
 * @file vector_ops.cuh
 * @brief Streaming n-length vector kernels the algorithm clients run between operators: fill, copy,
 * iota, and fused reductions. They stand in for the thrust::fill / copy_n / transform_reduce calls inside
 * the reference's timed region (pr.hxx:120-133,172-175; ppr.hxx:145; kcore.hxx:188-196 — SURVEY.md K20).
 */
#pragma once

#include <gunrock/b200/warp.cuh>
#include <gunrock/cuda/context.hxx>

namespace gunrock {
namespace b200 {

namespace kernels {
template <typename T>
__global__ void __launch_bounds__(256) fill_kernel(T* out, std::size_t n, T value) {
  for (std::size_t i = std::size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += std::size_t(gridDim.x) * blockDim.x)
    out[i] = value;
}
template <typename T>
__global__ void __launch_bounds__(256) copy_kernel(const T* __restrict__ in, T* __restrict__ out, std::size_t n) {
  for (std::size_t i = std::size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += std::size_t(gridDim.x) * blockDim.x)
    out[i] = in[i];
}
/// *result = max_i |a[i] - b[i]| over non-negative floats (bit pattern orders like unsigned).
static __global__ void __launch_bounds__(256)
    max_abs_diff_kernel(const float* __restrict__ a, const float* __restrict__ b, std::size_t n, unsigned* result) {
  float mine = 0.f;
  for (std::size_t i = std::size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += std::size_t(gridDim.x) * blockDim.x)
    mine = fmaxf(mine, fabsf(a[i] - b[i]));
  mine = warp_max(mine);
  if (lane_id() == 0) atomicMax(result, __float_as_uint(mine));
}
/// *result += Σ x[i] in double (one atomic per warp).
template <typename T>
__global__ void __launch_bounds__(256) sum_kernel(const T* __restrict__ x, std::size_t n, double* result) {
  double mine = 0;
  for (std::size_t i = std::size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += std::size_t(gridDim.x) * blockDim.x)
    mine += double(x[i]);
  mine = warp_sum(mine);
  if (lane_id() == 0) atomicAdd(result, mine);
}
/// *result &= all(flags) — counts zeros; result is the number of false entries.
static __global__ void __launch_bounds__(256) count_false_kernel(const bool* __restrict__ flags, std::size_t n, counter_t* result) {
  counter_t mine = 0;
  for (std::size_t i = std::size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += std::size_t(gridDim.x) * blockDim.x)
    mine += flags[i] ? 0 : 1;
  mine = warp_sum(mine);
  if (lane_id() == 0 && mine) atomicAdd(result, mine);
}
}  // namespace kernels

inline unsigned stream_grid(gcuda::standard_context_t& ctx, std::size_t n) {
  return gcuda::persistent_grid(ctx, (n + 255) / 256, 8);
}

template <typename T>
void fill(gcuda::standard_context_t& ctx, T* out, std::size_t n, T value) {
  if (n) kernels::fill_kernel<<<stream_grid(ctx, n), 256, 0, ctx.stream()>>>(out, n, value);
  ctx.profiler().launches_total += n ? 1 : 0;
}
template <typename T>
void copy(gcuda::standard_context_t& ctx, const T* in, T* out, std::size_t n) {
  if (n) kernels::copy_kernel<<<stream_grid(ctx, n), 256, 0, ctx.stream()>>>(in, out, n);
  ctx.profiler().launches_total += n ? 1 : 0;
}
template <typename T>
void set_one(gcuda::standard_context_t& ctx, T* at, T value) {
  kernels::fill_kernel<<<1, 32, 0, ctx.stream()>>>(at, std::size_t(1), value);
  ctx.profiler().launches_total += 1;
}

}  // namespace b200
}  // namespace gunrock
