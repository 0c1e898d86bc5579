/**
 * @file launch.hxx
 * @brief Launch geometry for B200 and the thread-index helpers of the reference's gcuda namespace.
 *
 * The reference picks launch parameters at compile time from an SM_TARGET macro through launch_box_t
 * (include/gunrock/cuda/launch_box.hxx:194-335, sm flags end at sm_86: include/gunrock/cuda/sm.hxx:23-39).
 * This build targets exactly one architecture (sm_100a), so inside the operators the "launch box" collapses to a
 * persistent-grid rule: min(needed CTAs, SMs x resident CTAs of that kernel from the occupancy API), grid-stride.
 * gcuda::launch_box (cuda/launch_box.hxx) keeps the reference's type for user kernels that name one.
 */
#pragma once

#include <cuda_runtime_api.h>
#include <gunrock/cuda/context.hxx>

namespace gunrock {
namespace gcuda {


namespace thread {
namespace global {
namespace id {
__device__ __forceinline__ std::size_t x() { return std::size_t(blockIdx.x) * blockDim.x + threadIdx.x; }
}  // namespace id
}  // namespace global
namespace local {
namespace id {
__device__ __forceinline__ unsigned x() { return threadIdx.x; }
}  // namespace id
}  // namespace local
}  // namespace thread
namespace block {
namespace id {
__device__ __forceinline__ unsigned x() { return blockIdx.x; }
}  // namespace id
namespace size {
__device__ __forceinline__ unsigned x() { return blockDim.x; }
}  // namespace size
}  // namespace block
namespace grid {
namespace size {
__device__ __forceinline__ unsigned x() { return gridDim.x; }
}  // namespace size
}  // namespace grid

/// CTAs for a grid-stride kernel: enough to cover `work_ctas`, capped at SMs x `resident`.
inline unsigned persistent_grid(standard_context_t& ctx, std::size_t work_ctas, int resident) {
  std::size_t cap = std::size_t(ctx.sm_count()) * std::size_t(resident);
  std::size_t g = work_ctas < cap ? work_ctas : cap;
  return unsigned(g ? g : 1);
}

/// Grid of a persistent kernel: every CTA the device can hold at once (SMs x occupancy of `kernel`), at most
/// `max_per_sm` per SM and never more than `work_ctas`.
template <typename kernel_t>
inline unsigned full_grid(standard_context_t& ctx, kernel_t kernel, std::size_t work_ctas = ~std::size_t(0),
                          int max_per_sm = 8, int threads = 256) {
  int per_sm = ctx.resident_ctas(kernel, threads);
  if (per_sm > max_per_sm) per_sm = max_per_sm;
  return persistent_grid(ctx, work_ctas, per_sm);
}

}  // namespace gcuda
}  // namespace gunrock
