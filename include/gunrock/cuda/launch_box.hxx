/**
 * @file launch_box.hxx
 * @brief gcuda::launch_box — compile-time launch configurations keyed by SM architecture, for user code that names
 * a launch box (reference: include/gunrock/cuda/launch_box.hxx:30-335, sm.hxx:23-39, detail/launch_kernels.hxx:20-55).
 *
 * Same spellings as the reference: `dim3_t<x,y,z>`, `dimensions_t`, `launch_params_t<flags, block, grid, items, smem>`,
 * `launch_params_dynamic_grid_t<flags, block, items, smem>` with calculate_grid_dimensions_{strided,blocked},
 * `launch_box_t<params...>` with launch / launch_strided / launch_blocked / launch_cooperative, and the `sm_flag_t`
 * names up to sm_86 so existing parameter packs keep compiling. This build targets exactly one architecture
 * (sm_100a), so the selection rule is: the first parameter set whose flags contain `sm_100` — every `fallback` set does —
 * and a static_assert when there is none. The operators of this tree do not use launch boxes (they size persistent
 * grids from the occupancy API, cuda/launch.hxx); the type exists for drop-in compatibility of user kernels.
 */
#pragma once

#include <cstddef>
#include <tuple>
#include <type_traits>
#include <utility>

#include <gunrock/cuda/context.hxx>
#include <gunrock/cuda/launch.hxx>

namespace gunrock {
namespace gcuda {
namespace launch_box {

/// Architecture bit flags (reference cuda/sm.hxx:23-39, extended to Hopper / Blackwell). `fallback` matches every SM.
enum sm_flag_t : unsigned {
  fallback = ~0u,
  sm_30 = 1u << 0,
  sm_35 = 1u << 1,
  sm_37 = 1u << 2,
  sm_50 = 1u << 3,
  sm_52 = 1u << 4,
  sm_53 = 1u << 5,
  sm_60 = 1u << 6,
  sm_61 = 1u << 7,
  sm_62 = 1u << 8,
  sm_70 = 1u << 9,
  sm_72 = 1u << 10,
  sm_75 = 1u << 11,
  sm_80 = 1u << 12,
  sm_86 = 1u << 13,
  sm_89 = 1u << 14,
  sm_90 = 1u << 15,
  sm_100 = 1u << 16
};
constexpr sm_flag_t operator|(sm_flag_t a, sm_flag_t b) { return sm_flag_t(unsigned(a) | unsigned(b)); }
constexpr sm_flag_t operator&(sm_flag_t a, sm_flag_t b) { return sm_flag_t(unsigned(a) & unsigned(b)); }

/// The architecture this tree is compiled for.
constexpr sm_flag_t target_sm = sm_100;

struct dimensions_t {
  unsigned int x, y, z;
  __host__ __device__ constexpr dimensions_t(unsigned int _x = 1, unsigned int _y = 1, unsigned int _z = 1)
      : x(_x), y(_y), z(_z) {}
  __host__ __device__ constexpr unsigned int size() const { return x * y * z; }
  __host__ __device__ operator dim3() const { return dim3(x, y, z); }
};

/// dim3 as a type (a dim3 value cannot be a template argument).
template <unsigned int x_ = 1, unsigned int y_ = 1, unsigned int z_ = 1>
struct dim3_t {
  enum : unsigned int { x = x_, y = y_, z = z_ };
  static constexpr unsigned int size() { return x_ * y_ * z_; }
  static constexpr dimensions_t dimensions() { return dimensions_t(x_, y_, z_); }
  constexpr operator dimensions_t() const { return dimensions(); }
};

namespace detail {
template <sm_flag_t flags, std::size_t items, std::size_t smem>
struct launch_params_base_t {
  static constexpr sm_flag_t sm_flags = flags;
  static constexpr std::size_t items_per_thread = items;
  static constexpr std::size_t shared_memory_bytes = smem;
};

template <typename T>
struct dependent_false : std::false_type {};
template <typename T>
struct no_matching_launch_params_t {
  static_assert(dependent_false<T>::value, "Launch box could not find valid launch parameters");
};

/// First parameter set of the pack whose flags contain the target architecture.
template <typename... lp_v>
struct first_match;
template <>
struct first_match<> {
  using type = no_matching_launch_params_t<void>;
};
template <typename lp, typename... rest>
struct first_match<lp, rest...> {
  using type = std::conditional_t<(unsigned(lp::sm_flags) & unsigned(target_sm)) != 0, lp,
                                  typename first_match<rest...>::type>;
};
}  // namespace detail

/// Launch parameters with a compile-time grid.
template <sm_flag_t flags, typename block_dimensions_, typename grid_dimensions_, std::size_t items_per_thread_ = 1,
          std::size_t shared_memory_bytes_ = 0>
struct launch_params_t : detail::launch_params_base_t<flags, items_per_thread_, shared_memory_bytes_> {
  using base_t = detail::launch_params_base_t<flags, items_per_thread_, shared_memory_bytes_>;
  using block_dimensions_t = block_dimensions_;
  using grid_dimensions_t = grid_dimensions_;
  static constexpr dimensions_t block_dimensions = block_dimensions_t::dimensions();
  static constexpr dimensions_t grid_dimensions = grid_dimensions_t::dimensions();
  void calculate_grid_dimensions_strided(std::size_t) {}  // fixed grid
  void calculate_grid_dimensions_blocked(std::size_t) {}
};

/// Launch parameters whose grid is computed from the number of elements at run time.
template <sm_flag_t flags, typename block_dimensions_, std::size_t items_per_thread_ = 1,
          std::size_t shared_memory_bytes_ = 0>
struct launch_params_dynamic_grid_t : detail::launch_params_base_t<flags, items_per_thread_, shared_memory_bytes_> {
  using base_t = detail::launch_params_base_t<flags, items_per_thread_, shared_memory_bytes_>;
  using block_dimensions_t = block_dimensions_;
  static constexpr dimensions_t block_dimensions = block_dimensions_t::dimensions();
  dimensions_t grid_dimensions;
  /// One thread per element.
  void calculate_grid_dimensions_strided(std::size_t num_elements) {
    const std::size_t per_cta = block_dimensions.x;
    grid_dimensions = dimensions_t(unsigned((num_elements + per_cta - 1) / per_cta), 1, 1);
  }
  /// items_per_thread elements per thread.
  void calculate_grid_dimensions_blocked(std::size_t num_elements) {
    const std::size_t per_cta = std::size_t(block_dimensions.x) * base_t::items_per_thread;
    grid_dimensions = dimensions_t(unsigned((num_elements + per_cta - 1) / per_cta), 1, 1);
  }
};

template <typename... lp_v>
using select_launch_params_t = typename detail::first_match<lp_v...>::type;

namespace kernels {
/// f(index, block id, args...) for every index below `bound`, grid-stride.
template <unsigned int threads_per_block, typename func_t, typename... args_t>
__global__ void __launch_bounds__(threads_per_block) strided_kernel(func_t f, const std::size_t bound, args_t... args) {
  const std::size_t stride = std::size_t(blockDim.x) * gridDim.x;
  for (std::size_t i = std::size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < bound; i += stride)
    f(int(i), int(blockIdx.x), args...);
}
/// The same with `items_per_thread` grid-strided elements per thread and trip.
template <unsigned int threads_per_block, unsigned int items_per_thread, typename func_t, typename... args_t>
__global__ void __launch_bounds__(threads_per_block) blocked_kernel(func_t f, const std::size_t bound, args_t... args) {
  const std::size_t stride = std::size_t(blockDim.x) * gridDim.x;
  for (std::size_t i = std::size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < bound; i += stride * items_per_thread) {
#pragma unroll
    for (unsigned int j = 0; j < items_per_thread; ++j)
      if (i + stride * j < bound) f(int(i + stride * j), int(blockIdx.x), args...);
  }
}
}  // namespace kernels

/**
 * @brief A pack of launch-parameter sets; the box IS the first set that matches the architecture compiled for.
 * launch(context, kernel, args...) launches `kernel` with the box's grid / block / shared memory on the context's
 * stream; launch_strided / launch_blocked wrap a functor f(index, block id, args...) in a grid-stride kernel.
 */
template <typename... lp_v>
struct launch_box_t : public select_launch_params_t<lp_v...> {
  using params_t = select_launch_params_t<lp_v...>;
  launch_box_t() {}

  template <typename func_t, typename... args_t>
  void launch_strided(gcuda::standard_context_t& context, func_t& f, const std::size_t num_elements, args_t&&... args) {
    params_t::calculate_grid_dimensions_strided(num_elements);
    kernels::strided_kernel<params_t::block_dimensions_t::size()>
        <<<dim3(this->grid_dimensions), dim3(params_t::block_dimensions), params_t::shared_memory_bytes,
           context.stream()>>>(f, num_elements, std::forward<args_t>(args)...);
  }

  template <typename func_t, typename... args_t>
  void launch_blocked(gcuda::standard_context_t& context, func_t& f, const std::size_t num_elements, args_t&&... args) {
    params_t::calculate_grid_dimensions_blocked(num_elements);
    kernels::blocked_kernel<params_t::block_dimensions_t::size(), unsigned(params_t::items_per_thread)>
        <<<dim3(this->grid_dimensions), dim3(params_t::block_dimensions), params_t::shared_memory_bytes,
           context.stream()>>>(f, num_elements, std::forward<args_t>(args)...);
  }

  template <typename func_t, typename... args_t>
  void launch_cooperative(gcuda::standard_context_t& context, const func_t& f, const std::size_t num_elements,
                          args_t&&... args) {
    params_t::calculate_grid_dimensions_strided(num_elements);
    void* pointers[sizeof...(args_t) == 0 ? 1 : sizeof...(args_t)] = {
        const_cast<void*>(static_cast<const void*>(&args))...};
    cudaLaunchCooperativeKernel((const void*)f, dim3(this->grid_dimensions), dim3(params_t::block_dimensions), pointers,
                                params_t::shared_memory_bytes, context.stream());
  }

  template <typename func_t, typename... args_t>
  void launch(gcuda::standard_context_t& context, const func_t& f, args_t&&... args) {
    f<<<dim3(this->grid_dimensions), dim3(params_t::block_dimensions), params_t::shared_memory_bytes,
        context.stream()>>>(std::forward<args_t>(args)...);
  }
};

/// Ratio of resident to maximum warps per SM for `kernel` under this box's block size (occupancy API).
template <typename launch_box_type, typename func_t>
inline float occupancy(func_t kernel) {
  int max_active_blocks = 0, device = 0;
  cudaGetDevice(&device);
  cudaDeviceProp props;
  cudaGetDeviceProperties(&props, device);
  const int block_size = int(launch_box_type::block_dimensions.size());
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&max_active_blocks, kernel, block_size,
                                                launch_box_type::shared_memory_bytes);
  return float(max_active_blocks * block_size) / float(props.maxThreadsPerMultiProcessor);
}

}  // namespace launch_box
}  // namespace gcuda
}  // namespace gunrock
