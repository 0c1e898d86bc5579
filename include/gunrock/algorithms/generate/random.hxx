/**
 * @file random.hxx
 * @brief generate::random::uniform_distribution — the seed-free per-index random stream colouring uses.
 * Same sequence as the reference (include/gunrock/algorithms/generate/random.hxx:20-33): element i is
 * thrust::default_random_engine (minstd_rand, x <- 48271 x mod 2^31-1, seed 1) after discard(i), mapped by
 * thrust::uniform_real_distribution<float>(begin, end). Computed here by modular exponentiation
 * (48271^(i+1) mod 2^31-1) in a plain kernel; the float mapping uses round-to-nearest intrinsics so no FMA
 * contraction can make the device stream differ from the host/oracle stream.
 */
#pragma once

#include <cstdint>
#include <gunrock/memory.hxx>

namespace gunrock {
namespace generate {
namespace random {

__host__ __device__ inline std::uint32_t minstd_power(std::uint64_t exponent) {
  const std::uint64_t M = 2147483647ull;
  std::uint64_t base = 48271ull, acc = 1ull;
  while (exponent) {
    if (exponent & 1ull) acc = (acc * base) % M;
    base = (base * base) % M;
    exponent >>= 1;
  }
  return std::uint32_t(acc);
}

__host__ __device__ inline float minstd_uniform(std::uint64_t i, float begin, float end) {
  const float raw = float(minstd_power(i + 1) - 1u);
#ifdef __CUDA_ARCH__
  const float unit = __fdiv_rn(raw, __fadd_rn(1.0f, float(2147483646u - 1u)));
  return __fadd_rn(__fmul_rn(unit, __fsub_rn(end, begin)), begin);
#else
  const float unit = raw / (1.0f + float(2147483646u - 1u));
  volatile float scaled = unit * (end - begin);  // volatile: keep the host from fusing mul+add
  return scaled + begin;
#endif
}

namespace kernels {
static __global__ void __launch_bounds__(256) uniform_kernel(float* out, std::size_t n, float begin, float end) {
  for (std::size_t i = std::size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += std::size_t(gridDim.x) * blockDim.x)
    out[i] = minstd_uniform(i, begin, end);
}
}  // namespace kernels

/// Fills a device array of floats (device_array_t or raw pointer + size).
inline void uniform_distribution(float* d_out, std::size_t n, float begin = 0.0f, float end = 1.0f,
                                 cudaStream_t stream = 0) {
  if (n) kernels::uniform_kernel<<<unsigned((n + 255) / 256 < 1184 ? (n + 255) / 256 : 1184), 256, 0, stream>>>(d_out, n, begin, end);
}

template <typename vector_t>
void uniform_distribution(vector_t& input, float begin = 0.0f, float end = 1.0f) {
  uniform_distribution(memory::raw_pointer_cast(input.data()), input.size(), begin, end);
}

}  // namespace random
}  // namespace generate
}  // namespace gunrock
