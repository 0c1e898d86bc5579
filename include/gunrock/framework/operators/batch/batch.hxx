/**
 * @file batch.hxx
 * @brief operators::batch::execute — run `number_of_jobs` independent jobs on as many host threads and sum the
 * milliseconds they return. Same signature as the reference (framework/operators/batch/batch.hxx:61-79); each
 * job is expected to build its own context (the operators keep no process-global mutable state; the device
 * memory pool is mutex-protected).
 */
#pragma once

#include <cstddef>
#include <thread>
#include <vector>
#include <cuda_runtime_api.h>

namespace gunrock {
namespace operators {
namespace batch {

template <typename function_t, typename... args_t>
void execute(function_t f, std::size_t number_of_jobs, float* total_elapsed, args_t&... args) {
  std::vector<float> elapsed(number_of_jobs, 0.f);
  std::vector<std::thread> workers;
  int device = 0;
  cudaGetDevice(&device);
  for (std::size_t job = 0; job < number_of_jobs; ++job)
    workers.emplace_back([&, job]() {
      cudaSetDevice(device);
      elapsed[job] = f(job);
    });
  float sum = 0.f;
  for (std::size_t job = 0; job < number_of_jobs; ++job) {
    workers[job].join();
    sum += elapsed[job];
  }
  if (total_elapsed) total_elapsed[0] = sum;
}

}  // namespace batch
}  // namespace operators
}  // namespace gunrock
