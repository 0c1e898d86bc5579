/**
 * @file timer.hxx
 * @brief cudaEvent stopwatch returned by standard_context_t::timer(); enact() reports its milliseconds
 * (reference: include/gunrock/util/timer.hxx:17-50, which records on the legacy default stream).
 * Here the events are recorded on the context's own stream so the measured interval brackets exactly
 * the work the operators enqueue.
 */
#pragma once

#include <cuda_runtime_api.h>

namespace gunrock {
namespace util {

struct timer_t {
  float time = 0.f;

  explicit timer_t(cudaStream_t stream = 0) : stream_(stream) {
    cudaEventCreate(&start_);
    cudaEventCreate(&stop_);
  }
  timer_t(const timer_t&) = delete;
  timer_t& operator=(const timer_t&) = delete;
  ~timer_t() {
    cudaEventDestroy(start_);
    cudaEventDestroy(stop_);
  }

  void set_stream(cudaStream_t s) { stream_ = s; }
  void begin() { cudaEventRecord(start_, stream_); }
  void start() { begin(); }
  float end() {
    cudaEventRecord(stop_, stream_);
    cudaEventSynchronize(stop_);
    cudaEventElapsedTime(&time, start_, stop_);
    return time;
  }
  float stop() { return end(); }
  float seconds() const { return time * 1e-3f; }
  float milliseconds() const { return time; }

 private:
  cudaEvent_t start_{}, stop_{};
  cudaStream_t stream_;
};

}  // namespace util
}  // namespace gunrock
