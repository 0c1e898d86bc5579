/**
 * @file timer.hxx
 * @brief cudaEvent stopwatch returned by standard_context_t::timer(); enact() reports its milliseconds
 * (reference: include/gunrock/util/timer.hxx:17-50, which records on the legacy default stream).
 * Here the two events are recorded on the context's own stream, so the interval brackets exactly the work the
 * operators enqueue, and they are created on first use (a context that never times anything owns no events).
 */
#pragma once

#include <cuda_runtime_api.h>

namespace gunrock {
namespace util {

struct timer_t {
  float time = 0.f;  ///< last measured interval, milliseconds

  explicit timer_t(cudaStream_t stream = 0) : stream_(stream) {}
  timer_t(const timer_t&) = delete;
  timer_t& operator=(const timer_t&) = delete;
  ~timer_t() {
    for (cudaEvent_t& e : mark_)
      if (e) cudaEventDestroy(e);
  }

  void set_stream(cudaStream_t stream) { stream_ = stream; }

  /// begin()/start() drop the opening mark, end()/stop() the closing one and wait for it.
  void begin() { drop(opening); }
  void start() { drop(opening); }
  float end() { return close(); }
  float stop() { return close(); }

  float milliseconds() const { return time; }
  float seconds() const { return time / 1000.f; }

 private:
  enum which_t { opening = 0, closing = 1 };
  void drop(which_t which) {
    if (!mark_[which]) cudaEventCreate(&mark_[which]);
    cudaEventRecord(mark_[which], stream_);
  }
  float close() {
    if (!mark_[opening]) return time = 0.f;  // never started
    drop(closing);
    cudaEventSynchronize(mark_[closing]);
    cudaEventElapsedTime(&time, mark_[opening], mark_[closing]);
    return time;
  }
  cudaEvent_t mark_[2] = {nullptr, nullptr};
  cudaStream_t stream_;
};

}  // namespace util
}  // namespace gunrock
