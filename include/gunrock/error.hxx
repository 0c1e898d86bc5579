/**
 * @file error.hxx
 * @brief Exception convention of the operator API: CUDA status codes and boolean failures become a
 * C++ exception carrying a readable report. Mirrors the contract of the reference's
 * include/gunrock/error.hxx:21-46 (exception_t / throw_if_exception, both overloads).
 */
#pragma once

#include <cuda_runtime_api.h>
#include <exception>
#include <string>

namespace gunrock {
namespace error {

using error_t = cudaError_t;

class exception_t : public std::exception {
 public:
  std::string report;
  explicit exception_t(std::string message = "") : report(std::move(message)) {}
  exception_t(error_t status, const std::string& message = "")
      : report(std::string(cudaGetErrorString(status)) + "\t: " + message), code(status) {}
  const char* what() const noexcept override { return report.c_str(); }
  error_t status() const noexcept { return code; }

 private:
  error_t code = cudaErrorUnknown;
};

inline void throw_if_exception(error_t status, std::string message = "") {
  if (status != cudaSuccess)
    throw exception_t(status, message);
}

inline void throw_if_exception(bool is_exception, std::string message = "") {
  if (is_exception)
    throw exception_t(message);
}

/// Checks the launch/async error state; used after every kernel launch in the operators.
inline void check_last(const char* where) {
  throw_if_exception(cudaGetLastError(), where);
}

}  // namespace error
}  // namespace gunrock
