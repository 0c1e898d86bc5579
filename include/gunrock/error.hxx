/**
 * @file error.hxx
 * @brief Exception convention of the operator API: CUDA status codes and boolean failures become a
 * C++ exception carrying a readable report. Keeps the contract of the reference's
 * include/gunrock/error.hxx:21-46 — error::error_t, error::exception_t with a public `report`, and the two
 * throw_if_exception overloads — and adds the status code (the C ABI returns it) and a post-launch check.
 */
#pragma once

#include <cuda_runtime_api.h>
#include <exception>
#include <string>
#include <utility>

namespace gunrock {
namespace error {

using error_t = cudaError_t;

class exception_t : public std::exception {
 public:
  std::string report;  ///< what() text; public because reference-side code prints it directly

  explicit exception_t(std::string message = "") : report(std::move(message)), code_(cudaErrorUnknown) {}
  exception_t(error_t status, const std::string& message = "") : report(describe(status, message)), code_(status) {}

  const char* what() const noexcept override { return report.c_str(); }
  error_t status() const noexcept { return code_; }

 private:
  static std::string describe(error_t status, const std::string& message) {
    std::string text(cudaGetErrorString(status));
    text += "\t: ";
    text += message;
    return text;
  }
  error_t code_;
};

namespace detail {
template <typename... args_t>
[[noreturn]] inline void raise(args_t&&... args) {
  throw exception_t(std::forward<args_t>(args)...);
}
}  // namespace detail

/// Throws when `status` is not cudaSuccess; the report is "<cuda error string>\t: <message>".
inline void throw_if_exception(error_t status, std::string message = "") {
  if (status == cudaSuccess) return;
  detail::raise(status, message);
}
/// Throws when the condition holds (argument checks, unsupported template combinations).
inline void throw_if_exception(bool is_exception, std::string message = "") {
  if (!is_exception) return;
  detail::raise(std::move(message));
}

/// Launch/async error state of the calling thread; the operators call it after every kernel launch.
inline void check_last(const char* where) { throw_if_exception(cudaGetLastError(), where); }

}  // namespace error
}  // namespace gunrock
