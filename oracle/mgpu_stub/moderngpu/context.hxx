// TEST INFRASTRUCTURE ONLY. Declaration-only stand-in for moderngpu (https://github.com/moderngpu/moderngpu,
// which the reference fetches at build time, unpinned `master`, cmake/FetchModernGPU.cmake:9-13, and which is
// absent from /root/reference). It lets the reference's OWN GPU code for bfs/sssp/pr/ppr/kcore/color compile
// for the "reference GPU" baseline (oracle/ref_gpu_shim.cu): those algorithms only instantiate the hand-written
// block_mapped advance and Thrust filters, so no moderngpu function is ever called or linked. Nothing of
// moderngpu is reproduced here: only the names the reference headers mention.
#pragma once
#include <cuda_runtime_api.h>
namespace mgpu {
struct standard_context_t {
  standard_context_t(bool, cudaStream_t) {}
};
template <typename... args_t>
void transform_lbs(args_t&&...);
template <typename... args_t>
void transform_segreduce(args_t&&...);
template <typename... args_t>
void lbs_segreduce(args_t&&...);
struct compact_stub_t {
  template <typename f_t>
  int upsweep(f_t);
  template <typename f_t>
  void downsweep(f_t);
};
template <typename... args_t>
compact_stub_t transform_compact(args_t&&...);
}  // namespace mgpu
