// TEST INFRASTRUCTURE ONLY — never linked into, imported by, or called from the product path.
//
// C-ABI wrapper around the REFERENCE's own GPU implementation, compiled for sm_100 from the sources where
// they lie under /root/reference (nothing copied): gunrock::bfs::run / sssp::run / pr::run / kcore::run /
// color::run / ppr::run (include/gunrock/algorithms/*.hxx) on the reference's operators
// (advance/block_mapped.hxx kernel + Thrust filters). moderngpu is only declared (oracle/mgpu_stub).
// Built by oracle/Makefile into oracle/_ref/libref_gpu.so; used to (a) check that our CUDA path returns the
// same arrays as the reference's GPU path on the same inputs and (b) report the "reference kernel on B200"
// time next to ours. Graph arrays are device pointers; int32 / float like the reference drivers.
#include <gunrock/algorithms/bfs.hxx>
#include <gunrock/algorithms/sssp.hxx>
#include <gunrock/algorithms/pr.hxx>
#include <gunrock/algorithms/ppr.hxx>
#include <gunrock/algorithms/kcore.hxx>
#include <gunrock/algorithms/color.hxx>

using namespace gunrock;
using namespace memory;

static auto make_graph(int n, int m, int* off, int* col, float* val) {
  return graph::build::from_csr<memory_space_t::device, graph::view_t::csr>(n, n, m, off, col, val);
}

#define REF_GUARD(body)                                    \
  try {                                                    \
    body                                                   \
  } catch (const std::exception& e) {                      \
    std::fprintf(stderr, "ref_gpu: %s\n", e.what());       \
    return -1.f;                                           \
  }

extern "C" {

float ref_gpu_bfs(int n, int m, int* d_off, int* d_col, float* d_val, int src, int* d_dist) {
  REF_GUARD(auto G = make_graph(n, m, d_off, d_col, d_val); thrust::device_vector<int> pred(1);
            return gunrock::bfs::run(G, src, d_dist, pred.data().get());)
}

float ref_gpu_sssp(int n, int m, int* d_off, int* d_col, float* d_val, int src, float* d_dist) {
  REF_GUARD(auto G = make_graph(n, m, d_off, d_col, d_val); thrust::device_vector<int> pred(1);
            return gunrock::sssp::run(G, src, d_dist, pred.data().get());)
}

float ref_gpu_pr(int n, int m, int* d_off, int* d_col, float* d_val, float alpha, float tol, float* d_p) {
  REF_GUARD(auto G = make_graph(n, m, d_off, d_col, d_val); return gunrock::pr::run(G, alpha, tol, d_p);)
}

float ref_gpu_ppr(int n, int m, int* d_off, int* d_col, float* d_val, int seed, float alpha, float eps, float* d_p) {
  REF_GUARD(auto G = make_graph(n, m, d_off, d_col, d_val); return gunrock::ppr::run(G, seed, d_p, alpha, eps);)
}

float ref_gpu_kcore(int n, int m, int* d_off, int* d_col, float* d_val, int* d_k) {
  REF_GUARD(auto G = make_graph(n, m, d_off, d_col, d_val); return gunrock::kcore::run(G, d_k);)
}

float ref_gpu_color(int n, int m, int* d_off, int* d_col, float* d_val, int* d_colors) {
  REF_GUARD(auto G = make_graph(n, m, d_off, d_col, d_val); return gunrock::color::run(G, d_colors);)
}

}  // extern "C"
