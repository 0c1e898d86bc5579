// TEST INFRASTRUCTURE ONLY — never linked into, imported by, or called from the product path.
//
// C-ABI wrapper around the REFERENCE's own single-threaded CPU implementations, compiled from the
// sources where they lie under /root/reference (nothing is copied into this repo):
//   examples/algorithms/bfs/bfs_cpu.hxx:21-68      -> ref_bfs
//   examples/algorithms/sssp/sssp_cpu.hxx:23-72    -> ref_sssp
//   examples/algorithms/kcore/kcore_cpu.hxx:8-62   -> ref_kcore
//   examples/algorithms/ppr/ppr_cpu.hxx:16-98      -> ref_ppr
//   examples/algorithms/color/color_cpu.hxx:16-72  -> ref_color (Gauss-Seidel; validity reference only)
//   include/gunrock/algorithms/generate/random.hxx:20-33 -> ref_randoms
//   include/gunrock/io/matrix_market.hxx:99-240 + formats/csr.hxx:79-157 -> ref_load_mtx
//   include/gunrock/formats/csr.hxx:159-236 (read_binary / write_binary) -> ref_read_csr_binary / ref_write_csr_binary
// Built by oracle/Makefile into oracle/_ref/libref_cpu.so with Thrust's CPP (host) backend, so that
// `memory_space_t::device` vectors are plain host memory and no GPU is needed.
// Used (a) to pin oracle/oracle.cpp against the reference, (b) to generate tests/golden/*, and
// (c) as bench.py's `cpu_baseline` / `--impl reference` arm ("kind": "reference").
#include <limits>
#include <cstring>
#include <cstdio>
#include <thrust/host_vector.h>
#include <thrust/device_vector.h>
#include <thrust/fill.h>

#include <gunrock/memory.hxx>
#include <gunrock/error.hxx>
#include <gunrock/container/vector.hxx>
#include <gunrock/formats/formats.hxx>
#include <gunrock/io/matrix_market.hxx>
#include <gunrock/algorithms/generate/random.hxx>

#include "bfs/bfs_cpu.hxx"
#include "sssp/sssp_cpu.hxx"
#include "kcore/kcore_cpu.hxx"
#include "ppr/ppr_cpu.hxx"
#include "color/color_cpu.hxx"

using namespace gunrock;
using namespace gunrock::memory;

using vertex_t = int;
using edge_t = int;
using weight_t = float;
using ref_csr_t = format::csr_t<memory_space_t::device, vertex_t, edge_t, weight_t>;

static ref_csr_t make_csr(int n, int m, const int* off, const int* col, const float* val) {
  ref_csr_t csr(n, n, m);
  for (int i = 0; i <= n; ++i) csr.row_offsets[i] = off[i];
  for (int e = 0; e < m; ++e) {
    csr.column_indices[e] = col[e];
    csr.nonzero_values[e] = val ? val[e] : 1.0f;
  }
  return csr;
}

extern "C" {

// Loads a MatrixMarket file with the reference loader and converts COO->CSR with the reference
// converter. Arrays are malloc'ed; free with ref_free.
int ref_load_mtx(const char* path, int* n, int* m, int** off, int** col, float** val) {
  io::matrix_market_t<vertex_t, edge_t, weight_t> mm;
  ref_csr_t csr;
  csr.from_coo(mm.load(path));
  *n = csr.number_of_rows;
  *m = csr.number_of_nonzeros;
  *off = (int*)malloc(sizeof(int) * (*n + 1));
  *col = (int*)malloc(sizeof(int) * (*m));
  *val = (float*)malloc(sizeof(float) * (*m));
  for (int i = 0; i <= *n; ++i) (*off)[i] = csr.row_offsets[i];
  for (int e = 0; e < *m; ++e) {
    (*col)[e] = csr.column_indices[e];
    (*val)[e] = csr.nonzero_values[e];
  }
  return 0;
}

void ref_free(void* p) { free(p); }

float ref_bfs(int n, int m, const int* off, const int* col, int src, int* dist) {
  auto csr = make_csr(n, m, off, col, nullptr);
  std::vector<int> pred(n);
  return bfs_cpu::run<ref_csr_t, vertex_t, edge_t>(csr, src, dist, pred.data());
}

// Persistent form for bench.py's reference arm: the graph is wrapped ONCE (make_csr copies 8 bytes per edge) and the
// reference's bfs_cpu::run / sssp_cpu::run are called on it from several host threads, one source each (they only
// read the csr_t; each call still makes the reference's own private host copy, bfs_cpu.hxx:25-27).
void* ref_graph_create(int n, int m, const int* off, const int* col, const float* val) {
  return new ref_csr_t(make_csr(n, m, off, col, val));
}
void ref_graph_destroy(void* g) { delete static_cast<ref_csr_t*>(g); }
float ref_bfs_on(void* g, int src, int* dist) {
  auto& csr = *static_cast<ref_csr_t*>(g);
  std::vector<int> pred(csr.number_of_rows);
  return bfs_cpu::run<ref_csr_t, vertex_t, edge_t>(csr, src, dist, pred.data());
}
float ref_sssp_on(void* g, int src, float* dist) {
  auto& csr = *static_cast<ref_csr_t*>(g);
  std::vector<int> pred(csr.number_of_rows);
  return sssp_cpu::run<ref_csr_t, vertex_t, edge_t, weight_t>(csr, src, dist, pred.data());
}

float ref_sssp(int n, int m, const int* off, const int* col, const float* val, int src, float* dist) {
  auto csr = make_csr(n, m, off, col, val);
  std::vector<int> pred(n);
  return sssp_cpu::run<ref_csr_t, vertex_t, edge_t, weight_t>(csr, src, dist, pred.data());
}

float ref_kcore(int n, int m, const int* off, const int* col, int* k_cores) {
  auto csr = make_csr(n, m, off, col, nullptr);
  return kcore_cpu::run<ref_csr_t, vertex_t, edge_t, weight_t>(csr, k_cores);
}

// p is n_seeds*n floats, zero-initialised by the caller exactly like examples/algorithms/ppr/ppr.cu.
float ref_ppr(int n, int m, const int* off, const int* col, int n_seeds, float* p, float alpha, float eps) {
  auto csr = make_csr(n, m, off, col, nullptr);
  return ppr_cpu::run<ref_csr_t, vertex_t, edge_t, weight_t>(csr, n_seeds, p, alpha, eps);
}

float ref_color(int n, int m, const int* off, const int* col, int* colors) {
  auto csr = make_csr(n, m, off, col, nullptr);
  return color_cpu::run<ref_csr_t, vertex_t, edge_t, weight_t>(csr, colors);
}

// The reference's `.csr` binary reader / writer (pins essentials_b200/io.py). Arrays from the reader are
// malloc'ed; free with ref_free.
int ref_write_csr_binary(const char* path, int n, int m, const int* off, const int* col, const float* val) {
  auto csr = make_csr(n, m, off, col, val);
  csr.write_binary(path);
  return 0;
}
int ref_read_csr_binary(const char* path, int* n, int* m, int** off, int** col, float** val) {
  ref_csr_t csr;
  csr.read_binary(path);
  *n = csr.number_of_rows;
  *m = csr.number_of_nonzeros;
  *off = (int*)malloc(sizeof(int) * (*n + 1));
  *col = (int*)malloc(sizeof(int) * (*m));
  *val = (float*)malloc(sizeof(float) * (*m));
  for (int i = 0; i <= *n; ++i) (*off)[i] = csr.row_offsets[i];
  for (int e = 0; e < *m; ++e) {
    (*col)[e] = csr.column_indices[e];
    (*val)[e] = csr.nonzero_values[e];
  }
  return 0;
}

// The reference's seed-free per-index random stream (color.hxx:65 calls it with (0, n)).
void ref_randoms(int n, float lo, float hi, float* out) {
  thrust::host_vector<float> r(n);
  generate::random::uniform_distribution(r, lo, hi);
  for (int i = 0; i < n; ++i) out[i] = r[i];
}

}  // extern "C"
