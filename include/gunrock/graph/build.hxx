/**
 * @file build.hxx
 * @brief graph::build::from_csr — wraps caller-owned CSR arrays into a graph_t and, when asked, derives
 * the COO row indices or the CSC (transpose) into caller-provided buffers.
 *
 * Signature kept from the reference (include/gunrock/graph/build.hxx:21-36):
 *   from_csr<space, views>(rows, cols, nnz, Ap, J, X, I = nullptr, Aj = nullptr)
 * with I = row_indices[nnz] and Aj = column_offsets[rows+1]. Unlike the reference
 * (graph/detail/build.hxx:85-111) CSR|CSC together is supported and the CSR arrays are never modified:
 * the transpose is a counting sort by column (histogram -> scan -> cursor scatter) that fills I and Aj and,
 * for weighted graphs, a separate `csc_values` array (extra trailing parameter). Passing I == J and
 * Aj == Ap declares the graph symmetric: the CSC view then aliases the CSR arrays (exact for undirected
 * graphs, zero cost). Host-space graphs are wrapped as-is.
 * Setup code: runs once per graph, outside every timed region.
 */
#pragma once

#include <type_traits>
#include <cuda_runtime_api.h>
#include <gunrock/error.hxx>
#include <gunrock/memory.hxx>

namespace gunrock {
namespace graph {
namespace build {
namespace detail {

template <typename edge_t>
__device__ __forceinline__ edge_t bump(edge_t* p) {
  if constexpr (sizeof(edge_t) == 8)
    return edge_t(atomicAdd(reinterpret_cast<unsigned long long*>(p), 1ull));
  else
    return edge_t(atomicAdd(reinterpret_cast<unsigned int*>(p), 1u));
}

template <typename vertex_t, typename edge_t>
__global__ void __launch_bounds__(256) rows_of_edges_kernel(vertex_t n, const edge_t* __restrict__ offsets,
                                                           vertex_t* __restrict__ row_of_edge) {
  const unsigned lane = threadIdx.x & 31;
  const std::size_t warps = (std::size_t(gridDim.x) * blockDim.x) >> 5;
  for (std::size_t row = (std::size_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5; row < std::size_t(n); row += warps) {
    edge_t b = offsets[row], e = offsets[row + 1];
    for (edge_t k = b + lane; k < e; k += 32) row_of_edge[k] = vertex_t(row);
  }
}

template <typename vertex_t, typename edge_t>
__global__ void __launch_bounds__(256) column_histogram_kernel(edge_t m, const vertex_t* __restrict__ columns,
                                                              edge_t* __restrict__ counts_shifted) {
  for (std::size_t e = std::size_t(blockIdx.x) * blockDim.x + threadIdx.x; e < std::size_t(m);
       e += std::size_t(gridDim.x) * blockDim.x)
    bump(&counts_shifted[columns[e] + 1]);
}

/// In-place inclusive scan of a[0..count) by ONE CTA walking 1024-element chunks (setup only).
template <typename edge_t>
__global__ void __launch_bounds__(1024) single_cta_inclusive_scan_kernel(edge_t* a, std::size_t count) {
  __shared__ edge_t warp_sums[33];
  __shared__ edge_t carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (std::size_t base = 0; base < count; base += 1024) {
    std::size_t i = base + threadIdx.x;
    edge_t x = i < count ? a[i] : edge_t(0);
    for (int d = 1; d < 32; d <<= 1) {
      edge_t y = __shfl_up_sync(0xffffffffu, x, d);
      if (lane >= unsigned(d)) x += y;
    }
    if (lane == 31) warp_sums[warp] = x;
    __syncthreads();
    if (warp == 0) {
      edge_t w = warp_sums[lane];
      for (int d = 1; d < 32; d <<= 1) {
        edge_t y = __shfl_up_sync(0xffffffffu, w, d);
        if (lane >= unsigned(d)) w += y;
      }
      warp_sums[lane] = w;
    }
    __syncthreads();
    edge_t before = carry + (warp ? warp_sums[warp - 1] : edge_t(0));
    if (i < count) a[i] = x + before;
    __syncthreads();
    if (threadIdx.x == 1023) carry = x + before;
    __syncthreads();
  }
}

template <typename vertex_t, typename edge_t, typename weight_t>
__global__ void __launch_bounds__(256)
    transpose_scatter_kernel(vertex_t n, const edge_t* __restrict__ offsets, const vertex_t* __restrict__ columns,
                             const weight_t* __restrict__ values, edge_t* __restrict__ cursor,
                             vertex_t* __restrict__ row_indices, weight_t* __restrict__ csc_values) {
  const unsigned lane = threadIdx.x & 31;
  const std::size_t warps = (std::size_t(gridDim.x) * blockDim.x) >> 5;
  for (std::size_t row = (std::size_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5; row < std::size_t(n); row += warps) {
    edge_t b = offsets[row], e = offsets[row + 1];
    for (edge_t k = b + lane; k < e; k += 32) {
      edge_t at = bump(&cursor[columns[k]]);
      row_indices[at] = vertex_t(row);
      if (csc_values) csc_values[at] = values ? values[k] : weight_t(1);
    }
  }
}

/// head[v] = in-neighbour of v with the largest degree (ties: first in the list), head_edge[v] = that in-edge;
/// head[v] = -1 for vertices without in-edges, and -2 - u for vertices whose single in-edge comes from u. `degree_offsets` is the offsets array degrees are read from
/// (the CSR when present, so "degree" is the out-degree of the in-neighbour). One lane per vertex, warp for long lists.
template <typename vertex_t, typename edge_t>
__global__ void __launch_bounds__(256)
    pull_hints_kernel(vertex_t first, vertex_t n, const edge_t* __restrict__ in_offsets,
                      const vertex_t* __restrict__ in_indices,
                      const edge_t* __restrict__ degree_offsets, const vertex_t* __restrict__ degree_of,
                      vertex_t* __restrict__ head, edge_t* __restrict__ head_edge) {
  // One lane per vertex: lists of up to 16 in-edges (most vertices of a power-law graph) are scanned by their own
  // lane, longer ones by the whole warp in coalesced strides (a warp per vertex, the first version, spent 34 ms at
  // scale-26 mostly on degree-1/2 vertices with one busy lane). Arg-max rule in both paths: largest degree, earliest edge.
  const unsigned lane = threadIdx.x & 31;
  const std::size_t warps = (std::size_t(gridDim.x) * blockDim.x) >> 5;
  auto degree = [&](vertex_t u) -> long long {
    return degree_of ? (long long)degree_of[u] : (long long)(degree_offsets[u + 1] - degree_offsets[u]);
  };
  // vertices [first, n): the streamed build (pull_hints_range) calls this once per arrived chunk of the indices
  for (std::size_t base = std::size_t(first) + ((std::size_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5) * 32;
       base < std::size_t(n); base += warps * 32) {
    const std::size_t v = base + lane;
    edge_t b = 0, e = 0;
    if (v < std::size_t(n)) {
      b = in_offsets[v];
      e = in_offsets[v + 1];
    }
    long long best_deg = -1;
    edge_t best_edge = e;
    const bool is_long = e - b > edge_t(16);
    if (!is_long)
      for (edge_t k = b; k < e; ++k) {
        const long long d = degree(in_indices[k]);
        if (d > best_deg) {  // strict: keeps the earliest edge on ties
          best_deg = d;
          best_edge = k;
        }
      }
    unsigned long_lanes = __ballot_sync(0xffffffffu, is_long);
    while (long_lanes) {
      const int owner = __ffs(long_lanes) - 1;
      long_lanes &= long_lanes - 1;
      const edge_t ob = __shfl_sync(0xffffffffu, b, owner), oe = __shfl_sync(0xffffffffu, e, owner);
      long long bd = -1;
      edge_t be = oe;
      for (edge_t k = ob + edge_t(lane); k < oe; k += 32) {
        const long long d = degree(in_indices[k]);
        if (d > bd) {
          bd = d;
          be = k;
        }
      }
      for (int s = 16; s > 0; s >>= 1) {  // (degree desc, edge asc) arg-max across the warp
        const long long od = __shfl_xor_sync(0xffffffffu, bd, s);
        const edge_t oe2 = __shfl_xor_sync(0xffffffffu, be, s);
        if (od > bd || (od == bd && oe2 < be)) {
          bd = od;
          be = oe2;
        }
      }
      if (int(lane) == owner) {
        best_deg = bd;
        best_edge = be;
      }
    }
    if (v < std::size_t(n)) {
      // >= 0: the hint; -1: no in-edges; <= -2: the ONLY in-neighbour is -2 - head[v] (nothing to walk on a miss)
      vertex_t h = vertex_t(-1);
      if (best_deg >= 0) {
        h = in_indices[best_edge];
        if (e - b == 1) h = vertex_t(-2) - h;
      }
      head[v] = h;
      head_edge[v] = best_deg >= 0 ? best_edge : edge_t(-1);
    }
  }
}

/// degree[v] = offsets[v+1] - offsets[v] as vertex_t (one coalesced pass): the hint kernel then needs ONE random
/// 4-byte gather per in-edge instead of two row-bound gathers (it is bound by the divergent-gather rate: 2.1 G edges
/// at scale-26, 34 ms with row bounds).
template <typename vertex_t, typename edge_t>
__global__ void __launch_bounds__(256)
    degrees_kernel(vertex_t n, const edge_t* __restrict__ offsets, vertex_t* __restrict__ degree) {
  for (std::size_t v = std::size_t(blockIdx.x) * blockDim.x + threadIdx.x; v < std::size_t(n);
       v += std::size_t(gridDim.x) * blockDim.x) {
    const edge_t d = offsets[v + 1] - offsets[v];
    degree[v] = d > edge_t(0x7fffffff) ? vertex_t(0x7fffffff) : vertex_t(d);
  }
}

/// Bitmap of vertices with an empty list (padding bits of the last word set too): 1 warp per 32-vertex word.
template <typename vertex_t, typename edge_t>
__global__ void __launch_bounds__(256)
    isolated_bitmap_kernel(vertex_t n, const edge_t* __restrict__ offsets, unsigned* __restrict__ words) {
  const unsigned lane = threadIdx.x & 31;
  const std::size_t n_words = (std::size_t(n) + 31) / 32;
  const std::size_t warps = (std::size_t(gridDim.x) * blockDim.x) >> 5;
  for (std::size_t w = (std::size_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5; w < n_words; w += warps) {
    const std::size_t v = w * 32 + lane;
    const bool skip = v >= std::size_t(n) || offsets[v + 1] == offsets[v];
    const unsigned word = __ballot_sync(0xffffffffu, skip);
    if (lane == 0) words[w] = word;
  }
}

template <typename vertex_t, typename edge_t, typename weight_t>
void transpose_on_device(vertex_t n, edge_t m, const edge_t* Ap, const vertex_t* J, const weight_t* X,
                         vertex_t* I, edge_t* Aj, weight_t* csc_values) {
  error::throw_if_exception(cudaMemset(Aj, 0, sizeof(edge_t) * (std::size_t(n) + 1)), "transpose memset");
  if (m > 0) column_histogram_kernel<<<2048, 256>>>(m, J, Aj);
  single_cta_inclusive_scan_kernel<<<1, 1024>>>(Aj, std::size_t(n) + 1);
  memory::device_array_t<edge_t> cursor(std::size_t(n) + 1);
  error::throw_if_exception(
      cudaMemcpy(cursor.data(), Aj, sizeof(edge_t) * (std::size_t(n) + 1), cudaMemcpyDeviceToDevice),
      "transpose cursor");
  transpose_scatter_kernel<<<2048, 256>>>(n, Ap, J, X, cursor.data(), I, csc_values);
  error::throw_if_exception(cudaDeviceSynchronize(), "transpose");
}

}  // namespace detail

template <memory_space_t space, view_t build_views, typename edge_t, typename vertex_t, typename weight_t>
auto from_csr(vertex_t const& r, vertex_t const& c, edge_t const& nnz, edge_t* Ap, vertex_t* J, weight_t* X,
              vertex_t* I = nullptr, edge_t* Aj = nullptr, weight_t* csc_values = nullptr) {
  constexpr bool want_csr = has(build_views, view_t::csr);
  constexpr bool want_csc = has(build_views, view_t::csc);
  constexpr bool want_coo = has(build_views, view_t::coo);
  static_assert(!(want_csc && want_coo), "CSC and COO views share the row-index buffer; build one of them");

  using csr_v_t = std::conditional_t<want_csr, graph_csr_t<vertex_t, edge_t, weight_t>, empty_csr_t>;
  using csc_v_t = std::conditional_t<want_csc, graph_csc_t<vertex_t, edge_t, weight_t>, empty_csc_t>;
  using coo_v_t = std::conditional_t<want_coo, graph_coo_t<vertex_t, edge_t, weight_t>, empty_coo_t>;
  using graph_type = graph_t<space, vertex_t, edge_t, weight_t, csr_v_t, csc_v_t, coo_v_t>;

  graph_type G;
  if constexpr (want_csr) G.template set<csr_v_t>(r, nnz, Ap, J, X);

  if constexpr (want_coo) {
    error::throw_if_exception(I == nullptr, "from_csr: COO view needs a row_indices buffer");
    if constexpr (space == memory_space_t::device) {
      detail::rows_of_edges_kernel<<<2048, 256>>>(r, Ap, I);
      error::throw_if_exception(cudaDeviceSynchronize(), "from_csr coo");
    } else {
      for (vertex_t row = 0; row < r; ++row)
        for (edge_t e = Ap[row]; e < Ap[row + 1]; ++e) I[e] = row;
    }
    G.template set<coo_v_t>(r, nnz, I, J, X);
  }

  if constexpr (want_csc) {
    error::throw_if_exception(I == nullptr || Aj == nullptr,
                              "from_csr: CSC view needs row_indices and column_offsets buffers");
    if (I == J && Aj == Ap) {  // declared symmetric: alias, nothing to compute
      G.get_properties().symmetric = true;
      G.template set<csc_v_t>(r, nnz, Ap, J, X);
    } else {
      static_assert(space == memory_space_t::device || !want_csc, "CSC build runs on the device");
      detail::transpose_on_device(r, nnz, Ap, J, X, I, Aj, csc_values);
      G.template set<csc_v_t>(r, nnz, Aj, I, csc_values);
    }
  }
  (void)c;
  return G;
}

/**
 * @brief Fill caller-owned hint arrays (head[n], head_edge[n]) for the CSC view of G and attach them.
 * Bottom-up advance probes head[v] first: it is read coalesced next to the row bounds, and the highest-degree
 * in-neighbour is the likeliest to be in a BFS frontier, so most vertices never touch their adjacency list
 * (one random 32-byte sector each otherwise). Setup cost, once per graph; results do not depend on it.
 */
template <typename graph_type>
void pull_hints(graph_type& G, typename graph_type::vertex_type* head, typename graph_type::edge_type* head_edge,
                cudaStream_t stream = 0, const typename graph_type::vertex_type* degree_of = nullptr,
                unsigned* isolated_words = nullptr) {
  // isolated_words (optional, ceil(n/32) words): receives the bitmap of vertices without in-edges, the start
  // state of the visited set of every direction-optimised traversal (saves an n-length pass per run).
  // degree_of (optional): degree of every vertex id that can appear as an in-neighbour. Needed when G holds
  // only a row range of a partitioned graph, whose offsets cannot answer degree queries for remote ids.
  using csr_v = typename graph_type::graph_csr_view_t;
  using csc_v = typename graph_type::graph_csc_view_t;
  static_assert(graph_type::template contains_representation<csc_v>(), "pull hints belong to the CSC view");
  csc_v& c = G;
  const auto* degree_offsets = c.get_column_offsets();
  if constexpr (graph_type::template contains_representation<csr_v>()) {
    const csr_v& r = G;
    if (r.get_row_offsets()) degree_offsets = r.get_row_offsets();
  }
  const auto n = c.get_number_of_vertices();
  using vertex_type = typename graph_type::vertex_type;
  memory::device_array_t<vertex_type> degree_scratch;
  if (n > 0 && degree_of == nullptr && sizeof(vertex_type) == 4) {  // whole graph: degrees by one coalesced pass
    degree_scratch.resize(std::size_t(n));
    detail::degrees_kernel<<<2048, 256, 0, stream>>>(n, degree_offsets, degree_scratch.data());
    degree_of = degree_scratch.data();
  }
  if (n > 0)
    detail::pull_hints_kernel<<<2048, 256, 0, stream>>>(decltype(n)(0), n, c.get_column_offsets(), c.get_row_indices(),
                                                        degree_offsets, degree_of, head, head_edge);
  if (n > 0 && isolated_words)
    detail::isolated_bitmap_kernel<<<2048, 256, 0, stream>>>(n, c.get_column_offsets(), isolated_words);
  error::throw_if_exception(cudaStreamSynchronize(stream), "pull_hints");
  c.set_pull_hints(head, head_edge, isolated_words);
}

/**
 * @brief Hints of the vertices [first, last) only, enqueued on `stream` without synchronising: the building block of a
 * graph that is still arriving from the host (ess_graph_create_from_host copies the column indices in chunks and calls
 * this for the vertices whose lists are complete, so the hint build hides behind the PCIe copy). `degree_of` must hold
 * the degree of every id that can appear in the lists.
 */
template <typename vertex_t, typename edge_t>
void pull_hints_range(vertex_t first, vertex_t last, const edge_t* in_offsets, const vertex_t* in_indices,
                      const vertex_t* degree_of, vertex_t* head, edge_t* head_edge, cudaStream_t stream) {
  if (last <= first) return;
  const std::size_t ctas = (std::size_t(last - first) + 255) / 256;
  detail::pull_hints_kernel<<<unsigned(ctas < 2048 ? ctas : 2048), 256, 0, stream>>>(
      first, last, in_offsets, in_indices, in_offsets, degree_of, head, head_edge);
}

/// Wrap pre-built CSR and CSC arrays (no computation): used by the C ABI when the caller owns both.
template <typename vertex_t, typename edge_t, typename weight_t>
auto from_csr_and_csc(vertex_t n, edge_t m, edge_t* Ap, vertex_t* J, weight_t* X, edge_t* Aj, vertex_t* I,
                      weight_t* Xt) {
  using csr_v_t = graph_csr_t<vertex_t, edge_t, weight_t>;
  using csc_v_t = graph_csc_t<vertex_t, edge_t, weight_t>;
  graph_t<memory_space_t::device, vertex_t, edge_t, weight_t, csr_v_t, csc_v_t, empty_coo_t> G;
  G.template set<csr_v_t>(n, m, Ap, J, X);
  G.template set<csc_v_t>(n, m, Aj, I, Xt);
  G.get_properties().symmetric = (Aj == Ap && I == J);
  return G;
}

}  // namespace build
}  // namespace graph
}  // namespace gunrock
