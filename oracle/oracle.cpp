// TEST INFRASTRUCTURE ONLY — the CPU checker for the frontier-operator hot path.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
// this library; the product path (include/, essentials_b200/) never does and has no CPU fallback.
//
// Plain C++17 restatement of what the reference's algorithms compute, one function per algorithm,
// each citing the reference file:line it follows (paths relative to /root/reference).
// PARITY PINNING: tests/test_oracle.py checks every function here against
//   (1) the reference's own CPU code compiled from source (oracle/_ref/libref_cpu.so, ref_shim.cpp) on
//       datasets/chesapeake/chesapeake.mtx (the graph the reference CI runs, .github/workflows/ubuntu.yml:79)
//       and on seeded RMAT / grid graphs, and
//   (2) the committed known-answer vectors tests/golden/*.npz produced by tests/golden/make_golden.py
//       from that same reference code.
// Offsets are int64 here so one oracle serves both edge_t widths; vertex ids are int32.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <queue>
#include <utility>
#include <vector>

namespace {
using i64 = std::int64_t;
using i32 = std::int32_t;

struct stopwatch {
  std::chrono::high_resolution_clock::time_point t0 = std::chrono::high_resolution_clock::now();
  float ms() const {
    return std::chrono::duration<float, std::milli>(std::chrono::high_resolution_clock::now() - t0).count();
  }
};
}  // namespace

extern "C" {

// ---------------------------------------------------------------------------------------------
// BFS depths. Follows examples/algorithms/bfs/bfs_cpu.hxx:21-68: unreachable = INT_MAX, source = 0,
// depth(v) = min hops. The reference runs a priority-queue Dijkstra with unit weights; hop counts are
// a unique fixed point, so a level-synchronous queue gives the identical array. Timed like the
// reference (clock around the search only, bfs_cpu.hxx:35,65-67). Returns ms.
float oracle_bfs(i64 n, const i64* off, const i32* col, i32 src, i32* depth) {
  std::fill(depth, depth + n, std::numeric_limits<i32>::max());
  stopwatch sw;
  std::vector<i32> cur, nxt;
  depth[src] = 0;
  cur.push_back(src);
  i32 level = 0;
  while (!cur.empty()) {
    ++level;
    for (i32 u : cur)
      for (i64 e = off[u]; e < off[u + 1]; ++e) {
        i32 v = col[e];
        if (level < depth[v]) {
          depth[v] = level;
          nxt.push_back(v);
        }
      }
    cur.swap(nxt);
    nxt.clear();
  }
  return sw.ms();
}

// ---------------------------------------------------------------------------------------------
// SSSP. Follows examples/algorithms/sssp/sssp_cpu.hxx:23-72 line by line in spirit: float distances,
// unreachable = FLT_MAX, lazy-deletion binary heap keyed on tentative distance, relax with a single
// float add `curr_dist + w` (same rounding as the GPU lambda, include/gunrock/algorithms/sssp.hxx:116-117).
// For non-negative weights fl(a+w) is monotone in a, so the fixed point is unique and the GPU's
// label-correcting order reaches bit-identical floats.
float oracle_sssp(i64 n, const i64* off, const i32* col, const float* w, i32 src, float* dist) {
  std::fill(dist, dist + n, std::numeric_limits<float>::max());
  stopwatch sw;
  using item = std::pair<float, i32>;
  std::priority_queue<item, std::vector<item>, std::greater<item>> heap;
  dist[src] = 0.0f;
  heap.push({0.0f, src});
  while (!heap.empty()) {
    auto [d, u] = heap.top();
    heap.pop();
    if (d > dist[u]) continue;  // stale entry: the reference re-scans it, which cannot change any label
    for (i64 e = off[u]; e < off[u + 1]; ++e) {
      float cand = d + w[e];
      i32 v = col[e];
      if (cand < dist[v]) {
        dist[v] = cand;
        heap.push({cand, v});
      }
    }
  }
  return sw.ms();
}

// ---------------------------------------------------------------------------------------------
// PageRank. The reference has NO pr_cpu (examples/algorithms/pr/ holds only pr.cu); this restates the GPU
// algorithm include/gunrock/algorithms/pr.hxx:
//   reset   :64-92   p = 1/n, iweights[v] = alpha / sum_w(v) (0 for dangling v)
//   loop    :120-146 plast = p; dsum = sum_{iweights==0} alpha*p; p = (1-alpha+dsum)/n;
//                    p[dst] += plast[src]*iweights[src]*w   for every edge
//   stop    :155-178 after >=1 iteration, stop when max|p-plast| < tol
// Accumulation is in double (the GPU sums in float with an unordered atomicAdd, so only a tolerance
// comparison is meaningful); inputs/outputs are float like the reference. `force_iters` > 0 runs exactly
// that many iterations (used to compare at the GPU's own iteration count); returns iterations run.
int oracle_pr(i64 n, const i64* off, const i32* col, const float* w, float alpha, float tol,
              int force_iters, int max_iters, float* p_out, float* ms_out) {
  std::vector<double> p(n, 1.0 / double(n)), plast(n, 0.0), iw(n, 0.0);
  for (i64 v = 0; v < n; ++v) {
    double s = 0;
    for (i64 e = off[v]; e < off[v + 1]; ++e) s += w ? double(w[e]) : 1.0;
    iw[v] = s != 0 ? double(alpha) / s : 0.0;
  }
  stopwatch sw;
  int it = 0;
  for (;;) {
    if (it > 0 && force_iters <= 0) {
      double err = 0;
      for (i64 v = 0; v < n; ++v) err = std::max(err, std::fabs(p[v] - plast[v]));
      if (err < double(tol)) break;
    }
    if (force_iters > 0 && it >= force_iters) break;
    if (it >= max_iters) break;
    plast = p;
    double dsum = 0;
    for (i64 v = 0; v < n; ++v)
      if (iw[v] == 0) dsum += double(alpha) * plast[v];
    std::fill(p.begin(), p.end(), (1.0 - double(alpha) + dsum) / double(n));
    for (i64 u = 0; u < n; ++u) {
      double base = plast[u] * iw[u];
      if (base == 0) continue;
      for (i64 e = off[u]; e < off[u + 1]; ++e) p[col[e]] += base * (w ? double(w[e]) : 1.0);
    }
    ++it;
  }
  if (ms_out) *ms_out = sw.ms();
  for (i64 v = 0; v < n; ++v) p_out[v] = float(p[v]);
  return it;
}

// ---------------------------------------------------------------------------------------------
// Personalised PageRank, one seed. Follows examples/algorithms/ppr/ppr_cpu.hxx:52-86 (float arithmetic,
// frontier order = discovery order, threshold test on the old/new residual) which itself mirrors
// include/gunrock/algorithms/ppr.hxx:120-146. p must be zeroed by the caller (ppr.cu zero-fills it).
float oracle_ppr(i64 n, const i64* off, const i32* col, i32 seed, float alpha, float eps, float* p) {
  stopwatch sw;
  std::vector<float> r(n, 0.0f), rp(n, 0.0f);
  std::vector<i32> f, fn;
  r[seed] = 1;
  rp[seed] = 1;
  f.push_back(seed);
  while (!f.empty()) {
    for (i32 v : f) {
      p[v] += (2 * alpha) / (1 + alpha) * r[v];
      rp[v] = 0;
    }
    for (i32 u : f) {
      i32 du = i32(off[u + 1] - off[u]);
      float inv = r[u] / du;
      for (i64 e = off[u]; e < off[u + 1]; ++e) {
        i32 v = col[e];
        float upd = ((1 - alpha) / (1 + alpha)) * inv;
        float oldv = rp[v], newv = rp[v] + upd;
        float th = i32(off[v + 1] - off[v]) * eps;
        rp[v] = newv;
        if (oldv < th && newv >= th) fn.push_back(v);
      }
    }
    r = rp;
    f.swap(fn);
    fn.clear();
  }
  return sw.ms();
}

// ---------------------------------------------------------------------------------------------
// k-core numbers. Follows examples/algorithms/kcore/kcore_cpu.hxx:8-62: peel all vertices with remaining
// degree <= k (k = 1,2,...) until none is left at that k; isolated vertices keep 0. This is the standard
// core decomposition, so a bucket-free restatement with the same peel rule gives the same integers as
// the GPU enactor (include/gunrock/algorithms/kcore.hxx:112-199).
float oracle_kcore(i64 n, const i64* off, const i32* col, i32* core) {
  std::vector<i32> deg(n);
  std::vector<i32> alive;
  for (i64 v = 0; v < n; ++v) {
    core[v] = 0;
    deg[v] = i32(off[v + 1] - off[v]);
    if (deg[v]) alive.push_back(i32(v));
  }
  stopwatch sw;
  std::vector<i32> keep, peel;
  for (i32 k = 1; !alive.empty(); ++k) {
    for (;;) {
      keep.clear();
      peel.clear();
      for (i32 v : alive) (deg[v] <= k ? peel : keep).push_back(v);
      alive.swap(keep);
      if (peel.empty()) break;
      for (i32 v : peel) {
        core[v] = k;
        for (i64 e = off[v]; e < off[v + 1]; ++e) --deg[col[e]];
      }
    }
  }
  return sw.ms();
}

// ---------------------------------------------------------------------------------------------
// The reference's seed-free random stream: include/gunrock/algorithms/generate/random.hxx:20-33 —
// thrust::default_random_engine (minstd_rand: x <- 48271*x mod (2^31-1), seed 1), discard(i), then
// thrust::uniform_real_distribution<float>(lo,hi): float(x - min) / (1 + float(max - min)) * (hi-lo) + lo
// with min = 1, max = 2^31-2 (thrust/random/detail/uniform_real_distribution.inl). color.hxx:65 calls
// it with (0, n).
void oracle_randoms(i64 n, float lo, float hi, float* out) {
  const std::uint64_t M = 2147483647ull, A = 48271ull;
  std::uint64_t x = 1;  // state before any draw
  for (i64 i = 0; i < n; ++i) {
    x = (x * A) % M;  // the (i+1)-th output is what discard(i) followed by one draw returns
    float r = static_cast<float>(std::uint32_t(x) - 1u);
    r /= (1.0f + static_cast<float>(2147483646u - 1u));
    out[i] = r * (hi - lo) + lo;
  }
}

// ---------------------------------------------------------------------------------------------
// Graph colouring, Jacobi restatement of include/gunrock/algorithms/color.hxx:99-146: iteration `it`
// hands out colours 2*it (local random maximum) and 2*it+1 (local minimum) among still-uncoloured
// vertices; a neighbour takes part in the comparison unless it already holds a colour from an EARLIER
// iteration. Every vertex of an iteration reads the colour array as it stood at the start of the
// iteration (snapshot) — the reference GPU kernel reads it while other threads write it, so it is not
// run-to-run deterministic; our filter kernel evaluates against the same snapshot semantics by design
// (DESIGN.md "colouring"). Ties on the random value break on vertex id exactly as color.hxx:129-132.
// Returns the number of iterations.
int oracle_color_jacobi(i64 n, const i64* off, const i32* col, const float* rnd, i32* color) {
  std::fill(color, color + n, -1);
  std::vector<i32> active(n), keep, snap(n);
  for (i64 v = 0; v < n; ++v) active[v] = i32(v);
  int it = 0;
  while (!active.empty()) {
    std::memcpy(snap.data(), color, sizeof(i32) * n);
    const i32 c0 = 2 * it;
    keep.clear();
    for (i32 v : active) {
      i64 deg = off[v + 1] - off[v];
      if (deg == 0) {
        color[v] = c0;
        continue;
      }
      bool cmax = true, cmin = true;
      for (i64 e = off[v]; e < off[v + 1]; ++e) {
        i32 u = col[e];
        if ((snap[u] != -1 && snap[u] != c0 && snap[u] != c0 + 1) || u == v) continue;
        if (rnd[v] < rnd[u] || (rnd[v] == rnd[u] && v < u)) cmax = false;
        if (rnd[v] > rnd[u] || (rnd[v] == rnd[u] && v > u)) cmin = false;
      }
      if (cmax)
        color[v] = c0;
      else if (cmin)
        color[v] = c0 + 1;
      else
        keep.push_back(v);
    }
    active.swap(keep);
    ++it;
  }
  return it;
}

// Validity check of a colouring, as examples/algorithms/color/color_cpu.hxx:75-106: counts directed
// edges (v,u), u != v, whose endpoints share a colour or whose source is uncoloured.
i64 oracle_color_errors(i64 n, const i64* off, const i32* col, const i32* color) {
  i64 bad = 0;
  for (i64 v = 0; v < n; ++v)
    for (i64 e = off[v]; e < off[v + 1]; ++e) {
      i32 u = col[e];
      if (u == v) continue;
      if (color[u] == color[v] || color[v] == -1) ++bad;
    }
  return bad;
}

// ---------------------------------------------------------------------------------------------
// Workload accounting used by bench.py and the tests: number of reached vertices n' and of directed
// edges leaving them m' (GTEPS = m'/t), from a depth/distance array.
void oracle_reached_i32(i64 n, const i64* off, const i32* depth, i64* n_reached, i64* m_reached) {
  i64 nr = 0, mr = 0;
  for (i64 v = 0; v < n; ++v)
    if (depth[v] != std::numeric_limits<i32>::max()) {
      ++nr;
      mr += off[v + 1] - off[v];
    }
  *n_reached = nr;
  *m_reached = mr;
}

}  // extern "C"
