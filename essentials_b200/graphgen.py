"""Synthetic graph inputs of the shapes BASELINE.json names (host-side setup; not on the timed path).

Everything here is integer/hash arithmetic written with torch ops so the *same* graph comes out on CPU
(tests, golden fixtures) and on the GPU (bench sizes): a counter-based splitmix64 stream replaces the
device RNG, edge weights are exact dyadic floats, and there is no floating-point comparison anywhere.

* ``rmat_csr``  — Graph500-style Kronecker/R-MAT (A,B,C,D)=(.57,.19,.19,.05), vertex ids permuted,
  then symmetrised, self-loops dropped, duplicates removed, adjacency sorted (BASELINE.md config 1/2/★),
  or kept directed (config 4, PageRank).
* ``grid_csr``  — 2-D 4-neighbour grid, the "road-like" high-diameter input (config 3).
* weights       — uniform in [1,64) from a symmetric hash of the undirected pair.

The reference reads graphs from MatrixMarket files (include/gunrock/io/matrix_market.hxx:99-240) and
converts COO->CSR on the host (include/gunrock/formats/csr.hxx:79-157); file IO is out of scope here
 and the CSR arrays these builders return are exactly what
``graph::build::from_csr`` (include/gunrock/graph/build.hxx:21-36) takes.
"""
from __future__ import annotations

from dataclasses import dataclass

import torch

_MASK64 = (1 << 64) - 1


def _s64(x: int) -> int:
    """Python int -> the int64 two's-complement value with the same low 64 bits."""
    x &= _MASK64
    return x - (1 << 64) if x >= (1 << 63) else x


_GOLDEN = _s64(0x9E3779B97F4A7C15)
_M1 = _s64(0xBF58476D1CE4E5B9)
_M2 = _s64(0x94D049BB133111EB)


def _lsr(z: torch.Tensor, k: int) -> torch.Tensor:
    return (z >> k) & ((1 << (64 - k)) - 1)


def mix64(x: torch.Tensor) -> torch.Tensor:
    """splitmix64 finaliser on int64 tensors (wrapping arithmetic; identical on CPU and CUDA)."""
    z = x + _GOLDEN
    z = (z ^ _lsr(z, 30)) * _M1
    z = (z ^ _lsr(z, 27)) * _M2
    return z ^ _lsr(z, 31)


def permute_ids(v: torch.Tensor, scale: int, seed: int) -> torch.Tensor:
    """A seeded bijection of [0, 2**scale): odd-multiply / xor-shift rounds modulo 2**scale."""
    mask = (1 << scale) - 1
    h = max(scale // 2, 1)
    k1 = (_s64(0x9E3779B97F4A7C15 * (2 * seed + 1)) | 1)
    k2 = (_s64(0xD6E8FEB86659FD93 * (2 * seed + 3)) | 1)
    v = (v * k1 + (seed * 0x51ED27 + 0x2545F491)) & mask
    v = v ^ (v >> h)
    v = (v * k2 + 0x6A09E667) & mask
    v = v ^ (v >> h)
    return v


def rmat_edges(scale: int, edge_factor: int = 16, seed: int = 1, a: float = 0.57, b: float = 0.19,
               c: float = 0.19, device="cpu", first: int = 0, count: int | None = None):
    """Directed R-MAT edge list (src, dst) as int64 tensors, edges ``first .. first+count``.

    One 32-bit draw per (edge, bit level) from mix64(seed, edge*scale+level); quadrant thresholds are
    integers, so the list is bit-reproducible across devices.  Vertex ids are permuted.
    """
    n_edges = edge_factor << scale
    count = n_edges - first if count is None else count
    idx = torch.arange(first, first + count, dtype=torch.int64, device=device)
    ta = int(a * 4294967296.0)
    tab = int((a + b) * 4294967296.0)
    tabc = int((a + b + c) * 4294967296.0)
    src = torch.zeros_like(idx)
    dst = torch.zeros_like(idx)
    base = idx * scale + _s64(seed * 0x632BE59BD9B4E019)
    for level in range(scale):
        r = _lsr(mix64(base + level), 32)
        src_bit = (r >= tab).to(torch.int64)  # quadrants C, D
        dst_bit = ((r >= ta) & (r < tab) | (r >= tabc)).to(torch.int64)  # quadrants B, D
        src = (src << 1) | src_bit
        dst = (dst << 1) | dst_bit
    return permute_ids(src, scale, seed), permute_ids(dst, scale, seed)


def pair_weights(u: torch.Tensor, v: torch.Tensor, seed: int = 7) -> torch.Tensor:
    """Symmetric edge weight in [1, 64): q / 2**18 with q a 24-bit integer, so every value is an exact
    float32 and both directions of an undirected edge agree."""
    lo = torch.minimum(u, v).to(torch.int64)
    hi = torch.maximum(u, v).to(torch.int64)
    h = mix64((lo << 32) ^ hi ^ _s64(seed * 0xA0761D6478BD642F))
    q = (1 << 18) + _lsr(h, 24) % ((1 << 24) - (1 << 18))
    return q.to(torch.float32) / float(1 << 18)


@dataclass
class CSR:
    """Caller-owned CSR arrays (what graph::build::from_csr takes): offsets[n+1], indices[m], values[m]."""
    n: int
    m: int
    offsets: torch.Tensor  # int32 or int64
    indices: torch.Tensor  # int32
    values: torch.Tensor | None  # float32 or None (all-ones semantics)
    name: str = ""
    symmetric: bool = False

    def to(self, device) -> "CSR":
        return CSR(self.n, self.m, self.offsets.to(device), self.indices.to(device),
                   None if self.values is None else self.values.to(device), self.name, self.symmetric)

    def pinned(self) -> "CSR":
        """Host copy in page-locked memory (what Graph.from_host overlaps its transfer from)."""
        pin = lambda t: None if t is None else t.cpu().contiguous().pin_memory()
        return CSR(self.n, self.m, pin(self.offsets), pin(self.indices), pin(self.values), self.name, self.symmetric)

    def host(self):
        """numpy views for the oracle (offsets widened to int64)."""
        off = self.offsets.cpu().numpy().astype("int64")
        col = self.indices.cpu().numpy()
        val = None if self.values is None else self.values.cpu().numpy()
        return off, col, val

    def degrees(self) -> torch.Tensor:
        return self.offsets[1:] - self.offsets[:-1]

    def nbytes(self) -> int:
        b = self.offsets.numel() * self.offsets.element_size() + self.indices.numel() * 4
        return b + (0 if self.values is None else self.values.numel() * 4)


def _offsets_from_sorted_rows(rows: torch.Tensor, n: int, wide: bool) -> torch.Tensor:
    counts = torch.bincount(rows, minlength=n)
    off = torch.zeros(n + 1, dtype=torch.int64, device=rows.device)
    torch.cumsum(counts, 0, out=off[1:])
    return off if wide else off.to(torch.int32)


def rmat_csr(scale: int, edge_factor: int = 16, seed: int = 1, symmetric: bool = True, weights: str = "none",
             device="cpu", offset_bits: int | None = None, row_chunks: int | None = None,
             row_range: tuple[int, int] | None = None) -> CSR:
    """Kronecker/R-MAT graph as CSR.

    symmetric=True : both directions stored, self-loops dropped, duplicates removed, rows sorted.
    symmetric=False: directed as generated, self-loops dropped, duplicates removed, rows sorted.
    weights        : "none" | "ones" | "hash" (pair_weights).
    row_range      : build only the CSR rows [lo, hi) (1-D vertex partition for multi-GPU; column ids stay
                     global) — every rank regenerates the counter-based edge list and keeps its slice.
    row_chunks     : sort/unique is done per block of source rows to bound temporary memory.
    """
    n = 1 << scale
    n_edges = edge_factor << scale
    lo, hi = (0, n) if row_range is None else row_range
    if row_chunks is None:
        row_chunks = max(1, (n_edges * (2 if symmetric else 1)) >> 27)
    gen_chunk = 1 << 26
    bounds = [lo + (hi - lo) * i // row_chunks for i in range(row_chunks + 1)]
    key_parts = [[] for _ in range(row_chunks)]
    for first in range(0, n_edges, gen_chunk):
        s, d = rmat_edges(scale, edge_factor, seed, device=device, first=first,
                          count=min(gen_chunk, n_edges - first))
        keep = s != d
        s, d = s[keep], d[keep]
        if symmetric:
            s, d = torch.cat([s, d]), torch.cat([d, s])
        if row_range is not None:
            keep = (s >= lo) & (s < hi)
            s, d = s[keep], d[keep]
        key = (s << scale) | d
        del s, d, keep
        if row_chunks == 1:
            key_parts[0].append(torch.unique(key))
        else:
            for ci in range(row_chunks):
                sel = key[(key >= (bounds[ci] << scale)) & (key < (bounds[ci + 1] << scale))]
                key_parts[ci].append(torch.unique(sel))
        del key
    cols, rows_counts = [], []
    for ci in range(row_chunks):
        k = torch.unique(torch.cat(key_parts[ci])) if key_parts[ci] else torch.empty(0, dtype=torch.int64,
                                                                                     device=device)
        key_parts[ci] = None
        rows_counts.append(torch.bincount((k >> scale) - lo, minlength=hi - lo))
        cols.append((k & (n - 1)).to(torch.int32))
        del k
    indices = torch.cat(cols) if len(cols) > 1 else cols[0]
    counts = rows_counts[0]
    for rc in rows_counts[1:]:
        counts = counts + rc
    m = int(indices.numel())
    wide = (offset_bits == 64) if offset_bits is not None else (m >= (1 << 31) - 1)
    off = torch.zeros(hi - lo + 1, dtype=torch.int64, device=device)
    torch.cumsum(counts, 0, out=off[1:])
    offsets = off if wide else off.to(torch.int32)
    values = None
    if weights == "ones":
        values = torch.ones(m, dtype=torch.float32, device=device)
    elif weights == "hash":
        rows = torch.repeat_interleave(torch.arange(lo, hi, device=device, dtype=torch.int64), counts)
        values = pair_weights(rows, indices.to(torch.int64))
        del rows
    kind = "kron" if symmetric else "rmat-directed"
    return CSR(hi - lo if row_range is not None else n, m, offsets, indices, values,
               f"{kind}-s{scale}-ef{edge_factor}-seed{seed}", symmetric)


def grid_csr(rows: int, cols: int, weights: str = "hash", device="cpu", offset_bits: int = 32) -> CSR:
    """2-D 4-neighbour grid (road-like, diameter rows+cols-2); adjacency sorted (up, left, right, down)."""
    n = rows * cols
    v = torch.arange(n, dtype=torch.int64, device=device)
    r, c = v // cols, v % cols
    cand = torch.stack([v - cols, v - 1, v + 1, v + cols], 1)
    ok = torch.stack([r > 0, c > 0, c < cols - 1, r < rows - 1], 1)
    counts = ok.sum(1)
    indices = cand[ok].to(torch.int32)
    off = torch.zeros(n + 1, dtype=torch.int64, device=device)
    torch.cumsum(counts, 0, out=off[1:])
    m = int(indices.numel())
    values = None
    if weights == "ones":
        values = torch.ones(m, dtype=torch.float32, device=device)
    elif weights == "hash":
        src = v.unsqueeze(1).expand(-1, 4)[ok]
        values = pair_weights(src, indices.to(torch.int64))
    return CSR(n, m, off if offset_bits == 64 else off.to(torch.int32), indices, values,
               f"grid-{rows}x{cols}", True)


def transpose_csr(g: CSR) -> CSR:
    """CSC of g expressed as the CSR of the transpose (setup cost, torch sort)."""
    dev = g.indices.device
    rows = torch.repeat_interleave(torch.arange(g.n, device=dev, dtype=torch.int64), g.degrees().to(torch.int64))
    key = (g.indices.to(torch.int64) << 32) | rows
    key, perm = torch.sort(key)
    off = _offsets_from_sorted_rows(key >> 32, g.n, g.offsets.dtype == torch.int64)
    vals = None if g.values is None else g.values[perm]
    return CSR(g.n, g.m, off, (key & 0xFFFFFFFF).to(torch.int32), vals, g.name + "-T", g.symmetric)


def pick_sources(g: CSR, count: int, seed: int = 2) -> list[int]:
    """`count` pseudo-random non-isolated vertices (Graph500 convention), reproducible across devices."""
    deg = g.degrees().cpu()
    out, i = [], 0
    n = g.n
    while len(out) < count and i < 64 * count + 1024:
        v = int(mix64(torch.tensor([seed * 1000003 + i], dtype=torch.int64))[0].item() % n)
        if deg[v] > 0 and v not in out:
            out.append(v)
        i += 1
    return out
