"""CPU suite: the synthetic input builders (host logic) — shapes BASELINE.json names, at small scale."""
import numpy as np
import torch

from essentials_b200 import graphgen as gg


def _rows(off):
    return np.repeat(np.arange(off.size - 1), np.diff(off))


def test_rmat_symmetric_sorted_unique_no_loops():
    g = gg.rmat_csr(12)
    off, col, _ = g.host()
    assert g.n == 4096 and off[-1] == g.m == col.size
    rows = _rows(off)
    key = rows.astype(np.int64) * g.n + col
    assert np.all(np.diff(key) > 0), "rows sorted, no duplicate edges"
    assert not np.any(rows == col), "no self loops"
    back = np.sort(col.astype(np.int64) * g.n + rows)
    assert np.array_equal(back, key), "every edge stored in both directions"
    assert g.m <= 2 * 16 * g.n
    deg = np.diff(off)
    assert deg.max() > 20 * deg.mean(), "Kronecker skew"


def test_rmat_is_reproducible_and_chunk_independent():
    a = gg.rmat_csr(11, seed=4)
    b = gg.rmat_csr(11, seed=4, row_chunks=5)
    c = gg.rmat_csr(11, seed=5)
    assert torch.equal(a.offsets, b.offsets) and torch.equal(a.indices, b.indices)
    assert not (a.m == c.m and torch.equal(a.indices, c.indices))


def test_row_range_partition_reassembles():
    full = gg.rmat_csr(11, weights="hash")
    n, parts = full.n, 4
    cols, vals, counts = [], [], []
    for r in range(parts):
        lo, hi = n * r // parts, n * (r + 1) // parts
        p = gg.rmat_csr(11, weights="hash", row_range=(lo, hi))
        assert p.n == hi - lo
        cols.append(p.indices), vals.append(p.values), counts.append(p.degrees())
    assert torch.equal(torch.cat(cols), full.indices)
    assert torch.equal(torch.cat(vals), full.values)
    assert torch.equal(torch.cat(counts), full.degrees())


def test_weights_symmetric_exact_range():
    g = gg.rmat_csr(10, weights="hash")
    off, col, val = g.host()
    rows = _rows(off)
    assert val.min() >= 1.0 and val.max() < 64.0
    w = {(int(r), int(c)): float(v) for r, c, v in zip(rows, col, val)}
    assert all(w[(c, r)] == v for (r, c), v in w.items())
    assert np.array_equal(val, (val * 2**18).round() / 2**18), "weights are dyadic (exact in float32)"


def test_directed_rmat_and_transpose():
    g = gg.rmat_csr(10, symmetric=False, weights="ones")
    t = gg.transpose_csr(g)
    off, col, _ = g.host()
    toff, tcol, _ = t.host()
    fwd = np.sort(_rows(off).astype(np.int64) * g.n + col)
    bwd = np.sort(tcol.astype(np.int64) * g.n + _rows(toff))
    assert np.array_equal(fwd, bwd)


def test_grid():
    g = gg.grid_csr(5, 7)
    off, col, val = g.host()
    assert g.n == 35 and g.m == 2 * (5 * 6 + 4 * 7)
    assert list(col[off[8]:off[9]]) == [1, 7, 9, 15]
    assert np.diff(off).max() == 4 and np.diff(off).min() == 2


def test_pick_sources_non_isolated():
    g = gg.rmat_csr(12)
    s = gg.pick_sources(g, 16)
    assert len(s) == len(set(s)) == 16 and all(g.degrees()[v] > 0 for v in s)
    assert s == gg.pick_sources(g, 16)
