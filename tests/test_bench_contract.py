"""CPU suite: bench.py's contract where it can be exercised without a GPU — the reference arm (`--impl reference`,
the reference's own bfs_cpu from oracle/_ref, or the oracle port) prints ONE JSON line with the agreed keys, and
the product arm refuses to run without a CUDA device instead of falling back to anything on the CPU."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--scale", "12",
                          "--steps", "2", "--warmup", "1"], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-1500:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "BFS GTEPS" and d["unit"] == "GTEPS"
    for key in ("value", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"):
        assert key in d, key
    assert d["steps"] == 2 and d["warmup"] == 1 and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert "workload" in d["config"] and d["value"] > 0 and d["gpu_launches"] == 0
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and 1 <= cb["cores"] <= 2 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "GTEPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # same `config` object as the product arm prints for this workload (the driver compares them)
    sys.path.insert(0, ROOT)
    import bench
    assert d["config"] == bench.workload_config(12, 16, d["config"]["n"], d["config"]["m"], 2)
    assert d["scaling"] == "strong"


def test_default_workload_is_the_north_star():
    """No flags = Kronecker scale-26 (BASELINE.json north_star) at every N: strong scaling."""
    sys.path.insert(0, ROOT)
    import bench
    old = sys.argv
    try:
        sys.argv = ["bench.py"]
        args = bench.parse()
    finally:
        sys.argv = old
    assert args.scale == 0 and bench.NORTH_STAR_SCALE == 26 and not args.weak


def test_certificates_accept_the_oracle_and_reject_corruption():
    """bench.py's full-size parity checks: the BFS / SSSP certificates pass on the oracle's answers (whole graph and
    a 2-way row partition) and flag any single corrupted entry."""
    import numpy as np
    import torch
    sys.path.insert(0, ROOT)
    import bench
    import oracle
    from essentials_b200 import graphgen as gg
    csr = gg.rmat_csr(11, 8, weights="hash", device="cpu")
    off, col, val = csr.host()
    src = gg.pick_sources(csr, 1)[0]
    depth = torch.from_numpy(oracle.bfs(off, col, src))
    dist = torch.from_numpy(oracle.sssp(off, col, val, src))
    assert bench.bfs_certificate(csr.offsets, csr.indices, depth, src) == 0
    assert bench.sssp_certificate(csr.offsets, csr.indices, csr.values, dist, src) == 0
    half = csr.n // 2
    for lo, hi in ((0, half), (half, csr.n)):  # a rank's rows against the full result array
        part = gg.rmat_csr(11, 8, weights="hash", device="cpu", row_range=(lo, hi))
        assert bench.bfs_certificate(part.offsets, part.indices, depth, src, row_begin=lo) == 0
        assert bench.sssp_certificate(part.offsets, part.indices, part.values, dist, src, row_begin=lo) == 0
    reached = (depth != bench.INF).nonzero().flatten()
    v = int(reached[reached != src][7])
    for delta in (1, -1):
        bad = depth.clone()
        bad[v] += delta
        assert bench.bfs_certificate(csr.offsets, csr.indices, bad, src) > 0
    bad = depth.clone()
    bad[v] = bench.INF  # a reached vertex reported unreached
    assert bench.bfs_certificate(csr.offsets, csr.indices, bad, src) > 0
    for factor in (0.5, 1.5):
        badd = dist.clone()
        badd[v] *= factor
        assert bench.sssp_certificate(csr.offsets, csr.indices, csr.values, badd, src) > 0


def test_product_arm_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        return  # on a GPU box the product arm is exercised by the driver itself
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "0",
                          "--scale", "10"], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode != 0 and not out.stdout.strip(), "no CPU fallback: nothing may be printed as a result"
