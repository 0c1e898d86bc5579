"""A small pass over the traversal kernels in one process (a few seconds on a GPU): graph built from pageable host arrays
in 1024-edge chunks, BFS through every balancer and both directions, SSSP (frontier, near-far, dense delta), k-core,
filters; every result is checked against the oracle. Written as the workload for `compute-sanitizer --tool memcheck`
(closed on this pool, so the recorded run is plain).
    python scripts/small_pass.py"""
import numpy as np
import torch

import essentials_b200 as ess
import oracle
from essentials_b200 import graphgen as gg

csr = gg.rmat_csr(10, weights="hash", device="cpu")
off, col, val = csr.host()
ctx = ess.Context(0)
ess.tune("host_chunk_edges", 1024)
g = ess.Graph.from_host(ctx, csr)
src = gg.pick_sources(csr, 1)[0]
want = oracle.bfs(off, col, src)
for lb, direction in (("merge_path", "optimized"), ("block_mapped", "forward"), ("bucketing", "optimized"),
                      ("thread_mapped", "forward"), ("merge_path", "forward")):
    d, _ = ess.bfs(ctx, g, src, lb=lb, direction=direction)
    assert np.array_equal(d.cpu().numpy(), want), (lb, direction)
want = oracle.sssp(off, col, val, src)
for run in (lambda: ess.sssp(ctx, g, src, lb="merge_path"), lambda: ess.sssp_near_far(ctx, g, src),
            lambda: ess.sssp_delta(ctx, g, src)):
    dist, _ = run()
    assert np.array_equal(dist.cpu().numpy(), want)
k, _ = ess.kcore(ctx, g)
assert np.array_equal(k.cpu().numpy(), oracle.kcore(off, col))
items = torch.arange(0, csr.n, dtype=torch.int32, device="cuda")
for alg in ("predicated", "compact", "bypass", "remove"):
    ess.filter_probe(ctx, g, items, alg=alg)
print("SMALL_PASS_OK")
