"""Reference GPU implementation vs ours on the same B200 (BASELINE.md §2b "the kernel to beat")."""
import argparse
import torch
import essentials_b200 as ess
import oracle
from essentials_b200 import graphgen as gg

ap = argparse.ArgumentParser()
ap.add_argument("--scale", type=int, default=24)
ap.add_argument("--grid", type=int, default=0)
ap.add_argument("--pr-scale", type=int, default=0)
args = ap.parse_args()
ctx = ess.Context(0)
if args.scale:
    csr = gg.rmat_csr(args.scale, device="cuda")
    g = ess.Graph(csr)
    deg = csr.degrees().long()
    for s in gg.pick_sources(csr, 3):
        for rep in range(2):
            want, ref_ms = oracle.ref_gpu_run("bfs", csr, s)
        m_r = int(deg[want != 2**31 - 1].sum())
        rows = [f"reference block_mapped {ref_ms:9.3f} ms {m_r/ref_ms/1e6:8.2f} GTEPS"]
        for lb, d in (("block_mapped", "forward"), ("merge_path", "forward"), ("merge_path", "optimized")):
            for rep in range(2):
                got, info = ess.bfs(ctx, g, s, lb=lb, direction=d)
            rows.append(f"ours {lb}/{d} {info['enact_ms']:9.3f} ms {m_r/info['enact_ms']/1e6:8.2f} GTEPS equal={torch.equal(got, want)}")
        print(f"BFS scale-{args.scale} src={s}: " + " | ".join(rows), flush=True)
    del g, csr
if args.grid:
    csr = gg.grid_csr(args.grid, args.grid, device="cuda")
    g = ess.Graph(csr)
    want, ref_ms = oracle.ref_gpu_run("sssp", csr, 0)
    got, info = ess.sssp(ctx, g, 0, lb="block_mapped")
    print(f"SSSP grid {args.grid}^2: reference {ref_ms:.1f} ms | ours block_mapped {info['enact_ms']:.1f} ms equal={torch.equal(got, want)}", flush=True)
    del g, csr
if args.pr_scale:
    csr = gg.rmat_csr(args.pr_scale, symmetric=False, weights="ones", device="cuda")
    g = ess.Graph(csr, csc=ess.transpose(csr))
    want, ref_ms = oracle.ref_gpu_run("pr", csr, 0.85, 1e-6)
    got, info = ess.pagerank(ctx, g, lb="merge_path", pull=False)
    rel = ((got.double() - want.double()).abs().sum() / want.double().sum()).item()
    print(f"PR scale-{args.pr_scale}: reference {ref_ms:.1f} ms | ours merge_path {info['enact_ms']:.1f} ms ({info['iterations']} iters) rel-L1 {rel:.2e}", flush=True)
