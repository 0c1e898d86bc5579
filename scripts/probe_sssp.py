"""Quick perf probe (development aid): single-GPU SSSP variants on a weighted Kronecker graph."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import essentials_b200 as ess  # noqa: E402
from essentials_b200 import graphgen as gg  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--scale", type=int, default=24)
ap.add_argument("--sources", type=int, default=3)
ap.add_argument("--repeats", type=int, default=2)
ap.add_argument("--deltas", default="0,2,4,8,16,32")
ap.add_argument("--near-far", action="store_true")
ap.add_argument("--lbs", default="merge_path")
args = ap.parse_args()
csr = gg.rmat_csr(args.scale, device="cuda", weights="hash")
ctx, g = ess.Context(0), ess.Graph(csr)
deg = csr.degrees().long()
for s in gg.pick_sources(csr, args.sources):
    base = None
    for rep in range(args.repeats):
        ess.tune("sssp_fused_unique", 0)
        for lb in args.lbs.split(","):
            d, info = ess.sssp(ctx, g, s, lb=lb)
            if base is None:
                base = d.clone()
                m_r = int(deg[d != 3.4028234663852886e38].sum())
            assert torch.equal(base, d)
            print(f"src={s} rep={rep} {lb:13s} enact={info['enact_ms']:8.2f} ms iters={info['iterations']} "
                  f"GTEPS={m_r / info['enact_ms'] / 1e6:.2f}", flush=True)
        ess.tune("sssp_fused_unique", 1)
        for lb in args.lbs.split(","):
            d, info = ess.sssp(ctx, g, s, lb=lb)
            assert torch.equal(base, d)
            print(f"src={s} rep={rep} {lb + '+unique':20s} enact={info['enact_ms']:8.2f} ms iters={info['iterations']} "
                  f"GTEPS={m_r / info['enact_ms'] / 1e6:.2f}", flush=True)
        for delta in (float(x) for x in args.deltas.split(",")):
            d, info = ess.sssp_delta(ctx, g, s, delta=delta if delta > 0 else 3e38)
            assert torch.equal(base, d)
            print(f"src={s} rep={rep} delta    d={delta:<5g} enact={info['enact_ms']:8.2f} ms rounds={info['rounds']} "
                  f"advances={info['threshold_advances']} expanded={info['expanded_vertices']} "
                  f"GTEPS={m_r / info['enact_ms'] / 1e6:.2f}", flush=True)
        for delta in (float(x) for x in args.deltas.split(",") if args.near_far):
            d, info = ess.sssp_near_far(ctx, g, s, delta=delta)
            assert torch.equal(base, d)
            print(f"src={s} rep={rep} near_far d={delta:<5g} enact={info['enact_ms']:8.2f} ms levels={info['levels']} "
                  f"splits={info['splits']} relax={info['relaxations']} GTEPS={m_r / info['enact_ms'] / 1e6:.2f}",
                  flush=True)
