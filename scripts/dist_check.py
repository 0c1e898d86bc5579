"""Multi-GPU parity check (run under torchrun, one rank per GPU): partitioned BFS == single-GPU BFS.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
      scripts/dist_check.py --scale 20
"""
import argparse
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import essentials_b200 as ess  # noqa: E402
from essentials_b200 import dist as edist  # noqa: E402
from essentials_b200 import graphgen as gg  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--scale", type=int, default=20)
ap.add_argument("--sources", type=int, default=4)
ap.add_argument("--python-loop", action="store_true", help="use the torch.distributed level loop instead of ess_dist_bfs")
args = ap.parse_args()
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
runner = edist.build_partitioned(args.scale, 16, rank, world, dev, native=not args.python_loop)
srcs = runner.pick_sources(args.sources)
ok = True
if rank == 0:
    full = gg.rmat_csr(args.scale, device=dev)
    ctx, g = ess.Context(local), ess.Graph(full)
    assert srcs == gg.pick_sources(full, args.sources)
for s in [0] + srcs:
    torch.cuda.synchronize()
    t = time.time()
    info = runner.bfs(s)
    torch.cuda.synchronize()
    ms = (time.time() - t) * 1e3
    depth = runner.gather_depth()
    if rank == 0:
        want, _ = ess.bfs(ctx, g, s, lb="merge_path", direction="optimized")
        same = bool(torch.equal(want, depth))
        ok &= same
        print(f"src={s} levels={info['iterations']} pull={info['pull_steps']} wall={ms:.2f} ms "
              f"device={info['enact_ms']:.2f} ms equal={same}", flush=True)
trace = []
runner.bfs(srcs[-1], trace=trace)
if rank == 0 and trace:
    import collections
    agg = collections.OrderedDict()
    for lvl, phase, dt in trace:
        agg.setdefault(phase, []).append(dt * 1e6)
    print("phase breakdown (us, per level, synchronised after each phase): " +
          "; ".join(f"{k}: n={len(v)} avg={sum(v)/len(v):.0f} max={max(v):.0f}" for k, v in agg.items()), flush=True)
flag = torch.tensor([int(ok)], device=dev)
dist.broadcast(flag, 0)
dist.destroy_process_group()
if rank == 0:
    print("DIST_CHECK_OK" if ok else "DIST_CHECK_FAILED", flush=True)
sys.exit(0 if int(flag.item()) else 1)
