"""Multi-GPU parity check (run under torchrun, one rank per GPU): partitioned BFS / SSSP == single-GPU result.

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
ap.add_argument("--nccl-exchange", action="store_true", help="ess_dist_bfs: NCCL collectives instead of peer memory")
ap.add_argument("--compare-exchange", action="store_true", help="time the BFS sources with both exchanges")
ap.add_argument("--alg", default="bfs", help="bfs, sssp or bfs,sssp (one graph build for both)")
ap.add_argument("--no-check", action="store_true", help="timing only (scales whose full graph does not fit one GPU)")
ap.add_argument("--enactor", action="store_true",
                help="also run gunrock::bfs::run / sssp::run through the enactor contract with the partitioned context "
                     "(advance::execute + operators::exchange::execute per level) and compare with the single GPU")
args = ap.parse_args()
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
algs = args.alg.split(",")
weights = "hash" if "sssp" in algs else "none"
runner = edist.build_partitioned(args.scale, 16, rank, world, dev, native=not args.python_loop, weights=weights)
if args.nccl_exchange:
    ess.tune("dist_peer_exchange", 0)
if rank == 0 and hasattr(runner, "exchange_kind"):
    print(f"exchange: {runner.exchange_kind()}", flush=True)
srcs = runner.pick_sources(args.sources)
ok = True
if rank == 0 and not args.no_check:
    full = gg.rmat_csr(args.scale, device=dev, weights=weights)
    ctx, g = ess.Context(local), ess.Graph(full)
    assert srcs == gg.pick_sources(full, args.sources)


def check_sssp():
    """Partitioned SSSP (native loop, or the torch.distributed loop of PartitionedSSSP over the same backend)."""
    global ok
    solver = runner
    if args.python_loop:
        solver = edist.PartitionedSSSP(runner.csr, runner.row_begin, runner.n_global, rank, world, runner.backend, dev)
    per_source = []
    for s in [0] + srcs:
        torch.cuda.synchronize()
        dist.barrier()
        t = time.time()
        info = solver.sssp(s)
        torch.cuda.synchronize()
        ms = (time.time() - t) * 1e3
        n_r, m_r = solver.reached_work() if args.python_loop else solver.reached_work_sssp()
        line = (f"src={s} rounds={info['iterations']} wall={ms:.2f} ms device={info['enact_ms']:.2f} ms "
                f"reached={n_r} m'={m_r} relaxed={info['relaxed_edges']} "
                f"GTEPS={m_r / max(ms, 1e-9) / 1e6:.2f} recvMB={info['nvlink_bytes_received'] / 1e6:.1f} "
                f"exchange={info.get('exchange', 'torch.distributed')}")
        if not args.no_check:
            got = solver.gather_dist()
            if rank == 0:
                want, winfo = ess.sssp(ctx, g, s, lb="merge_path")
                same = bool(torch.equal(want, got))
                ok &= same
                line += f" equal={same} single_gpu_ms={winfo['enact_ms']:.2f}"
        if rank == 0:
            print(line, flush=True)
        per_source.append(ms)


def check_bfs():
    global ok
    for s in [0] + srcs:
        torch.cuda.synchronize()
        dist.barrier()
        t = time.time()
        info = runner.bfs(s)
        torch.cuda.synchronize()
        ms = (time.time() - t) * 1e3
        n_r, m_r = runner.reached_work()
        line = (f"src={s} levels={info['iterations']} pull={info['pull_steps']} wall={ms:.2f} ms "
                f"device={info['enact_ms']:.2f} ms m'={m_r} GTEPS={m_r / max(ms, 1e-9) / 1e6:.2f}")
        if not args.no_check:
            depth = runner.gather_depth()
            if rank == 0:
                want, _ = ess.bfs(ctx, g, s, lb="merge_path", direction="optimized")
                same = bool(torch.equal(want, depth))
                ok &= same
                line += f" equal={same}"
        if rank == 0:
            print(line, flush=True)
    trace = []
    runner.bfs(srcs[-1], trace=trace)
    if rank == 0 and trace:
        import collections
        agg = collections.OrderedDict()
        for lvl, phase, dt in trace:
            agg.setdefault(phase, []).append(dt * 1e6)
        print("phase breakdown (us, per level, synchronised after each phase): " +
              "; ".join(f"{k}: n={len(v)} avg={sum(v)/len(v):.0f} max={max(v):.0f}" for k, v in agg.items()), flush=True)


def check_enactor():
    """The operator-API form: same enactor loop as on one GPU, frontier exchange by operators::exchange."""
    global ok
    for alg in algs:
        for lb in ("merge_path", "block_mapped", "thread_mapped"):
            for s in [0] + srcs[:2]:
                dist.barrier()
                if alg == "bfs":
                    mine, info = runner.bfs_enactor(s, lb=lb)
                    full = torch.empty(runner.n_global, dtype=torch.int32, device=dev)
                else:
                    mine, info = runner.sssp_enactor(s, lb=lb)
                    full = torch.empty(runner.n_global, dtype=torch.float32, device=dev)
                dist.all_gather_into_tensor(full, mine.contiguous())
                line = f"enactor {alg} lb={lb} src={s} iterations={info['iterations']} device={info['enact_ms']:.2f} ms"
                if rank == 0 and not args.no_check:
                    if alg == "bfs":
                        want, _ = ess.bfs(ctx, g, s, lb=lb, direction="forward")
                    else:
                        want, _ = ess.sssp(ctx, g, s, lb=lb)
                    same = bool(torch.equal(want, full))
                    ok &= same
                    line += f" equal={same}"
                if rank == 0:
                    print(line, flush=True)


if args.enactor:
    check_enactor()
    algs = []  # the enactor form is checked on its own
if "bfs" in algs:
    check_bfs()
    if args.compare_exchange:
        ess.tune("dist_trace", 1)
        for mode in (1, 0):
            ess.tune("dist_peer_exchange", mode)
            for s in srcs[:3]:
                dist.barrier()
                torch.cuda.synchronize()
                runner.bfs(s)
        ess.tune("dist_trace", 0)
        for mode in (1, 0, 1, 0):
            ess.tune("dist_peer_exchange", mode)
            times = []
            for s in srcs:
                dist.barrier()
                torch.cuda.synchronize()
                times.append(runner.bfs(s)["enact_ms"])
            if rank == 0:
                print(f"exchange={runner.exchange_kind():11s} device ms per source: " +
                      " ".join(f"{t:.3f}" for t in times) + f"  mean={sum(times) / len(times):.3f}", flush=True)
if "sssp" in algs:
    check_sssp()
    if args.compare_exchange and not args.python_loop:
        for mode in (1, 0, 1, 0):
            ess.tune("dist_peer_exchange", mode)
            times, kinds = [], set()
            for s in srcs:
                dist.barrier()
                torch.cuda.synchronize()
                info = runner.sssp(s)
                times.append(info["enact_ms"])
                kinds.add(info["exchange"])
            if rank == 0:
                print(f"sssp exchange={'/'.join(sorted(kinds)):11s} device ms per source: " +
                      " ".join(f"{t:.3f}" for t in times) + f"  mean={sum(times) / len(times):.3f}", flush=True)
        ess.tune("dist_peer_exchange", 1)
flag = torch.tensor([int(ok)], device=dev)
dist.broadcast(flag, 0)
dist.destroy_process_group()
if rank == 0:
    print("DIST_CHECK_OK" if ok else "DIST_CHECK_FAILED", flush=True)
sys.exit(0 if int(flag.item()) else 1)
