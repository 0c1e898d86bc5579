#!/usr/bin/env python
"""bench.py — BASELINE.json's metric on its config: BFS GTEPS on a Kronecker scale-24 edge-factor-16 graph.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

A step is one BFS (one pass of the frontier-operator hot path) from one pseudo-random non-isolated source
of the synthetic graph; the K timed steps use K different sources. GTEPS = directed edges leaving reached
vertices (m') summed over the steps / time. Rank 0 prints ONE JSON line.

  value        device-resident graph, K x ess_bfs, CUDA events on the library's stream (max over ranks).
  e2e          same K sources through the public host API with HOST buffers: every step copies the CSR
               arrays and the source from pinned host memory, runs BFS and copies the depth array back.
  roofline     dominant kernel class of an instrumented repeat of the K steps (CUDA events around every
               launch, essentials_b200.Context.profile): algorithmic bytes / kernel time vs the measured HBM
               peak in MEASURED_PEAKS.json. Byte model in DESIGN.md §Measurement.
  cpu_baseline the reference's own bfs_cpu (oracle/_ref, 1 thread) on a bounded sample, rank 0, N=1 only.
  --impl reference   times that CPU implementation alone (the only other place oracle/ is executed).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

INF = 2**31 - 1


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=16)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--scale", type=int, default=0,
                    help="Kronecker scale; default 24 + log2(gpus): BASELINE config 2 on one GPU and the same "
                         "number of edges PER GPU on more (weak scaling of the 1-D partitioned BFS); an explicit "
                         "--scale fixes the total work (strong scaling)")
    ap.add_argument("--edge-factor", type=int, default=16)
    ap.add_argument("--lb", default="merge_path")
    ap.add_argument("--direction", default="optimized")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--nccl-exchange", action="store_true",
                    help="multi-GPU: NCCL send/recv + all_gather instead of the peer-memory exchange kernels")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the SSSP-grid / PageRank / reference-GPU side runs")
    ap.add_argument("--cpu-budget-s", type=float, default=150.0, help="wall budget of the reference arm")
    return ap.parse_args()


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(kernel_class):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed
    `ncu --set full` capture (profiles/ncu_traffic.json says which capture); None when there is none."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            return json.load(f).get(kernel_class, {}).get("bytes_per_launch")
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i] == "Active"})
        mx = max((int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()), default=None)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": reasons,
                "samples": len(sm)}


def reached_work(csr, depth):
    """(n', m') of a depth array: reached vertices and the directed edges leaving them."""
    import torch
    r = depth != INF
    return int(r.sum()), int(csr.degrees().to(torch.int64)[r].sum())


def cpu_reference_bfs(off, col, source):
    """One run of the reference's bfs_cpu (or our port when oracle/_ref is absent). Returns (depth, ms, kind)."""
    import oracle
    if oracle.have_ref() and off.dtype.itemsize == 4:
        d, ms = oracle.ref_bfs(off, col, source, return_ms=True)
        return d, ms, "reference"
    d, ms = oracle.bfs(off, col, source, return_ms=True)
    return d, ms, "port"


# --------------------------------------------------------------------------------------------- reference arm
def run_reference(args, rank, world):
    if rank != 0:
        return
    import numpy as np
    import torch

    from essentials_b200 import graphgen as gg
    steps_total = args.steps + args.warmup
    # bounded sample: the largest Kronecker scale <= the configured one whose (K+W) CPU traversals fit the budget
    # (bfs_cpu sustains roughly 0.08-0.12 GTEPS on one core incl. its array copies; measured figure is what gets printed)
    scale = args.scale
    while scale > 16 and (args.edge_factor << scale) * 2 * steps_total / 80e6 > args.cpu_budget_s:
        scale -= 1
    dev = "cuda" if torch.cuda.is_available() else "cpu"
    csr = gg.rmat_csr(scale, args.edge_factor, device=dev)
    srcs = gg.pick_sources(csr, steps_total)
    off = csr.offsets.cpu().numpy()
    col = csr.indices.cpu().numpy()
    deg = np.diff(off.astype(np.int64))
    for s in srcs[: args.warmup]:
        cpu_reference_bfs(off, col, s)
    edges, ms_total, kind = 0, 0.0, "port"
    t0 = time.time()
    for s in srcs[args.warmup:]:
        d, ms, kind = cpu_reference_bfs(off, col, s)
        edges += int(deg[d != INF].sum())
        ms_total += ms
    wall = time.time() - t0
    val = edges / (ms_total * 1e-3) / 1e9
    sample = (f"{args.steps} full BFS runs of the reference's single-threaded bfs_cpu on kron scale-{scale} "
              f"ef-{args.edge_factor} (configured scale {args.scale}; search time only, as bfs_cpu.hxx:35,65-67; "
              f"wall {wall:.1f}s)")
    line = {
        "impl": "reference", "metric": "BFS GTEPS", "value": val, "unit": "GTEPS", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
        "scaling": args.scaling, "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "config": {"workload": f"BFS kron scale-{args.scale} ef-{args.edge_factor} (sample: scale-{scale})",
                   "n": csr.n, "m": csr.m},
        "cpu_baseline": {"value": val, "unit": "GTEPS", "cores": 1, "kind": kind, "sample": sample},
        "e2e": {"value": val, "unit": "GTEPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------- our arm
def run_b200(args, rank, world, local_rank):
    import numpy as np
    import torch
    import torch.distributed as dist

    import essentials_b200 as ess
    from essentials_b200 import graphgen as gg

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    distributed = world > 1
    stream = torch.cuda.Stream(device=dev)
    peak_gbs, peak_src = peaks()
    K, W = args.steps, max(args.warmup, 0)

    if distributed:
        from essentials_b200 import dist as edist
        runner = edist.build_partitioned(args.scale, args.edge_factor, rank, world, dev, stream)
        if args.nccl_exchange:
            ess.tune("dist_peer_exchange", 0)
        stream.synchronize()
        n, m, offset_bits = runner.n_global, runner.m_global, runner.offset_bits
        with torch.cuda.stream(stream):
            srcs = runner.pick_sources(K + W)
    else:
        with torch.cuda.stream(stream):
            csr = gg.rmat_csr(args.scale, args.edge_factor, device=dev)
            ctx = ess.Context(local_rank, stream=stream)
            graph = ess.Graph(csr)
            depth = torch.empty(csr.n, dtype=torch.int32, device=dev)
        stream.synchronize()
        n, m, offset_bits = csr.n, csr.m, graph.offset_bits
        srcs = gg.pick_sources(csr, K + W)

    def one_bfs(s):
        if distributed:
            return runner.bfs(s)
        return ess.bfs(ctx, graph, s, lb=args.lb, direction=args.direction, out=depth)[1]

    # ---- untimed pre-pass: per-source work (n', m') and a property check of the result ----------------
    work = {}
    with torch.cuda.stream(stream):
        for s in srcs:
            info = one_bfs(s)
            if distributed:
                work[s] = runner.reached_work()
            else:
                work[s] = reached_work(csr, depth)
        for s in srcs[:W]:  # warm-up proper (allocator pools, clocks)
            one_bfs(s)
    stream.synchronize()
    timed = srcs[W:]
    edges = sum(work[s][1] for s in timed)
    verts = sum(work[s][0] for s in timed)

    # ---- timed region: K steps, device events on the launching stream, barrier + sync on both sides ----
    launches0 = runner.backend.launches() if distributed else ctx.launches()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if distributed:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    enact_ms = 0.0
    with torch.cuda.stream(stream):
        e0.record(stream)
        for s in timed:
            enact_ms += one_bfs(s)["enact_ms"]
        e1.record(stream)
    torch.cuda.synchronize()
    if distributed:
        dist.barrier()
    ms = e0.elapsed_time(e1)
    if distributed:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    clocks = sampler.stop() if sampler else None
    launches = (runner.backend.launches() if distributed else ctx.launches()) - launches0
    value = edges / (ms * 1e-3) / 1e9

    out = {
        "metric": "BFS GTEPS", "value": value, "unit": "GTEPS", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms / K, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
        "dtype": "int32", "data": "synthetic",
        "config": {"workload": f"BFS kron scale-{args.scale} ef-{args.edge_factor} symmetrised, {K} random sources",
                   "n": n, "m": m, "edge_t_bits": offset_bits, "advance": f"{args.lb}/{args.direction}",
                   "parallelism": f"1d-vertex-partition x{world}" if distributed else "single-gpu",
                   "l2": "CSR (%.2f GB) exceeds the 126 MB L2; no flush between steps" % ((m * 4 + n * offset_bits / 8) / 1e9),
                   "teps_edges": "directed edges leaving reached vertices (m')"},
        "enact_ms_per_step": enact_ms / K,
        "gpu_launches": launches,
        "clocks": clocks,
    }

    if not distributed:
        # ---- roofline: instrumented repeat of the same K steps --------------------------------------
        sE = offset_bits // 8
        acc = {"pull_vertices": 0, "pull_edges": 0, "push_vertices": 0, "push_edges": 0, "pull_steps": 0}
        with torch.cuda.stream(stream):
            ctx.profile(True)
            for s in timed:
                info = one_bfs(s)
                for k in acc:
                    acc[k] += info[k]
            prof = ctx.profile_read()
            ctx.profile(False)
        bytes_by_class = {
            # bottom-up levels: 3 bitmaps streamed per level + row bounds and head hint of every walked vertex +
            # in-edges read from the adjacency lists
            "pull_step": acc["pull_steps"] * 3 * (n / 8) + acc["pull_vertices"] * (2 * sE + 4) + acc["pull_edges"] * 4,
            # top-down levels: frontier id + row bounds per expanded vertex + column ids of its out-edges
            "push_expand": acc["push_vertices"] * (4 + 2 * sE) + acc["push_edges"] * 4,
        }
        kernel_ms = {k: v[0] for k, v in prof.items()}
        dom = max(bytes_by_class, key=lambda k: kernel_ms.get(k, 0.0))
        d_ms, d_launches = prof[dom]
        achieved = bytes_by_class[dom] / (d_ms * 1e-3) / 1e9 if d_ms > 0 else 0.0
        graph500_bytes = 8 * edges + (2 * sE + 12) * verts
        out["roofline"] = {
            "bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak_gbs, "unit": "GB/s",
            "frac": achieved / peak_gbs, "traffic": ncu_traffic(dom), "peak_source": peak_src,
            "frac_of_nominal_8000_gbs": achieved / 8000.0,
            "launches": d_launches, "avg_launch_ms": d_ms / max(d_launches, 1),
            "algorithmic_bytes_per_launch": bytes_by_class[dom] / max(d_launches, 1),
            "kernel_ms_per_step": {k: v / K for k, v in kernel_ms.items() if v},
            "share_of_step": d_ms / ms if ms else None,
            "whole_bfs_effective_gbs": graph500_bytes / (ms * 1e-3) / 1e9,
            "note": "achieved counts bytes the kernel actually has to move (early-exited edges are not counted); "
                    "traffic (ncu DRAM bytes per launch, mean over every launch of the class) exceeds the algorithmic "
                    "bytes because each walked vertex costs whole 32-byte sectors for 4-12 useful bytes, not because "
                    "anything is re-read; "
                    "whole_bfs_effective_gbs uses SURVEY §8d's 8*m'+(2*sE+12)*n' over the step time and can exceed "
                    "the peak because direction optimisation skips most edges",
        }

        # ---- e2e: host buffers in, host buffer out, every step ---------------------------------------
        if not args.no_e2e:
            h_off = csr.offsets.cpu().pin_memory()
            h_col = csr.indices.cpu().pin_memory()
            h_depth = torch.empty(n, dtype=torch.int32).pin_memory()
            h_src = torch.tensor(timed, dtype=torch.int32).pin_memory()
            d_src = torch.empty(1, dtype=torch.int32, device=dev)
            f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            with torch.cuda.stream(stream):
                f0.record(stream)
                for i, s in enumerate(timed):
                    csr.offsets.copy_(h_off, non_blocking=True)
                    csr.indices.copy_(h_col, non_blocking=True)
                    d_src.copy_(h_src[i:i + 1], non_blocking=True)
                    one_bfs(s)
                    h_depth.copy_(depth, non_blocking=True)
                f1.record(stream)
            torch.cuda.synchronize()
            ems = f0.elapsed_time(f1)
            out["e2e"] = {"value": edges / (ems * 1e-3) / 1e9, "unit": "GTEPS",
                          "h2d_bytes_per_step": int(h_off.numel() * h_off.element_size() + h_col.numel() * 4 + 4),
                          "d2h_bytes_per_step": int(n * 4), "ms_per_step": ems / K,
                          "what": "per step: CSR offsets+indices and the source H2D from pinned memory, ess_bfs, depth D2H"}
            last_depth = h_depth.numpy().copy()
        else:
            last_depth = depth.cpu().numpy()

        # ---- CPU baseline: the reference's bfs_cpu, one source of the same graph -------------------------
        if not args.no_cpu and rank == 0:
            off = csr.offsets.cpu().numpy()
            col = csr.indices.cpu().numpy()
            s = timed[-1]
            d_cpu, cpu_ms, kind = cpu_reference_bfs(off, col, s)
            same = bool(np.array_equal(d_cpu, last_depth))
            out["cpu_baseline"] = {
                "value": work[s][1] / (cpu_ms * 1e-3) / 1e9, "unit": "GTEPS", "cores": 1, "kind": kind,
                "sample": f"1 BFS (source {s}) on the full workload graph, search time only ({cpu_ms / 1e3:.1f} s)",
                "depths_equal_gpu": same}
            if not same:
                out["parity_error"] = "GPU depths differ from the CPU reference"
        # ---- side measurements (BASELINE configs 3 and 4, and the reference's own GPU path); never fatal ----
        if not args.no_extras and rank == 0:
            try:
                out["other_configs"] = side_runs(args, ctx, csr, graph, timed, work, dev)
            except Exception as e:  # the headline line must survive anything here
                out["other_configs"] = {"error": repr(e)[:300]}
    else:
        # ---- multi-GPU: NVLink-side accounting + e2e with this rank's partition in host memory ---------
        kind = runner.exchange_kind() if hasattr(runner, "exchange_kind") else "nccl"
        how = ("the library's own peer-memory kernels: 8-byte stores into the receivers' IPC-mapped windows over "
               "NVLink + epoch flags (NCCL only bootstraps)" if kind == "peer-memory" else "NCCL over NVLink")
        out["config"]["exchange"] = ("per level: all_to_all of candidate bitmap slices (top-down levels only) + one "
                                     "all_gather of the next-frontier slice and Beamer counters; " + how)
        out["config"]["exchange_kind"] = kind
        out["config"]["levels"] = runner.levels
        out["config"]["pull_levels"] = runner.pull_levels
        nv_bytes = runner.bytes_exchanged  # received per rank in the last BFS
        out["roofline"] = {"bound": "hbm", "kernel": "partitioned level kernels", "achieved": None, "peak": peak_gbs,
                           "unit": "GB/s", "frac": None, "traffic": None, "peak_source": peak_src,
                           "nvlink_bytes_received_per_rank_per_bfs": nv_bytes,
                           "nvlink_time_floor_ms": nv_bytes / 770e9 * 1e3,
                           "note": "multi-GPU steps are latency-bound by the per-level exchange; single-GPU run "
                                   "carries the kernel roofline"}
        if not args.no_e2e:
            csr = runner.csr
            h_off, h_col = csr.offsets.cpu().pin_memory(), csr.indices.cpu().pin_memory()
            h_depth = torch.empty(runner.per, dtype=torch.int32).pin_memory()
            f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            dist.barrier()
            torch.cuda.synchronize()
            with torch.cuda.stream(stream):
                f0.record(stream)
                for s in timed:
                    csr.offsets.copy_(h_off, non_blocking=True)
                    csr.indices.copy_(h_col, non_blocking=True)
                    one_bfs(s)
                    h_depth.copy_(runner.depth_local, non_blocking=True)
                f1.record(stream)
            torch.cuda.synchronize()
            t = torch.tensor([f0.elapsed_time(f1)], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ems = float(t.item())
            out["e2e"] = {"value": edges / (ems * 1e-3) / 1e9, "unit": "GTEPS",
                          "h2d_bytes_per_step": int(h_off.numel() * h_off.element_size() + h_col.numel() * 4 + 4) * world,
                          "d2h_bytes_per_step": int(n * 4), "ms_per_step": ems / K,
                          "what": "per step and rank: this rank's CSR partition H2D from pinned memory, the BFS, "
                                  "the owned depth slice D2H (bytes summed over ranks)"}

    if rank == 0:
        print(json.dumps(out), flush=True)


def side_runs(args, ctx, csr, graph, timed, work, dev):
    """BASELINE configs 3/4 and the reference GPU implementation on this same device (one run each)."""
    import torch

    import essentials_b200 as ess
    import oracle
    from essentials_b200 import graphgen as gg
    res = {}
    # the reference's own GPU BFS (block_mapped + Thrust, built for sm_100) on the bench graph, same source
    if oracle.have_ref_gpu() and csr.offsets.dtype == torch.int32:
        s = timed[0]
        for _ in range(2):
            ref_depth, ref_ms = oracle.ref_gpu_run("bfs", csr, s)
        ours, info = ess.bfs(ctx, graph, s, lb=args.lb, direction=args.direction)
        res["reference_gpu_bfs"] = {"workload": f"same graph, source {s}", "reference_enact_ms": ref_ms,
                                    "reference_gteps": work[s][1] / ref_ms / 1e6, "ours_enact_ms": info["enact_ms"],
                                    "ours_gteps": work[s][1] / info["enact_ms"] / 1e6,
                                    "depths_equal": bool(torch.equal(ours, ref_depth))}
        del ref_depth
    # SSSP on the bench graph itself (weights from the symmetric pair hash): dense delta variant vs the frontier recipe
    if csr.symmetric:
        from dataclasses import replace
        deg = (csr.offsets[1:] - csr.offsets[:-1]).to(torch.int64)
        rows = torch.repeat_interleave(torch.arange(csr.n, device=dev, dtype=torch.int64), deg)
        weighted = replace(csr, values=gg.pair_weights(rows, csr.indices.to(torch.int64)))
        del rows, deg
        wg = ess.Graph(weighted)
        runs = []
        for s in timed[:3]:
            for _ in range(2):  # second run: buffers warm
                d_delta, i_delta = ess.sssp_delta(ctx, wg, s)
                d_front, i_front = ess.sssp(ctx, wg, s, lb="merge_path")
            runs.append({"source": s, "delta_enact_ms": i_delta["enact_ms"], "delta_rounds": i_delta["rounds"],
                         "delta_gteps": work[s][1] / i_delta["enact_ms"] / 1e6,
                         "frontier_merge_path_enact_ms": i_front["enact_ms"],
                         "distances_equal": bool(torch.equal(d_delta, d_front))})
        res["sssp_kron_bench_graph"] = {"weights": "uniform [1,64) dyadic, symmetric hash", "runs": runs}
        del wg, weighted, d_delta, d_front
    # config 3: SSSP on the 4900 x 4900 grid
    grid = gg.grid_csr(4900, 4900, device=dev)
    gg_graph = ess.Graph(grid)
    d_nf, i_nf = ess.sssp_near_far(ctx, gg_graph, 0)
    d_lc, i_lc = ess.sssp(ctx, gg_graph, 0, lb="block_mapped")
    res["sssp_grid_4900"] = {"n": grid.n, "m": grid.m, "near_far_enact_ms": i_nf["enact_ms"],
                             "near_far_gteps": grid.m / i_nf["enact_ms"] / 1e6, "near_far_levels": i_nf["levels"],
                             "relaxations_per_edge": i_nf["relaxations"] / grid.m,
                             "label_correcting_block_mapped_enact_ms": i_lc["enact_ms"],
                             "distances_equal": bool(torch.equal(d_nf, d_lc))}
    if oracle.have_ref_gpu():
        ref_dist, ref_ms = oracle.ref_gpu_run("sssp", grid, 0)
        res["sssp_grid_4900"].update(reference_gpu_enact_ms=ref_ms, equal_reference_gpu=bool(torch.equal(d_nf, ref_dist)))
    del gg_graph, grid, d_nf, d_lc
    # config 4: PageRank on directed RMAT scale-25
    pr_csr = gg.rmat_csr(25, args.edge_factor, symmetric=False, weights="ones", device=dev)
    pr_graph = ess.Graph(pr_csr, csc=ess.transpose(pr_csr))
    p_pull, i_pull = ess.pagerank(ctx, pr_graph, pull=True)
    p_push, i_push = ess.pagerank(ctx, pr_graph, lb="merge_path")
    rel = ((p_pull.double() - p_push.double()).abs().sum() / p_pull.double().sum()).item()
    res["pagerank_rmat25"] = {"n": pr_csr.n, "m": pr_csr.m, "iterations": i_pull["iterations"],
                              "pull_ms_per_iteration": i_pull["enact_ms"] / max(i_pull["iterations"], 1),
                              "push_merge_path_ms_per_iteration": i_push["enact_ms"] / max(i_push["iterations"], 1),
                              "pull_gteps": pr_csr.m * i_pull["iterations"] / i_pull["enact_ms"] / 1e6,
                              "rel_l1_push_vs_pull": rel}
    if oracle.have_ref_gpu():
        ref_p, ref_ms = oracle.ref_gpu_run("pr", pr_csr, 0.85, 1e-6)
        res["pagerank_rmat25"].update(reference_gpu_enact_ms=ref_ms, ours_pull_enact_ms=i_pull["enact_ms"],
                                      rel_l1_vs_reference_gpu=((p_pull.double() - ref_p.double()).abs().sum()
                                                               / ref_p.double().sum()).item())
    return res


def main():
    args = parse()
    world_env = int(os.environ.get("WORLD_SIZE", "1"))
    args.scaling = "strong" if args.scale > 0 and world_env > 1 else "weak"
    if args.scale <= 0:
        args.scale = 24 + max(world_env.bit_length() - 1, 0)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        run_b200(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
