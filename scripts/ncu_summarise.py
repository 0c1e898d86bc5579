"""Summarise an `ncu --csv` launch list of a bench.py run.

    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
        --csv --log-file gpurun_out/launches.csv python bench.py --steps 4 --warmup 3 --no-extras --no-cpu --no-e2e
    python scripts/ncu_summarise.py gpurun_out/launches.csv --shares profiles/rNN_launch_shares.txt \
        --traffic profiles/ncu_traffic.json

Per kernel class (the same classes gunrock::gcuda::profiler_t times live inside bench.py) it reports the number of
launches, the share of the summed kernel time and the MEAN DRAM bytes per launch over every launch of the class in
that command — the per-launch figure bench.py's `roofline.traffic` quotes next to its algorithmic bytes per launch
(both averaged over all levels, light and heavy).  ncu's times are cold-cache and serialised, so only the shares
are comparable with the live CUDA-event numbers.
"""
import argparse
import collections
import csv
import json
import re

CLASSES = [  # first match wins
    ("pull_step", r"pull_step_kernel|pull_chunk_kernel"),
    ("push_expand", r"merge_path_kernel|merge_path_small_kernel|merge_path_quad_kernel|merge_path_small_quad_kernel|"
                    r"thread_mapped|block_mapped|warp_mapped|bucket|big_list|expand_.*kernel"),
    ("work_prepare", r"prepare_work_kernel|prepare_quads_kernel|max_degree_kernel|degree_sum|bin_"),
    ("dense_state", r"init_visited_kernel|gather_bits_kernel|scatter_bits_kernel|mark_frontier_kernel|fill_kernel"),
    ("counters", r"publish_counters_kernel"),
    ("setup", r"pull_hints_kernel|isolated_bitmap_kernel|transpose"),
    ("graph_generation", r"at_cuda_detail|native::|kernelHistogram|at::"),  # torch kernels of graphgen.py, untimed
]
UNTIMED = ("setup", "graph_generation")


def classify(name: str) -> str:
    for cls, pat in CLASSES:
        if re.search(pat, name):
            return cls
    return "other"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("csv")
    ap.add_argument("--shares", help="write the per-class table here")
    ap.add_argument("--traffic", help="update this ncu_traffic.json with mean DRAM bytes per launch per class")
    ap.add_argument("--label", default=None, help="provenance string stored with the traffic numbers")
    args = ap.parse_args()

    rows = list(csv.reader(line for line in open(args.csv) if line.startswith('"')))
    hdr = rows[0]
    col = {k: hdr.index(k) for k in ("ID", "Kernel Name", "Metric Name", "Metric Unit", "Metric Value")}
    launches = collections.OrderedDict()  # id -> {name, ns, rd, wr}
    for r in rows[1:]:
        rec = launches.setdefault(r[col["ID"]], {"name": r[col["Kernel Name"]], "ns": 0.0, "rd": 0.0, "wr": 0.0})
        val = float(r[col["Metric Value"]].replace(",", ""))
        unit = r[col["Metric Unit"]].lower()
        scale = {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "ns": 1.0, "us": 1e3, "ms": 1e6,
                 "usecond": 1e3, "msecond": 1e6, "nsecond": 1.0}.get(unit, 1.0)
        metric = r[col["Metric Name"]]
        if metric == "gpu__time_duration.sum":
            rec["ns"] = val * scale
        elif metric == "dram__bytes_read.sum":
            rec["rd"] = val * scale
        elif metric == "dram__bytes_write.sum":
            rec["wr"] = val * scale

    agg = collections.defaultdict(lambda: {"n": 0, "ns": 0.0, "bytes": 0.0, "names": collections.Counter()})
    for rec in launches.values():
        a = agg[classify(rec["name"])]
        a["n"] += 1
        a["ns"] += rec["ns"]
        a["bytes"] += rec["rd"] + rec["wr"]
        a["names"][re.sub(r"[<(].*", "", rec["name"]).replace("void ", "")] += 1
    traversal = [c for c in agg if c not in UNTIMED]
    total_ns = sum(agg[c]["ns"] for c in traversal) or 1.0

    lines = [f"# {args.csv}: {len(launches)} launches; shares exclude graph generation / set-up kernels",
             f"{'class':14s} {'launches':>8s} {'ms':>9s} {'share':>7s} {'MB/launch':>10s} {'GB/s':>8s}  kernels"]
    for c in sorted(agg, key=lambda c: -agg[c]["ns"]):
        a = agg[c]
        share = a["ns"] / total_ns if c not in UNTIMED else float("nan")
        gbs = a["bytes"] / a["ns"] if a["ns"] else 0.0
        lines.append(f"{c:14s} {a['n']:8d} {a['ns'] / 1e6:9.3f} {share:7.3f} {a['bytes'] / max(a['n'], 1) / 1e6:10.2f} "
                     f"{gbs:8.1f}  {', '.join(a['names'])}")
    text = "\n".join(lines)
    print(text)
    if args.shares:
        with open(args.shares, "w") as f:
            f.write(text + "\n")
    if args.traffic and any(a["bytes"] for a in agg.values()):
        try:
            cur = json.load(open(args.traffic))
        except Exception:
            cur = {}
        for c, a in agg.items():
            if a["bytes"] and c in ("pull_step", "push_expand", "work_prepare", "dense_state"):
                cur[c] = {"bytes_per_launch": a["bytes"] / a["n"], "launches": a["n"],
                          "ncu_ms_per_launch": a["ns"] / a["n"] / 1e6,
                          "source": args.label or f"{args.csv} (mean over all {a['n']} launches of the class)"}
        with open(args.traffic, "w") as f:
            json.dump(cur, f, indent=1)


if __name__ == "__main__":
    main()
