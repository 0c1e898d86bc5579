#!/usr/bin/env python
"""Summarise an `ncu --set full` report: one line per launch with the counters the roofline argument needs
(duration, DRAM bytes and GB/s, L1TEX / L2 throughput %, hit rates, issue utilisation, lanes per instruction,
top stall reasons, global-atomic counters).  Usage: ncu_key_metrics.py report.ncu-rep [out.txt]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}


def f(r, name, default=0.0):
    try:
        return float(r[idx[name]].replace(",", ""))
    except Exception:
        return default


def unit(name):
    return units[idx[name]] if name in idx else ""


out = []
for r in data:
    name = r[idx["Kernel Name"]]
    short = name.split("<")[0].replace("void ", "")
    dur = f(r, "gpu__time_duration.sum")
    dur_us = dur / 1e3 if unit("gpu__time_duration.sum") in ("ns", "nsecond") else dur
    rd, wr = f(r, "dram__bytes_read.sum"), f(r, "dram__bytes_write.sum")
    scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}
    rd *= scale.get(unit("dram__bytes_read.sum"), 1.0)
    wr *= scale.get(unit("dram__bytes_write.sum"), 1.0)
    stalls = {h.split("issue_stalled_")[1].split("_per")[0]: f(r, h) for h in hdr
              if "smsp__average_warps_issue_stalled" in h and h.endswith("per_issue_active.ratio")
              and "not_issued" not in h}
    top = sorted(stalls.items(), key=lambda kv: -kv[1])[:4]
    line = (f"{short:32s} {dur_us:9.1f} us | dram R {rd/1e6:8.1f} W {wr/1e6:7.1f} MB = {(rd+wr)/dur_us/1e3 if dur_us else 0:7.1f} GB/s "
            f"({f(r,'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):4.1f}% of ncu peak) | "
            f"l1tex {f(r,'l1tex__throughput.avg.pct_of_peak_sustained_active'):4.1f}% lsu-wavefronts {f(r,'l1tex__data_pipe_lsu_wavefronts.sum')/1e6:7.1f}M "
            f"| lts {f(r,'lts__throughput.avg.pct_of_peak_sustained_elapsed'):4.1f}% L2hit {f(r,'lts__t_sector_hit_rate.pct'):4.1f}% L1hit {f(r,'l1tex__t_sector_hit_rate.pct'):4.1f}% "
            f"| issue {f(r,'smsp__issue_active.avg.pct_of_peak_sustained_active'):4.1f}% lanes/inst {f(r,'smsp__thread_inst_executed_per_inst_executed.ratio'):4.1f} "
            f"warp-inst {f(r,'smsp__inst_executed.sum')/1e6:6.1f}M regs {int(f(r,'launch__registers_per_thread'))} "
            f"occ {f(r,'sm__warps_active.avg.pct_of_peak_sustained_active'):4.1f}% "
            f"| atom {f(r,'l1tex__t_set_accesses_pipe_lsu_mem_global_op_atom.sum')/1e6:6.2f}M red {f(r,'l1tex__t_set_accesses_pipe_lsu_mem_global_op_red.sum')/1e6:6.2f}M "
            f"| stalls " + " ".join(f"{k}={v:.1f}" for k, v in top))
    out.append(line)
text = "\n".join(out) + "\n"
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write(text)
print(text)
