#!/bin/bash
# round-2 GPU call F: where does a near-far level spend its time? ncu --set full with source counters on a 1500^2 grid
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
export PYTHONPATH=$PWD
mkdir -p gpurun_out
P="python scripts/probe_grid.py --grid 1500 --deltas 0 --modes 1 --reps 1"
timeout 300 $P > gpurun_out/r02f_probe.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
  -k regex:near_far_cluster -c 1 -o gpurun_out/r02f_nearfar $P > gpurun_out/r02f_ncu.log 2>&1
timeout 300 python scripts/probe_bfs.py --scale 24 --sources 4 --variants merge_path:forward,block_mapped:forward --engines 11 > gpurun_out/r02f_probe24.log 2>&1
cat gpurun_out/r02f_probe.log; cut -c1-200 gpurun_out/r02f_probe24.log
