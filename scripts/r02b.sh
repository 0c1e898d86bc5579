#!/bin/bash
# round-2 GPU call B: hub-kernel fix timing + ncu --set full of the pull and push kernels
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
export PYTHONPATH=$PWD
mkdir -p gpurun_out
timeout 600 python scripts/probe_bfs.py --scale 24 --sources 3 \
  --variants block_mapped:forward,bucketing:forward,merge_path:forward,block_mapped:optimized --engines 11 \
  > gpurun_out/r02b_probe24.log 2>&1
P26="python scripts/probe_bfs.py --scale 26 --sources 2 --variants merge_path:optimized --engines 11"
timeout 600 $P26 > gpurun_out/r02b_probe26.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'pull_chunk|merge_path_quad' -c 14 \
  -o gpurun_out/r02b_bfs26 $P26 > gpurun_out/r02b_ncu26.log 2>&1
P24="python scripts/probe_bfs.py --scale 24 --sources 1 --variants merge_path:forward --engines 11"
timeout 600 $P24 > gpurun_out/r02b_probe24f.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'merge_path_quad' -c 7 \
  -o gpurun_out/r02b_fwd24 $P24 > gpurun_out/r02b_ncu24.log 2>&1
ls -la gpurun_out | tail -12
