#!/bin/bash
# round-2 GPU call A: A/B timings of the quad / chunk engines against the round-1 kernels
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
export PYTHONPATH=$PWD
mkdir -p gpurun_out
timeout 600 python scripts/probe_bfs.py --scale 24 --sources 4 \
  --variants block_mapped:forward,merge_path:forward,bucketing:forward,merge_path:optimized,block_mapped:optimized \
  --engines 00,11 > gpurun_out/r02a_probe24.log 2>&1
timeout 900 python scripts/probe_bfs.py --scale 26 --sources 6 --variants merge_path:optimized \
  --engines 00,01,10,11 > gpurun_out/r02a_probe26.log 2>&1
tail -3 gpurun_out/r02a_probe26.log
