#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
export PYTHONPATH=$PWD
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 \
  scripts/dist_check.py --scale 26 --sources 3 --compare-exchange --no-check > gpurun_out/r02j_dist_trace_n2.log 2>&1
grep -v "^\*\|OMP_NUM" gpurun_out/r02j_dist_trace_n2.log | tail -30
