#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
export PYTHONPATH=$PWD
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_dist.py -m gpu -q -x 2>&1 | tail -30 > gpurun_out/r02k_tests_dist.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 \
  scripts/dist_check.py --scale 26 --sources 3 --compare-exchange > gpurun_out/r02k_dist_trace_n2.log 2>&1
tail -5 gpurun_out/r02k_tests_dist.log; grep -v "^\*\|OMP_NUM" gpurun_out/r02k_dist_trace_n2.log | tail -22
