#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
export PYTHONPATH=$PWD
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "near_far or threshold" 2>&1 | tail -5 > gpurun_out/r02i_tests.log
timeout 300 python scripts/probe_grid.py --deltas 0,128,256 > gpurun_out/r02i_probe_grid.log 2>&1
tail -3 gpurun_out/r02i_tests.log; cat gpurun_out/r02i_probe_grid.log | tail -8
