#!/bin/bash
# round-2 GPU call H: parity suite + grid SSSP after the two-phase batching
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
export PYTHONPATH=$PWD
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/r02h_tests.log
timeout 300 python scripts/probe_grid.py --deltas 0,64,128,256 > gpurun_out/r02h_probe_grid.log 2>&1
tail -4 gpurun_out/r02h_tests.log; cat gpurun_out/r02h_probe_grid.log | tail -8
