#!/bin/bash
# Round-2 call T (1 GPU): the small all-kernels pass of scripts/small_pass.py (compute-sanitizer is closed on this
# pool, so it runs plain: every result is compared with the oracle).
export PYTHONPATH=$PWD
mkdir -p gpurun_out
timeout 35 python scripts/small_pass.py > gpurun_out/r02t_small_pass.log 2>&1
echo "rc=$?"; tail -3 gpurun_out/r02t_small_pass.log
