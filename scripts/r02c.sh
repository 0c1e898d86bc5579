#!/bin/bash
# round-2 GPU call C: parity suite on pull v3 / degree-sum epilogue, probes, first run of the new bench.py
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
export PYTHONPATH=$PWD
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/r02c_tests.log
timeout 600 python scripts/probe_bfs.py --scale 26 --sources 6 --variants merge_path:optimized --engines 11 \
  --alphas 14,30,60 > gpurun_out/r02c_probe26.log 2>&1
timeout 600 python scripts/probe_bfs.py --scale 24 --sources 4 --variants merge_path:optimized,block_mapped:forward,merge_path:forward \
  --engines 11 > gpurun_out/r02c_probe24.log 2>&1
( time timeout 1200 python bench.py --steps 8 --warmup 3 ) > gpurun_out/r02c_bench_n1.json 2> gpurun_out/r02c_bench_n1.err
( time timeout 900 python bench.py --impl reference --steps 4 --warmup 1 ) > gpurun_out/r02c_bench_ref.json 2> gpurun_out/r02c_bench_ref.err
tail -3 gpurun_out/r02c_tests.log; tail -4 gpurun_out/r02c_bench_n1.err; tail -4 gpurun_out/r02c_bench_ref.err; nproc; free -g | head -2
