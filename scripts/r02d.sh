#!/bin/bash
# round-2 GPU call D: parity suite (cluster near-far, atomics), grid SSSP probe, ncu launch list of the bench, bench lines
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
export PYTHONPATH=$PWD
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/r02d_tests.log
timeout 600 python scripts/probe_grid.py > gpurun_out/r02d_probe_grid.log 2>&1
B="python bench.py --steps 4 --warmup 3 --no-extras --no-cpu --no-e2e"
timeout 600 $B > gpurun_out/r02d_bench_short.json 2> gpurun_out/r02d_bench_short.err && \
timeout 1200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
  --csv --log-file gpurun_out/r02d_launches_bfs_kron26.csv $B > gpurun_out/r02d_ncu_launches.log 2>&1
( time timeout 1500 python bench.py ) > gpurun_out/r02d_bench_n1.json 2> gpurun_out/r02d_bench_n1.err
( time timeout 1500 python bench.py --impl reference ) > gpurun_out/r02d_bench_ref.json 2> gpurun_out/r02d_bench_ref.err
tail -4 gpurun_out/r02d_tests.log; cat gpurun_out/r02d_probe_grid.log | tail -12; tail -3 gpurun_out/r02d_bench_n1.err gpurun_out/r02d_bench_ref.err
