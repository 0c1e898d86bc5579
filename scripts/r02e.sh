#!/bin/bash
# round-2 GPU call E: parity suite, grid SSSP probe after batching, PageRank balancers, ncu launch list + full capture
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
export PYTHONPATH=$PWD
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/r02e_tests.log
timeout 300 python scripts/probe_grid.py --deltas 0,128,256,512 > gpurun_out/r02e_probe_grid.log 2>&1
timeout 400 python scripts/probe_configs.py --skip-sssp > gpurun_out/r02e_probe_pr.log 2>&1
B="python bench.py --steps 4 --warmup 3 --no-extras --no-cpu --no-e2e"
timeout 600 $B > gpurun_out/r02e_bench_short.json 2> gpurun_out/r02e_bench_short.err && \
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
  --kernel-name-base demangled -k regex:gunrock --csv --log-file gpurun_out/r02e_launches_bfs_kron26.csv $B \
  > gpurun_out/r02e_ncu_launches.log 2>&1
P26="python scripts/probe_bfs.py --scale 26 --sources 1 --variants merge_path:optimized --engines 11"
timeout 600 $P26 > gpurun_out/r02e_probe26.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
  -k regex:'pull_chunk|merge_path_quad' -c 8 -o gpurun_out/r02e_bfs26 $P26 > gpurun_out/r02e_ncu26.log 2>&1
tail -4 gpurun_out/r02e_tests.log; cat gpurun_out/r02e_probe_grid.log | tail -8; grep "^pr " gpurun_out/r02e_probe_pr.log
