#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
export PYTHONPATH=$PWD
mkdir -p gpurun_out
( time timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29613 \
    bench.py --gpus 8 --steps 16 --warmup 3 ) > gpurun_out/r02l_bench_n8.json 2> gpurun_out/r02l_bench_n8.err
tail -6 gpurun_out/r02l_bench_n8.err; cut -c1-300 gpurun_out/r02l_bench_n8.json | tail -3
