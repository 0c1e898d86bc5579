#!/bin/bash
# round-2 GPU call G (2 GPUs): multi-GPU parity tests (native / NCCL / python loops, partitioned enactor), bench at N=2
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
export PYTHONPATH=$PWD
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_dist.py -m gpu -q -x 2>&1 | tail -40 > gpurun_out/r02g_tests_dist.log
( time timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 \
    bench.py --gpus 2 --steps 8 --warmup 3 ) > gpurun_out/r02g_bench_n2.json 2> gpurun_out/r02g_bench_n2.err
tail -12 gpurun_out/r02g_tests_dist.log; tail -5 gpurun_out/r02g_bench_n2.err; cut -c1-600 gpurun_out/r02g_bench_n2.json
