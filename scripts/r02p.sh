#!/bin/bash
# Round-2 call P (4 GPUs): the driver's own N=4 launch of bench.py on the final tree.
export PYTHONPATH=$PWD
mkdir -p gpurun_out
timeout 280 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29612 \
  bench.py --gpus 4 --steps 8 --warmup 3 > gpurun_out/r02p_bench_n4.json 2> gpurun_out/r02p_bench_n4.err
echo "exit $?"; tail -c 3000 gpurun_out/r02p_bench_n4.json; tail -5 gpurun_out/r02p_bench_n4.err
