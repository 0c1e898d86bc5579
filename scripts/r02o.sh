#!/bin/bash
# Round-2 call O (2 GPUs): the driver's own N=2 launch of bench.py on the final tree.
export PYTHONPATH=$PWD
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 \
  bench.py --gpus 2 --steps 8 --warmup 3 > gpurun_out/r02o_bench_n2.json 2> gpurun_out/r02o_bench_n2.err
echo "exit $?"; tail -c 3000 gpurun_out/r02o_bench_n2.json; tail -5 gpurun_out/r02o_bench_n2.err
