#!/bin/bash
# Round-2 call R (2 GPUs): the N=2 bench line and one small SSSP/BFS dist_check on the tree with the wall-clock
# peer-wait timeout.
export PYTHONPATH=$PWD
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29613 \
  bench.py --gpus 2 --steps 8 --warmup 3 > gpurun_out/r02r_bench_n2.json 2> gpurun_out/r02r_bench_n2.err
echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02r_bench_n2.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','parity_ok','gpu_launches')}); print(d['e2e']['value'], d['e2e']['ms_per_step'], d['sssp']['gteps'], d['other_configs'])
PY
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29614 \
  scripts/dist_check.py --scale 17 --alg bfs,sssp > gpurun_out/r02r_dist_check.log 2>&1
echo "dist_check rc=$?"; grep -c "DIST_CHECK_OK" gpurun_out/r02r_dist_check.log; tail -3 gpurun_out/r02r_dist_check.log
