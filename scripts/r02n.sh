#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
export PYTHONPATH=$PWD
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py smoke > gpurun_out/r02n_smoke.log 2>&1
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "bfs or advance or frontier" 2>&1 | tail -4 > gpurun_out/r02n_tests.log
( time timeout 600 python bench.py --steps 6 --warmup 3 --no-extras --no-cpu ) > gpurun_out/r02n_bench_short.json 2> gpurun_out/r02n_bench_short.err
tail -1 gpurun_out/r02n_smoke.log; tail -2 gpurun_out/r02n_tests.log; tail -4 gpurun_out/r02n_bench_short.err
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/r02n_bench_short.json").read().splitlines() if l.startswith("{")][-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["parity_ok"])
PY
