#!/bin/bash
# Round-2 call Q (1 GPU): full GPU suite + smoke + short bench on the tree with the host-array graph build,
# the lane-per-vertex hint kernel and the stream-aware device pool.
export PYTHONPATH=$PWD
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/r02q_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02q_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02q_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r02q_smoke.log
timeout 300 python bench.py --steps 6 --warmup 3 --no-extras --no-cpu > gpurun_out/r02q_bench.json 2> gpurun_out/r02q_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02q_bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','parity_ok','gpu_launches')}); print(d['e2e'])
PY
tail -3 gpurun_out/r02q_bench.err
