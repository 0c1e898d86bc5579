#!/bin/bash
# Round-2 call S (1 GPU): final tree — smoke (incl. the host-array graph build), full GPU suite, the default bench line.
export PYTHONPATH=$PWD
mkdir -p gpurun_out
timeout 120 python __graft_entry__.py smoke > gpurun_out/r02s_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r02s_smoke.log
timeout 200 python -m pytest tests -m gpu -x -q > gpurun_out/r02s_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r02s_pytest.log
timeout 400 python bench.py > gpurun_out/r02s_bench_n1.json 2> gpurun_out/r02s_bench_n1.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02s_bench_n1.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','parity_ok','gpu_launches')}); print(d['e2e']['value'], d['e2e']['ms_per_step'], d['cpu_baseline']['value'])
print(d['roofline']['frac'], d['roofline']['kernel_share_of_step'])
oc=d['other_configs']; print({k:(list(v) if isinstance(v,dict) else v) for k,v in oc.items()})
PY
tail -3 gpurun_out/r02s_bench_n1.err
