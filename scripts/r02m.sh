#!/bin/bash
# round-2 GPU call M: final single-GPU validation — smoke, full parity suite, the default bench line
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
export PYTHONPATH=$PWD
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py smoke > gpurun_out/r02m_smoke.log 2>&1
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/r02m_tests.log
( time timeout 900 python bench.py ) > gpurun_out/r02m_bench_n1.json 2> gpurun_out/r02m_bench_n1.err
tail -2 gpurun_out/r02m_smoke.log; tail -5 gpurun_out/r02m_tests.log; tail -4 gpurun_out/r02m_bench_n1.err; cut -c1-200 gpurun_out/r02m_bench_n1.json | tail -2
