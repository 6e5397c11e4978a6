#!/usr/bin/env bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5) > gpurun_out/exp8_tests.log 2>&1
timeout 600 python tools/profile_search.py --sustain 2 --reps 10 --envs ";" > gpurun_out/exp8_default.log 2>&1
JV_B200_SO=$PWD/jvector_b200/lib_gencas/libjvector_b200.so timeout 600 python tools/profile_search.py --sustain 2 --reps 10 --envs ";" > gpurun_out/exp8_gencas.log 2>&1
JV_B200_SO=$PWD/jvector_b200/lib_minb9/libjvector_b200.so timeout 600 python tools/profile_search.py --sustain 2 --reps 10 --workloads c2 --envs ";" > gpurun_out/exp8_minb9.log 2>&1
tail -2 gpurun_out/exp8_tests.log; grep -H "^c[23] " gpurun_out/exp8_*.log
