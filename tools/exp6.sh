#!/usr/bin/env bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15) > gpurun_out/exp6_tests.log 2>&1
timeout 900 python tools/profile_search.py --envs ";JV_EARLY_PREFETCH=0;JV_EARLY_PREFETCH=1" > gpurun_out/exp6_time.log 2>&1
JV_B200_SO=$PWD/jvector_b200/lib_prof/libjvector_b200.so timeout 900 python tools/profile_search.py --reps 1 --envs "JV_EARLY_PREFETCH=0;JV_EARLY_PREFETCH=1" > gpurun_out/exp6_prof.log 2>&1
grep -h "^c[23] \|passed\|failed\|nq=10000" gpurun_out/exp6_*.log
