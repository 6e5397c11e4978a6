#!/usr/bin/env bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python tools/profile_search.py --sustain 2 --reps 10 --envs ";JV_SEARCH_OVERLAP=0;JV_ROW_PREFETCH=0;JV_VISITED=global;JV_EARLY_PREFETCH=0,JV_PQ_SCORE=group" > gpurun_out/exp7_default.log 2>&1
JV_B200_SO=$PWD/jvector_b200/lib_t256/libjvector_b200.so timeout 900 python tools/profile_search.py --sustain 2 --reps 10 --workloads c2 --envs ";JV_SEARCH_WIDE=1" > gpurun_out/exp7_t256.log 2>&1
grep -h "^c[23] " gpurun_out/exp7_*.log
