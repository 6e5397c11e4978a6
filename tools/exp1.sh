#!/usr/bin/env bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
(JV_VISITED=smem JV_ROW_PREFETCH=1 timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15) > gpurun_out/exp1_tests.log 2>&1
E='; JV_ROW_PREFETCH=1; JV_ROW_PREFETCH=2; JV_VISITED=smem; JV_VISITED=smem,JV_ROW_PREFETCH=1; JV_VISITED=smem,JV_VISITED_SMEM_SLOTS=16384,JV_ROW_PREFETCH=1'
E=${E// /}
timeout 900 python tools/profile_search.py --envs "$E" > gpurun_out/exp1_default.log 2>&1
JV_B200_SO=$PWD/jvector_b200/lib_t128/libjvector_b200.so timeout 900 python tools/profile_search.py --envs ";JV_VISITED=smem;JV_VISITED=smem,JV_ROW_PREFETCH=1" > gpurun_out/exp1_t128.log 2>&1
JV_B200_SO=$PWD/jvector_b200/lib_prof/libjvector_b200.so timeout 900 python tools/profile_search.py --reps 1 --envs ";JV_VISITED=smem,JV_ROW_PREFETCH=1" > gpurun_out/exp1_prof.log 2>&1
tail -n 30 gpurun_out/exp1_tests.log gpurun_out/exp1_default.log gpurun_out/exp1_t128.log gpurun_out/exp1_prof.log
