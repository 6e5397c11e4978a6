#!/usr/bin/env bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15) > gpurun_out/exp2_tests.log 2>&1
timeout 900 python tools/profile_search.py --workloads c2 --envs ";JV_SEARCH_WIDE=1;JV_VISITED=global,JV_ROW_PREFETCH=0;JV_VISITED=global;JV_ROW_PREFETCH=0" > gpurun_out/exp2_c2.log 2>&1
timeout 900 python tools/profile_search.py --workloads c3 --envs ";JV_PQ_LUT_SMEM_M=32;JV_PQ_LUT_SMEM_M=48;JV_PQ_LUT_SMEM_M=64;JV_PQ_LUT_SMEM_M=84;JV_VISITED=global" > gpurun_out/exp2_c3.log 2>&1
JV_B200_SO=$PWD/jvector_b200/lib_t128/libjvector_b200.so timeout 900 python tools/profile_search.py --workloads c2 --envs ";JV_SEARCH_WIDE=1" > gpurun_out/exp2_t128.log 2>&1
JV_B200_SO=$PWD/jvector_b200/lib_minb6/libjvector_b200.so timeout 900 python tools/profile_search.py --workloads c2 --envs ";" > gpurun_out/exp2_minb6.log 2>&1
JV_B200_SO=$PWD/jvector_b200/lib_prof/libjvector_b200.so timeout 900 python tools/profile_search.py --reps 1 --envs ";JV_PQ_LUT_SMEM_M=84" > gpurun_out/exp2_prof.log 2>&1
grep -h "^c[23] \|passed\|failed\|nq=10000" gpurun_out/exp2_*.log
