#!/usr/bin/env bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
(JV_PQ_SCORE=rows timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15) > gpurun_out/exp4_tests.log 2>&1
timeout 900 python tools/profile_search.py --workloads c3 --envs ";JV_PQ_SCORE=rows;JV_PQ_SCORE=rows,JV_PQ_WIDE=1;JV_PQ_SCORE=rows,JV_PQ_LUT_SMEM_M=48;JV_PQ_SCORE=rows,JV_PQ_WIDE=1,JV_PQ_LUT_SMEM_M=48;JV_PQ_SCORE=rows,JV_FUSED_PQ=0" > gpurun_out/exp4_c3.log 2>&1
JV_B200_SO=$PWD/jvector_b200/lib_prof/libjvector_b200.so timeout 900 python tools/profile_search.py --reps 1 --workloads c3 --envs "JV_PQ_SCORE=rows;JV_PQ_SCORE=rows,JV_PQ_WIDE=1" > gpurun_out/exp4_prof.log 2>&1
grep -h "^c[23] \|passed\|failed\|nq=10000\|score_ragged" gpurun_out/exp4_*.log
