#!/usr/bin/env bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
(JV_PQ_SCORE=thread timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15) > gpurun_out/exp3_tests.log 2>&1
timeout 900 python tools/profile_search.py --workloads c3 --envs ";JV_PQ_SCORE=thread;JV_PQ_SCORE=thread,JV_PQ_WIDE=1;JV_PQ_SCORE=thread,JV_PQ_WIDE=1,JV_PQ_LUT_SMEM_M=48;JV_PQ_SCORE=thread,JV_FUSED_PQ=0" > gpurun_out/exp3_c3.log 2>&1
JV_B200_SO=$PWD/jvector_b200/lib_t128/libjvector_b200.so timeout 900 python tools/profile_search.py --envs ";JV_PQ_SCORE=thread;JV_PQ_SCORE=thread,JV_PQ_WIDE=1" > gpurun_out/exp3_t128.log 2>&1
JV_B200_SO=$PWD/jvector_b200/lib_prof/libjvector_b200.so timeout 900 python tools/profile_search.py --reps 1 --workloads c3 --envs "JV_PQ_SCORE=thread;JV_PQ_SCORE=thread,JV_PQ_WIDE=1" > gpurun_out/exp3_prof.log 2>&1
timeout 600 python tools/profile_search.py --workload seam > gpurun_out/exp3_seam.log 2>&1
grep -h "^c[23] \|passed\|failed\|nq=10000\|score_ragged" gpurun_out/exp3_*.log
