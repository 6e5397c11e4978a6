#!/usr/bin/env bash
# Record of the gpurun batches behind profiles/r2b_search_experiments.md (round 2, second half). Each function is one gpurun call;
# the variant libraries are built first with tools/build_variant.sh <name> <-D flags> (jvector_b200/lib_<name>/, git- and gpurun-shipped
# only while they exist). Usage: tools/r2b_experiments.sh expN
cd "$(dirname "$0")/.."; mkdir -p gpurun_out; export PYTHONUNBUFFERED=1

exp1() {
  (JV_VISITED=smem JV_ROW_PREFETCH=1 timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15) > gpurun_out/exp1_tests.log 2>&1
  E='; JV_ROW_PREFETCH=1; JV_ROW_PREFETCH=2; JV_VISITED=smem; JV_VISITED=smem,JV_ROW_PREFETCH=1; JV_VISITED=smem,JV_VISITED_SMEM_SLOTS=16384,JV_ROW_PREFETCH=1'
  E=${E// /}
  timeout 900 python tools/profile_search.py --envs "$E" > gpurun_out/exp1_default.log 2>&1
  JV_B200_SO=$PWD/jvector_b200/lib_t128/libjvector_b200.so timeout 900 python tools/profile_search.py --envs ";JV_VISITED=smem;JV_VISITED=smem,JV_ROW_PREFETCH=1" > gpurun_out/exp1_t128.log 2>&1
  JV_B200_SO=$PWD/jvector_b200/lib_prof/libjvector_b200.so timeout 900 python tools/profile_search.py --reps 1 --envs ";JV_VISITED=smem,JV_ROW_PREFETCH=1" > gpurun_out/exp1_prof.log 2>&1
  tail -n 30 gpurun_out/exp1_tests.log gpurun_out/exp1_default.log gpurun_out/exp1_t128.log gpurun_out/exp1_prof.log
}

exp2() {
  (timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15) > gpurun_out/exp2_tests.log 2>&1
  timeout 900 python tools/profile_search.py --workloads c2 --envs ";JV_SEARCH_WIDE=1;JV_VISITED=global,JV_ROW_PREFETCH=0;JV_VISITED=global;JV_ROW_PREFETCH=0" > gpurun_out/exp2_c2.log 2>&1
  timeout 900 python tools/profile_search.py --workloads c3 --envs ";JV_PQ_LUT_SMEM_M=32;JV_PQ_LUT_SMEM_M=48;JV_PQ_LUT_SMEM_M=64;JV_PQ_LUT_SMEM_M=84;JV_VISITED=global" > gpurun_out/exp2_c3.log 2>&1
  JV_B200_SO=$PWD/jvector_b200/lib_t128/libjvector_b200.so timeout 900 python tools/profile_search.py --workloads c2 --envs ";JV_SEARCH_WIDE=1" > gpurun_out/exp2_t128.log 2>&1
  JV_B200_SO=$PWD/jvector_b200/lib_minb6/libjvector_b200.so timeout 900 python tools/profile_search.py --workloads c2 --envs ";" > gpurun_out/exp2_minb6.log 2>&1
  JV_B200_SO=$PWD/jvector_b200/lib_prof/libjvector_b200.so timeout 900 python tools/profile_search.py --reps 1 --envs ";JV_PQ_LUT_SMEM_M=84" > gpurun_out/exp2_prof.log 2>&1
  grep -h "^c[23] \|passed\|failed\|nq=10000" gpurun_out/exp2_*.log
}

exp3() {
  (JV_PQ_SCORE=thread timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15) > gpurun_out/exp3_tests.log 2>&1
  timeout 900 python tools/profile_search.py --workloads c3 --envs ";JV_PQ_SCORE=thread;JV_PQ_SCORE=thread,JV_PQ_WIDE=1;JV_PQ_SCORE=thread,JV_PQ_WIDE=1,JV_PQ_LUT_SMEM_M=48;JV_PQ_SCORE=thread,JV_FUSED_PQ=0" > gpurun_out/exp3_c3.log 2>&1
  JV_B200_SO=$PWD/jvector_b200/lib_t128/libjvector_b200.so timeout 900 python tools/profile_search.py --envs ";JV_PQ_SCORE=thread;JV_PQ_SCORE=thread,JV_PQ_WIDE=1" > gpurun_out/exp3_t128.log 2>&1
  JV_B200_SO=$PWD/jvector_b200/lib_prof/libjvector_b200.so timeout 900 python tools/profile_search.py --reps 1 --workloads c3 --envs "JV_PQ_SCORE=thread;JV_PQ_SCORE=thread,JV_PQ_WIDE=1" > gpurun_out/exp3_prof.log 2>&1
  timeout 600 python tools/profile_search.py --workload seam > gpurun_out/exp3_seam.log 2>&1
  grep -h "^c[23] \|passed\|failed\|nq=10000\|score_ragged" gpurun_out/exp3_*.log
}

exp4() {
  (JV_PQ_SCORE=rows timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15) > gpurun_out/exp4_tests.log 2>&1
  timeout 900 python tools/profile_search.py --workloads c3 --envs ";JV_PQ_SCORE=rows;JV_PQ_SCORE=rows,JV_PQ_WIDE=1;JV_PQ_SCORE=rows,JV_PQ_LUT_SMEM_M=48;JV_PQ_SCORE=rows,JV_PQ_WIDE=1,JV_PQ_LUT_SMEM_M=48;JV_PQ_SCORE=rows,JV_FUSED_PQ=0" > gpurun_out/exp4_c3.log 2>&1
  JV_B200_SO=$PWD/jvector_b200/lib_prof/libjvector_b200.so timeout 900 python tools/profile_search.py --reps 1 --workloads c3 --envs "JV_PQ_SCORE=rows;JV_PQ_SCORE=rows,JV_PQ_WIDE=1" > gpurun_out/exp4_prof.log 2>&1
  grep -h "^c[23] \|passed\|failed\|nq=10000\|score_ragged" gpurun_out/exp4_*.log
}

exp5() {
  (timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15) > gpurun_out/exp5_tests.log 2>&1
  (JV_PQ_SCORE=group JV_VISITED=global timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15) > gpurun_out/exp5_tests_old.log 2>&1
  timeout 900 python tools/profile_search.py --envs ";JV_PQ_SCORE=group;JV_SEARCH_WIDE=1" > gpurun_out/exp5_time.log 2>&1
  NCU="ncu --set full --clock-control none --profile-from-start off -f"
  timeout 900 $NCU --import-source on -k regex:graph_search_kernel -c 1 -o gpurun_out/r2b_search_c2 python tools/profile_search.py --workload c2 --reps 1 --ncu > gpurun_out/exp5_ncu_c2.log 2>&1
  timeout 900 $NCU -k regex:graph_search_kernel -c 1 -o gpurun_out/r2b_search_c3 python tools/profile_search.py --workload c3 --reps 1 --ncu > gpurun_out/exp5_ncu_c3.log 2>&1
  timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2b_launches_c2.csv python bench.py --workload c2 --steps 2 --warmup 3 --no-cpu --gt-queries 100 > gpurun_out/exp5_launches.log 2>&1
  ls -la gpurun_out/*.ncu-rep gpurun_out/*.csv
  grep -h "^c[23] \|passed\|failed" gpurun_out/exp5_*.log
}

exp6() {
  (timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15) > gpurun_out/exp6_tests.log 2>&1
  timeout 900 python tools/profile_search.py --envs ";JV_EARLY_PREFETCH=0;JV_EARLY_PREFETCH=1" > gpurun_out/exp6_time.log 2>&1
  JV_B200_SO=$PWD/jvector_b200/lib_prof/libjvector_b200.so timeout 900 python tools/profile_search.py --reps 1 --envs "JV_EARLY_PREFETCH=0;JV_EARLY_PREFETCH=1" > gpurun_out/exp6_prof.log 2>&1
  grep -h "^c[23] \|passed\|failed\|nq=10000" gpurun_out/exp6_*.log
}

exp7() {
  timeout 900 python tools/profile_search.py --sustain 2 --reps 10 --envs ";JV_SEARCH_OVERLAP=0;JV_ROW_PREFETCH=0;JV_VISITED=global;JV_EARLY_PREFETCH=0,JV_PQ_SCORE=group" > gpurun_out/exp7_default.log 2>&1
  JV_B200_SO=$PWD/jvector_b200/lib_t256/libjvector_b200.so timeout 900 python tools/profile_search.py --sustain 2 --reps 10 --workloads c2 --envs ";JV_SEARCH_WIDE=1" > gpurun_out/exp7_t256.log 2>&1
  grep -h "^c[23] " gpurun_out/exp7_*.log
}

exp8() {
  (timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5) > gpurun_out/exp8_tests.log 2>&1
  timeout 600 python tools/profile_search.py --sustain 2 --reps 10 --envs ";" > gpurun_out/exp8_default.log 2>&1
  JV_B200_SO=$PWD/jvector_b200/lib_gencas/libjvector_b200.so timeout 600 python tools/profile_search.py --sustain 2 --reps 10 --envs ";" > gpurun_out/exp8_gencas.log 2>&1
  JV_B200_SO=$PWD/jvector_b200/lib_minb9/libjvector_b200.so timeout 600 python tools/profile_search.py --sustain 2 --reps 10 --workloads c2 --envs ";" > gpurun_out/exp8_minb9.log 2>&1
  tail -2 gpurun_out/exp8_tests.log; grep -H "^c[23] " gpurun_out/exp8_*.log
}

"$@"
