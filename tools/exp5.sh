#!/usr/bin/env bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15) > gpurun_out/exp5_tests.log 2>&1
(JV_PQ_SCORE=group JV_VISITED=global timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15) > gpurun_out/exp5_tests_old.log 2>&1
timeout 900 python tools/profile_search.py --envs ";JV_PQ_SCORE=group;JV_SEARCH_WIDE=1" > gpurun_out/exp5_time.log 2>&1
NCU="ncu --set full --clock-control none --profile-from-start off -f"
timeout 900 $NCU --import-source on -k regex:graph_search_kernel -c 1 -o gpurun_out/r2b_search_c2 python tools/profile_search.py --workload c2 --reps 1 --ncu > gpurun_out/exp5_ncu_c2.log 2>&1
timeout 900 $NCU -k regex:graph_search_kernel -c 1 -o gpurun_out/r2b_search_c3 python tools/profile_search.py --workload c3 --reps 1 --ncu > gpurun_out/exp5_ncu_c3.log 2>&1
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2b_launches_c2.csv python bench.py --workload c2 --steps 2 --warmup 3 --no-cpu --gt-queries 100 > gpurun_out/exp5_launches.log 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/*.csv
grep -h "^c[23] \|passed\|failed" gpurun_out/exp5_*.log
