#!/usr/bin/env bash
# Round-end evidence in one gpurun call (one GPU): GPU tests, launch list of the timed region of the default bench workload,
# the default bench line, the reference arm.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5) > gpurun_out/final_tests.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2b_launches_bench_c2.csv python bench.py --workload c2 --steps 2 --warmup 3 --no-cpu --gt-queries 100 --ncu-range > gpurun_out/final_launches.log 2>&1
(time python bench.py) > gpurun_out/r2b_bench_n1.json 2> gpurun_out/r2b_bench_n1.err
(time python bench.py --impl reference) > gpurun_out/r2b_bench_n1_reference.json 2> gpurun_out/r2b_bench_n1_reference.err
tail -2 gpurun_out/final_tests.log; tail -4 gpurun_out/r2b_launches_bench_c2.csv | cut -c1-300; python tools/bench_table.py gpurun_out/r2b_bench_n1.json gpurun_out/r2b_bench_n1_reference.json | cut -c1-330
