#!/usr/bin/env bash
# Round-2 ncu captures (one GPU, run under gpurun): one `--set full` capture per dominant kernel, written to gpurun_out/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on --profile-from-start off -f"
$NCU -k regex:graph_search_kernel -c 1 -o gpurun_out/r2_search_c2 python tools/profile_search.py --workload c2 --reps 1 --ncu > gpurun_out/r2_ncu_c2.log 2>&1
$NCU -k regex:graph_search_kernel -c 1 -o gpurun_out/r2_search_c3 python tools/profile_search.py --workload c3 --reps 1 --ncu > gpurun_out/r2_ncu_c3.log 2>&1
$NCU -k regex:score_ragged_kernel -c 1 -o gpurun_out/r2_seam python tools/profile_search.py --workload seam --ncu > gpurun_out/r2_ncu_seam.log 2>&1
$NCU -k regex:bq_imma_kernel -c 2 -o gpurun_out/r2_bq_imma python tools/profile_c4.py --ncu > gpurun_out/r2_ncu_c4.log 2>&1
JV_BQ_FILTER=umma $NCU -k regex:bq_umma_filter_kernel -c 1 -o gpurun_out/r2_bq_umma python tools/profile_c4.py --ncu > gpurun_out/r2_ncu_c4u.log 2>&1
ls -la gpurun_out/*.ncu-rep
