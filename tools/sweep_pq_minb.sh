#!/usr/bin/env bash
cd "$(dirname "$0")/.."
python tools/profile_search.py --prepare --workload c3 2>&1 | tail -1
for v in 5 6 8; do
  JV_NVCC_EXTRA="-DJV_SEARCH_MINB_PQ=$v" python jvector_b200/build.py --force > /dev/null 2>&1
  regs=$(grep -A2 "graph_search_kernelILi1ELi1" jvector_b200/lib/ptxas_info.log | tail -1 | sed 's/.*Used \([0-9]*\) registers.*/\1/')
  echo "== PQ minb=$v regs=$regs"
  python tools/profile_search.py --run --workload c3 --reps 2 2>&1 | tail -1
done
python jvector_b200/build.py --force > /dev/null 2>&1
