#!/usr/bin/env bash
# builds graph_search_kernel with different (threads, min blocks/SM) and times workload c2/c3 on a prepared graph
cd "$(dirname "$0")/.."
python tools/profile_search.py --prepare --workload c3 2>&1 | tail -1
for v in "256 4" "256 5" "256 6" "256 8" "128 8" "128 12" "128 16"; do
  set -- $v
  JV_NVCC_EXTRA="-DJV_SEARCH_THREADS=$1 -DJV_SEARCH_MINB=$2" python jvector_b200/build.py --force > /dev/null 2>&1
  regs=$(grep -A2 "graph_search_kernelILi0ELi1" jvector_b200/lib/ptxas_info.log | tail -1 | sed 's/.*Used \([0-9]*\) registers.*/\1/')
  echo "== threads=$1 minb=$2 regs(f32,dot)=$regs"
  python tools/profile_search.py --run --workload c2 2>&1 | tail -1
  python tools/profile_search.py --run --workload c3 2>&1 | tail -1
done
