#!/usr/bin/env bash
# per-phase cycle counts of graph_search_kernel (profile build), c2 and c3, on a prepared graph; restores the product build
cd "$(dirname "$0")/.."
python tools/profile_search.py --prepare --workload c3 2>&1 | tail -1
JV_NVCC_EXTRA="-DJV_SEARCH_PROFILE" python jvector_b200/build.py --force > /dev/null 2>&1
python tools/profile_search.py --run --workload c2 --reps 2 2>&1 | grep -E "profile|device_ms" | tail -3
python tools/profile_search.py --run --workload c3 --reps 2 2>&1 | grep -E "profile|device_ms" | tail -3
python jvector_b200/build.py --force > /dev/null 2>&1
