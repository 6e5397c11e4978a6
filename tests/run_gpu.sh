#!/usr/bin/env bash
# Runs every GPU test in its own process under a hard timeout, so a hung kernel costs one test, not the whole lease.
# usage: tests/run_gpu.sh [per-test-seconds] [pytest -k expression]
cd "$(dirname "$0")/.."
T=${1:-240}
K=${2:-}
mkdir -p gpurun_out
LOG=gpurun_out/gpu_tests.log
: > "$LOG"
ids=$(python -m pytest tests -m gpu --collect-only -q ${K:+-k "$K"} 2>/dev/null | grep "::")
pass=0; fail=0
for id in $ids; do
  if timeout "$T" python -m pytest "$id" -q -x --no-header -p no:cacheprovider >> "$LOG" 2>&1; then pass=$((pass+1)); echo "PASS $id";
  else fail=$((fail+1)); echo "FAIL($?) $id"; fi
done
echo "gpu tests: $pass passed, $fail failed"
[ "$fail" = 0 ]
