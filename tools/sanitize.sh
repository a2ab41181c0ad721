#!/usr/bin/env bash
# compute-sanitizer passes over the CI-sized kernel parity tests (race / memory checking of the
# warp-specialised mbarrier pipelines).  Run on the GPU box:
#     gpurun --timeout 1500 -- 'bash tools/sanitize.sh'
# Logs land in gpurun_out/sanitizer_<tool>.log; a one-line-per-tool summary in gpurun_out/sanitizer_summary.txt
# (copied to profiles/ when judged).  The caching allocator is disabled so every tensor is its own cudaMalloc
# and an out-of-bounds access cannot hide inside a pooled block.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTORCH_NO_CUDA_MEMORY_CACHING=1
SAN=/usr/local/cuda/bin/compute-sanitizer
SMALL='not 128-256 and not 128-128 and not 256-256'
SUMMARY=gpurun_out/sanitizer_summary.txt
: > "$SUMMARY"

run() { # tool, per-tool budget [s], pytest -k expression, extra sanitizer flags...
  local tool=$1 budget=$2 expr=$3
  shift 3
  local log=gpurun_out/sanitizer_${tool}.log
  local t0=$SECONDS
  timeout "$budget" "$SAN" --tool "$tool" --target-processes all --error-exitcode 86 "$@" \
    python -m pytest tests/test_kernels_gpu.py -q -x -m gpu -k "$expr" -p no:cacheprovider > "$log" 2>&1
  local rc=$?
  local errs
  errs=$(grep -c "^========= \(Invalid\|Error\|Race\|Hazard\|Barrier error\|Program hit\|Uninitialized\)" "$log" || true)
  local tail_line
  tail_line=$(grep -E "passed|failed|error" "$log" | tail -1)
  local san_line
  san_line=$(grep -E "^========= (ERROR SUMMARY|RACECHECK SUMMARY)" "$log" | tail -1)
  echo "$tool rc=$rc secs=$((SECONDS - t0)) findings=$errs | pytest: ${tail_line:-none} | ${san_line:-no summary line}" >> "$SUMMARY"
}

run memcheck 420 "$SMALL" --leak-check no
run synccheck 180 "$SMALL and (gate or head or bn or xstitch)"
run racecheck 300 "$SMALL and (gate or head or bn or xstitch)" --racecheck-report all
cat "$SUMMARY"
