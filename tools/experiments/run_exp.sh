#!/bin/bash
# One parametrised launcher for the kernel experiments of rounds 1-2 (replaces the run_exp1..23.sh one-offs).
#   tools/experiments/run_exp.sh sweep [TUNE]        per-layer sweep of both nets       (B2F_TUNE="key=value,...")
#   tools/experiments/run_exp.sh layer N H W CIN COUT K STRIDE [act res bias9 reps]     one layer, CUDA events
#   tools/experiments/run_exp.sh bench [TUNE] [bench.py args]                           A/B of the whole step
#   tools/experiments/run_exp.sh ncu REGEX CMD...                                       plain run, then ncu --set full on REGEX
# Outputs go to gpurun_out/exp_<mode>_<timestamp>.log; meant to be run under gpurun from the repo root.
set -u
mode=${1:?mode}; shift
mkdir -p gpurun_out
out=gpurun_out/exp_${mode}_$(date +%H%M%S).log
case "$mode" in
  sweep) B2F_TUNE="${1:-}" python tools/conv_sweep.py > "$out" 2>&1 ;;
  layer) python tools/conv_bench.py "$@" > "$out" 2>&1 ;;
  bench) tune="${1:-}"; shift || true; B2F_TUNE="$tune" python bench.py --no-cpu-baseline "$@" > "$out" 2>&1 ;;
  ncu)   regex=${1:?kernel regex}; shift
         "$@" > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k "regex:$regex" -c 3 \
           -o "gpurun_out/exp_ncu_$(date +%H%M%S)" "$@" > "$out" 2>&1 ;;
  *) echo "unknown mode $mode" >&2; exit 2 ;;
esac
tail -5 "$out"
