cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
for dbg in 0 16 4 20 28; do
  echo "=== debug $dbg"
  B2F_DEBUG=$dbg timeout 300 python tools/conv_sweep.py rec old,direct,m1,m2,m2g2 2>&1 | head -8
done > gpurun_out/exp3_debug.log 2>&1
cat gpurun_out/exp3_debug.log
B2F_TRACE=1 timeout 300 python tools/conv_sweep.py rec 2>&1 | tail -40 > gpurun_out/exp3_trace.log
tail -30 gpurun_out/exp3_trace.log
