#!/bin/bash
# Runs the GPU test groups as separate processes (a CUDA fault in one group must not poison the rest)
# and leaves one log per group under gpurun_out/.  Usage: tools/gpu_check.sh [group ...]
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
groups=("$@")
[ ${#groups[@]} -eq 0 ] && groups=(tma conv layers exact decode match models)
declare -A sel=(
  [tma]="tests/test_gpu_kernels.py -k tma_box"
  [conv]="tests/test_gpu_kernels.py -k 'conv2d or conv_kernel_variants'"
  [layers]="tests/test_gpu_kernels.py -k 'stem or depthwise or pool or im2col'"
  [exact]="tests/test_gpu_kernels.py -k 'letterbox or blob or warp'"
  [decode]="tests/test_gpu_kernels.py -k 'decode or forward_view'"
  [match]="tests/test_gpu_kernels.py -k 'l2norm or match or duplicate'"
  [models]="tests/test_gpu_models.py"
)
for g in "${groups[@]}"; do
  echo "=== $g" | tee -a gpurun_out/summary.txt
  eval timeout 600 python -m pytest ${sel[$g]} -m gpu -q -s --tb=short -p no:cacheprovider > gpurun_out/test_$g.log 2>&1
  echo "exit $?" | tee -a gpurun_out/summary.txt
  tail -n 3 gpurun_out/test_$g.log | tee -a gpurun_out/summary.txt
done
