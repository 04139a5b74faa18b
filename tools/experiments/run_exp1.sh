# bottleneck isolation of the conv kernel: debug bits (1 no stores, 2 no residual, 4 no MMA, 8 no A loads, 16 no epilogue)
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -k "conv_kernel_variants" -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/exp1_variants.log 2>&1
echo "variants rc $?"
tail -n 15 gpurun_out/exp1_variants.log
{
for cfg in "1024 112 112 64 64 3 1 2 0 1" "1024 56 56 64 64 3 1 0 1 0" "1024 28 28 128 128 3 1 2 0 1" "64 160 160 56 56 3 1 1 1 0" "64 80 80 88 88 3 1 1 1 0" "1024 14 14 256 256 3 1 2 0 1"; do
  for v in 1 2; do
    for dbg in 0 1 3 4 8 16 19 12 28; do
      B2F_VHALO=$v B2F_DEBUG=$dbg timeout 120 python tools/conv_bench.py $cfg
    done
  done
done
} > gpurun_out/exp1_bench.log 2>&1
cat gpurun_out/exp1_bench.log
