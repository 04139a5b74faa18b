cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -k "conv2d or conv_kernel_variants" -m gpu -q -x --tb=short -p no:cacheprovider > gpurun_out/exp2_tests.log 2>&1
echo "tests rc $?"
tail -n 5 gpurun_out/exp2_tests.log
timeout 900 python tools/conv_sweep.py rec,det old,auto,m0,m1,m2,m1mt2,m2mt2,m0mt2,g2,direct > gpurun_out/exp2_sweep.log 2>&1
echo "sweep rc $?"
cat gpurun_out/exp2_sweep.log
B2F_PLAN_TRACE=1 timeout 300 python tools/conv_sweep.py rec,det auto 2>&1 | grep "b2f plan" | sort -u > gpurun_out/exp2_plans.log
