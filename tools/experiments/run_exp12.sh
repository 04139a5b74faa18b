cd "${GRAFT_REPO_ROOT:-.}"
timeout 300 python -m pytest tests/test_gpu_kernels.py -k "conv2d or conv_kernel_variants" -m gpu -q -x --tb=short -p no:cacheprovider 2>&1 | tail -4
S="1024 14 14 256 256 3 1 2 0 1"
for d in 0 16 4 28; do B2F_DEBUG=$d B2F_PERSISTENT=3 B2F_AMODE=0 timeout 60 python tools/conv_bench.py $S; done
B2F_TUNE="11=0" B2F_PERSISTENT=3 B2F_AMODE=0 timeout 60 python tools/conv_bench.py $S
