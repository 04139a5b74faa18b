cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_scale.py -k "match or duplicate or config3 or config4 or qdrant or l2norm" -m gpu -q -p no:cacheprovider 2>&1 | tail -3
for t in "10=0" "10=1"; do
B2F_TUNE=$t python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$t', d['value'], d['ms_per_step'], d['match'])"
done
