cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider 2>&1 | tail -3
for t in "13=0" "13=1" "13=0" "13=1"; do
B2F_TUNE=$t timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$t', d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['top1_correct'])"
done
