cd "${GRAFT_REPO_ROOT:-.}"
timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider 2>&1 | tail -4
for i in 1 2 3; do
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --layer-report gpurun_out/layers_v3.csv 2>gpurun_out/bench_v3.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['top1_correct'], d['kernels_per_step'])"
done
