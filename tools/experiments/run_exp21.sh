cd "${GRAFT_REPO_ROOT:-.}"
timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider 2>&1 | tail -3
B2F_PLAN_TRACE=1 B2F_SWEEP_REPS=1500 timeout 300 python tools/conv_sweep.py fc auto,cg0 2>&1 | grep -v "^\s*$" | sort -u | cut -c1-250 | tail -6
for i in 1 2; do
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench_v4.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['sustained']['value'], d['top1_correct'], d['kernels_per_step'])"
done
