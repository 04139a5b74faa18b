cd "${GRAFT_REPO_ROOT:-.}"
timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider 2>&1 | tail -5
for v in 1 0 1 0; do
B2F_STEM8=$v timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench_v6.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('stem8 $v', d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['sustained']['value'], d['top1_correct'], d['kernels_per_step'])"
done
tail -3 gpurun_out/bench_v6.err
