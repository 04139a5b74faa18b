cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider 2>&1 | tail -3
timeout 600 python tools/conv_sweep.py rec,det cg0,auto,cg2 > gpurun_out/exp11_sweep.log 2>&1
echo "sweep rc $?"; cat gpurun_out/exp11_sweep.log
for i in 1 2 3; do
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --layer-report gpurun_out/layers_v3.csv > gpurun_out/bench_v3.json 2> gpurun_out/bench_v3.err
echo bench rc $?; python -c "
import json; d=json.load(open('gpurun_out/bench_v3.json')); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['top1_correct'], d['clocks'])"
done
B2F_TUNE="11=0" python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('no pairs', d['value'], d['ms_per_step'], d['roofline']['frac'])"
