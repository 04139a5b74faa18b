cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
bash tools/gpu_check.sh conv models > gpurun_out/exp5_check.log 2>&1
cat gpurun_out/summary.txt
timeout 600 python -m pytest tests/test_gpu_scale.py -m gpu -q -p no:cacheprovider 2>&1 | tail -3
for i in 1 2; do
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --layer-report gpurun_out/layers_v3.csv > gpurun_out/bench_v3.json 2> gpurun_out/bench_v3.err
echo bench rc $?; python -c "
import json; d=json.load(open('gpurun_out/bench_v3.json')); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['clocks'])"
done
B2F_TUNE="2=1" python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('old-gen', d['value'], d['ms_per_step'], d['roofline']['frac'])"
