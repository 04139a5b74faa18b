cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/round_pytest.log 2>&1
echo "pytest rc $?"; tail -n 6 gpurun_out/round_pytest.log
for i in 1 2; do
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --layer-report gpurun_out/layers_v3.csv > gpurun_out/bench_v3.json 2> gpurun_out/bench_v3.err
echo bench rc $?; python -c "
import json; d=json.load(open('gpurun_out/bench_v3.json')); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['kernels_per_step'], d['top1_correct'], d['clocks'])"
done
tail -3 gpurun_out/bench_v3.err
python bench.py --steps 2 --warmup 1 --no-graph --no-cpu-baseline > gpurun_out/round_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 250 -c 125 --csv --log-file gpurun_out/round_launches.csv python bench.py --steps 2 --warmup 1 --no-graph --no-cpu-baseline > gpurun_out/round_ncu1.log 2>&1
echo "ncu list rc $?"
