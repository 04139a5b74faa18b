cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
bash tools/gpu_check.sh conv models > gpurun_out/exp5_check.log 2>&1
cat gpurun_out/summary.txt
timeout 900 python tools/conv_sweep.py rec,det old,auto,new,g2,g4,m2mt1,m2mt2,m1mt2,m0mt2 > gpurun_out/exp5_sweep.log 2>&1
cat gpurun_out/exp5_sweep.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --layer-report gpurun_out/layers_v3.csv > gpurun_out/bench_v3.json 2> gpurun_out/bench_v3.err
echo bench rc $?; cat gpurun_out/bench_v3.json; tail -3 gpurun_out/bench_v3.err
