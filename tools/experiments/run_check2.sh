cd "${GRAFT_REPO_ROOT:-.}"
bash tools/gpu_check.sh conv models
for cfg in "1024 112 112 64 64 3 1 2 0 1" "1024 56 56 64 64 3 1 0 1 0" "1024 28 28 128 128 3 1 2 0 1" "64 320 320 28 28 3 1 1 0 0" "64 160 160 56 56 3 1 1 1 0" "1024 14 14 256 256 3 1 2 0 1" "1024 7 7 512 512 3 1 0 1 0" "1024 56 56 64 128 3 1 2 0 1" "1024 112 112 27 64 1 1 2 0 0"; do
  for v in 0 1; do B2F_VHALO=$v python tools/conv_bench.py $cfg; done
done > gpurun_out/convbench.log 2>&1
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --layer-report gpurun_out/layers.csv > gpurun_out/bench7.json 2> gpurun_out/bench7.err
echo bench rc $? >> gpurun_out/summary.txt
tail -3 gpurun_out/bench7.err
cat gpurun_out/convbench.log
