cd "${GRAFT_REPO_ROOT:-.}"
for cfg in "1024 112 112 64 64 3 1 2 0 1" "1024 56 56 64 64 3 1 0 1 0" "1024 28 28 128 128 3 1 2 0 1" "64 320 320 28 28 3 1 1 0 0" "64 160 160 56 56 3 1 1 1 0"; do
  for v in 0 1; do B2F_VHALO=$v python tools/conv_bench.py $cfg; done
done > gpurun_out/convbench.log 2>&1
python tools/conv_bench.py 1024 112 112 64 64 3 1 2 0 1 > gpurun_out/plain8.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:vhalo -s 3 -c 1 -o gpurun_out/prof_vhalo python tools/conv_bench.py 1024 112 112 64 64 3 1 2 0 1 10 0 3 > gpurun_out/ncu8.log 2>&1
cat gpurun_out/convbench.log
