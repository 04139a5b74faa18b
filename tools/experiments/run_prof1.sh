cd "${GRAFT_REPO_ROOT:-.}"
python tools/conv_bench.py 1024 112 112 64 64 3 1 2 0 1 10 3 > gpurun_out/plain9.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:vhalo -s 3 -c 1 -o gpurun_out/prof_vhalo2 python tools/conv_bench.py 1024 112 112 64 64 3 1 2 0 1 10 3 > gpurun_out/ncu9.log 2>&1
