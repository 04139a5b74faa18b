cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
python tools/conv_bench.py 1024 112 112 64 64 3 1 2 0 1 > gpurun_out/exp7_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:conv_tile -s 2 -c 1 -f -o gpurun_out/prof_tile_112 python tools/conv_bench.py 1024 112 112 64 64 3 1 2 0 1 5 > gpurun_out/exp7_ncu.log 2>&1
echo rc $?
python tools/conv_bench.py 1024 56 56 64 64 3 1 0 1 0 >> gpurun_out/exp7_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:conv_tile -s 2 -c 1 -f -o gpurun_out/prof_tile_56res python tools/conv_bench.py 1024 56 56 64 64 3 1 0 1 0 5 >> gpurun_out/exp7_ncu.log 2>&1
echo rc $?
cat gpurun_out/exp7_plain.log
