cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
B2F_AMODE=2 B2F_EPI=0 python tools/conv_bench.py 1024 56 56 64 64 3 1 2 0 1 > gpurun_out/exp4_plain.log 2>&1 && \
B2F_AMODE=2 B2F_EPI=0 ncu --set full --clock-control none --import-source on -k regex:conv_tile -s 2 -c 1 -f -o gpurun_out/prof_tile_dbg python tools/conv_bench.py 1024 56 56 64 64 3 1 2 0 1 5 > gpurun_out/exp4_ncu.log 2>&1
echo rc $?
cat gpurun_out/exp4_plain.log; tail -5 gpurun_out/exp4_ncu.log
