cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
S="1024 14 14 256 256 3 1 2 0 1"
B2F_PERSISTENT=1 python tools/conv_bench.py $S > gpurun_out/exp8_plain.log 2>&1
B2F_PERSISTENT=3 B2F_AMODE=0 python tools/conv_bench.py $S >> gpurun_out/exp8_plain.log 2>&1
B2F_PLAN_TRACE=1 B2F_PERSISTENT=3 B2F_AMODE=0 python tools/conv_bench.py $S 2>&1 | grep "b2f plan" | sort -u >> gpurun_out/exp8_plain.log
B2F_PERSISTENT=1 ncu --set full --clock-control none --import-source on -k regex:umma_conv -s 2 -c 1 -f -o gpurun_out/prof_old_256 python tools/conv_bench.py $S 5 > gpurun_out/exp8_ncu.log 2>&1
B2F_PERSISTENT=3 B2F_AMODE=0 ncu --set full --clock-control none --import-source on -k regex:conv_tile -s 2 -c 1 -f -o gpurun_out/prof_new_256 python tools/conv_bench.py $S 5 >> gpurun_out/exp8_ncu.log 2>&1
cat gpurun_out/exp8_plain.log
