cd "${GRAFT_REPO_ROOT:-.}"
python bench.py --steps 2 --warmup 1 --no-graph --no-cpu-baseline > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 260 -c 130 --csv --log-file gpurun_out/launches2.csv python bench.py --steps 2 --warmup 1 --no-graph --no-cpu-baseline > gpurun_out/ncu1.log 2>&1
echo ncu list rc $?
ncu --set full --clock-control none --import-source on -k regex:umma_conv_persistent -s 60 -c 2 -o gpurun_out/prof_persistent python bench.py --steps 2 --warmup 1 --no-graph --no-cpu-baseline > gpurun_out/ncu2.log 2>&1
echo ncu full rc $?
