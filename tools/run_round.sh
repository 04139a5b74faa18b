#!/bin/bash
# Round evidence on one B200: GPU tests, bench line, ncu launch list (kernel shares + DRAM bytes) and one
# `ncu --set full` capture of the dominant kernel.  Outputs land in gpurun_out/.
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/round_pytest.log 2>&1
echo "pytest rc $?"; tail -n 3 gpurun_out/round_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/round_smoke.log 2>&1
echo "smoke rc $?"; tail -n 2 gpurun_out/round_smoke.log
python bench.py > gpurun_out/round_bench.json 2> gpurun_out/round_bench.err
echo "bench rc $?"; cat gpurun_out/round_bench.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/round_bench_ref.json 2> gpurun_out/round_bench_ref.err
echo "ref rc $?"; cat gpurun_out/round_bench_ref.json
python bench.py --steps 2 --warmup 1 --no-graph --no-cpu-baseline > gpurun_out/round_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 260 -c 130 --csv --log-file gpurun_out/round_launches.csv python bench.py --steps 2 --warmup 1 --no-graph --no-cpu-baseline > gpurun_out/round_ncu1.log 2>&1
echo "ncu list rc $?"
ncu --set full --clock-control none --import-source on -k regex:conv_tile -s 40 -c 3 -f -o gpurun_out/round_prof_conv_tile python bench.py --steps 2 --warmup 1 --no-graph --no-cpu-baseline > gpurun_out/round_ncu2.log 2>&1
echo "ncu full rc $?"
