#!/bin/bash
# Round evidence on one B200: GPU tests, bench line, ncu launch list (kernel shares + DRAM bytes) and `ncu --set full`
# captures of the dominant kernel.  Outputs land in gpurun_out/.
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
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 250 -c 125 --csv --log-file gpurun_out/round_launches.csv python bench.py --steps 2 --warmup 1 --no-graph --no-cpu-baseline > gpurun_out/round_ncu1.log 2>&1
echo "ncu list rc $?"
# full captures: a 256-channel ArcFace layer (CTA pair) and the 64-channel 112x112 layer (single CTA, full-halo boxes)
python tools/conv_bench.py 1024 14 14 256 256 3 1 2 0 1 > gpurun_out/round_cb.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:conv_tile -s 2 -c 1 -f -o gpurun_out/round_prof_pair_256 python tools/conv_bench.py 1024 14 14 256 256 3 1 2 0 1 5 > gpurun_out/round_ncu2.log 2>&1
echo "ncu full (pair) rc $?"
python tools/conv_bench.py 1024 112 112 64 64 3 1 2 0 1 >> gpurun_out/round_cb.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:conv_tile -s 2 -c 1 -f -o gpurun_out/round_prof_halo_64 python tools/conv_bench.py 1024 112 112 64 64 3 1 2 0 1 5 > gpurun_out/round_ncu3.log 2>&1
echo "ncu full (halo) rc $?"
cat gpurun_out/round_cb.log
