#!/bin/bash
# Round-end measurement pass on ONE B200 (run under gpurun from the repo root): tests, the four bench configs, the
# reference arm, an ncu launch list of the headline step and `ncu --set full` captures of the dominant kernels.
# Everything lands in gpurun_out/final/; tools/summarize_launches.py and tools/ncu_brief.py digest it for profiles/.
set -u
out=gpurun_out/final
mkdir -p $out
python -m pytest tests -m gpu -q 2>&1 | tail -3 > $out/pytest_gpu.txt
for c in 2 3 4 5; do
  python bench.py --config $c --layer-report $out/layers_cfg$c.csv > $out/bench_cfg${c}_1gpu.json 2> $out/bench_cfg$c.err
done
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $out/bench_cfg2_1gpu_steps20.json 2>> $out/bench_cfg2.err
python bench.py --impl reference --steps 5 --warmup 1 > $out/bench_cfg2_reference_arm.json 2> $out/bench_ref.err
# launch list of the headline step (eager, so every kernel is its own launch): durations + DRAM bytes
python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-graph > $out/plain_launches.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1200 --csv \
      --log-file $out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-graph > $out/ncu_launches.log 2>&1
# full captures of the dominant kernels on single-layer runs
python tools/conv_bench.py 1024 14 14 256 256 3 1 2 1 1 5 > $out/plain_c256.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:conv_tile -s 3 -c 1 -o $out/ncu_conv_256ch \
      python tools/conv_bench.py 1024 14 14 256 256 3 1 2 1 1 5 > $out/ncu_c256.log 2>&1
python tools/conv_bench.py 1024 112 112 64 64 3 1 2 0 1 5 > $out/plain_c64.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:conv_tile -s 3 -c 1 -o $out/ncu_conv_64ch_112 \
      python tools/conv_bench.py 1024 112 112 64 64 3 1 2 0 1 5 > $out/ncu_c64.log 2>&1
B2F_POOL=1 python tools/conv_bench.py 64 320 320 28 56 3 1 1 0 0 5 > $out/plain_pool.log 2>&1 &&
  B2F_POOL=1 ncu --set full --clock-control none --import-source on -k regex:conv_tile -s 3 -c 1 -o $out/ncu_conv_pool \
      python tools/conv_bench.py 64 320 320 28 56 3 1 1 0 0 5 > $out/ncu_pool.log 2>&1
python tools/match_one.py --q 1024 --g 1000000 > $out/plain_match.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:match_pair -s 1 -c 1 -o $out/ncu_match_pair_1024x1M \
      python tools/match_one.py --q 1024 --g 1000000 > $out/ncu_match.log 2>&1
ls -la $out | tail -30
