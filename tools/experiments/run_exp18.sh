cd "${GRAFT_REPO_ROOT:-.}"
for set in w14 w28 r2; do
nvidia-smi --query-gpu=clocks.sm,clocks.mem,power.draw,clocks_event_reasons.sw_power_cap --format=csv,noheader -lms 50 > gpurun_out/clk_$set.txt &
SMI=$!
B2F_SWEEP_REPS=6000 timeout 120 python tools/conv_sweep.py $set auto 2>&1 | tail -3
kill $SMI
sort gpurun_out/clk_$set.txt | uniq -c | sort -rn | head -8
done
