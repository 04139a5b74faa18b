cd "${GRAFT_REPO_ROOT:-.}"
B2F_PLAN_TRACE=1 timeout 600 python tools/conv_sweep.py rec,sc,det auto 2>&1 | grep "b2f plan" | sort -u > gpurun_out/plans_auto.txt
timeout 600 python tools/conv_sweep.py rec,sc,det auto,g2,g4 2>&1 | tail -45
