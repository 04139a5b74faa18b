cd "${GRAFT_REPO_ROOT:-.}"
for args in "4 0 0" "4 1 0" "4 1 1" "130 1 0" "130 1 2" "260 1 2" "4 1 0 512 3" "4 1 0 64 7" "4 0 0 64 7"; do
B2F_PLAN_TRACE=1 timeout 120 python tools/experiments/fc_diag.py $args 2>&1 | grep -v "^\s*$" | tail -3 | cut -c1-260
done
