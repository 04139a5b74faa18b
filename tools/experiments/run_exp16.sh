cd "${GRAFT_REPO_ROOT:-.}"
for dbg in 0 1 16 17 8 4 25; do
echo "== debug $dbg"
B2F_DEBUG=$dbg timeout 200 python tools/conv_sweep.py r2 cg0,cg2,m0cg0,m0cg2,g2,g4 2>&1 | tail -4
done
