cd "${GRAFT_REPO_ROOT:-.}"
for S in "1024 112 112 27 64 1 1 2 0 0" "64 320 320 28 28 3 1 1 0 0" "1024 56 56 64 64 3 1 0 1 0" "64 80 80 56 88 1 1 0 0 0"; do
for d in 0 64; do B2F_DEBUG=$d timeout 60 python tools/conv_bench.py $S; done; done
