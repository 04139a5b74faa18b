cd "${GRAFT_REPO_ROOT:-.}"
for S in "1024 112 112 64 64 3 1 2 0 1" "1024 56 56 64 64 3 1 0 1 0" "64 160 160 56 56 3 1 1 0 0" "64 320 320 28 56 3 1 1 0 0"; do
for d in 0 128 16 144; do B2F_DEBUG=$d timeout 60 python tools/conv_bench.py $S; done; done
