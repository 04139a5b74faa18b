cd "${GRAFT_REPO_ROOT:-.}"
for S in "1024 14 14 256 256 3 1 2 0 1" "1024 14 14 256 512 3 1 2 0 1" "1024 28 28 128 256 3 1 2 0 1" "1024 7 7 512 512 3 1 2 0 1"; do
for t in "14=1" "14=2" "14=1" "14=2"; do B2F_TUNE=$t timeout 60 python tools/conv_bench.py $S 2>&1 | sed "s/^/$t /"; done; done
