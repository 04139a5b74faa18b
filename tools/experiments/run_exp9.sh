cd "${GRAFT_REPO_ROOT:-.}"
S="1024 14 14 256 256 3 1 2 0 1"
for d in 0 1 16 4 8 12 28; do B2F_DEBUG=$d B2F_PERSISTENT=1 B2F_VHALO=0 python tools/conv_bench.py $S; done
for d in 0 1 16 4 8 12 28; do B2F_DEBUG=$d B2F_PERSISTENT=3 B2F_AMODE=0 python tools/conv_bench.py $S; done
for d in 0 16; do B2F_DEBUG=$d B2F_PERSISTENT=3 B2F_AMODE=0 B2F_EPI=1 python tools/conv_bench.py $S; done
