for shape in "64 80 80 80 80" "64 80 80 88 88" "64 160 160 56 56" "64 20 20 224 224" "64 40 40 88 88" "64 80 80 56 80"; do
  for dbg in 0 21 4 17; do B2F_DEBUG=$dbg python tools/conv_bench.py $shape 3 1 1 0 0 30; done
done
