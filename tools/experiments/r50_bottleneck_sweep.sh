# where the time of the early ArcFace-R50 layers goes: whole layer (D0), loads only (D21), loads + epilogue (D4), loads + MMA (D17)
for shape in "1024 112 112 64 64 3 1 2" "1024 112 112 64 64 3 2 2" "1024 56 56 64 64 3 1 2" "1024 56 56 64 128 3 1 2" "1024 56 56 128 128 3 2 2" "1024 28 28 128 128 3 1 2"; do
  for dbg in 0 21 4 17; do B2F_DEBUG=$dbg python tools/conv_bench.py $shape 0 0 20; done
done
