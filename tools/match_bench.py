"""Time the gallery match (coarse tcgen05 top-k + exact merge) for Q queries vs G rows."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from scrfd_arcface_facerecognition_b200.gallery import Gallery
q_n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
g_n = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
splits = int(sys.argv[3]) if len(sys.argv) > 3 else None
dev = torch.device("cuda", 0)
gen = torch.Generator(device=dev).manual_seed(2)
gal = Gallery()
gal.set_shard(torch.randn((g_n, 512), generator=gen, device=dev), 0)
q = torch.randn((q_n, 512), generator=gen, device=dev)
for _ in range(3):
    gal.match_local(q, 1, 0.4, True, splits)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    gal.match_local(q, 1, 0.4, True, splits)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f"match Q={q_n} G={g_n} splits={splits}: {ms:.3f} ms  {2.0*q_n*g_n*512/ms/1e9:.1f} TFLOP/s  gallery stream {g_n*1024/ms/1e6:.0f} GB/s x m-tiles {(q_n+127)//128}")
