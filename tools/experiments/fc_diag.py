"""One case of the 7x7 'embedding' conv through the forced conv_tile kernel: python fc_diag.py N F32 CG2 [CIN K]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch, torch.nn.functional as F
from scrfd_arcface_facerecognition_b200 import _lib
from tests.test_gpu_kernels import run_conv, _q
n, f32, cg2 = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
cin = int(sys.argv[4]) if len(sys.argv) > 4 else 512
k = int(sys.argv[5]) if len(sys.argv) > 5 else 7
lib = _lib.lib()
g = torch.Generator().manual_seed(1)
x = _q(torch.randn((n, cin, k, k), generator=g))
w = _q(torch.randn((512, cin, k, k), generator=g) * (1.0 / (cin * k * k)) ** 0.5)
b = torch.randn(512, generator=g) * 0.1
ref = F.conv2d(x, w, b)
for key, val in ((2, 3), (11, cg2)):
    _lib.check(lib.b2f_set_tuning(key, val))
out = run_conv(lib, x, w, b, 1, 0, out_f32=bool(f32))
print(f"n {n} f32 {f32} cg2 {cg2} cin {cin} k {k}: max err {(out - ref).abs().max().item():.2e}", flush=True)
