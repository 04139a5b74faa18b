"""One matching launch per kernel generation for ncu: python tools/match_one.py [--q 8192 --g 125000]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scrfd_arcface_facerecognition_b200 import _lib  # noqa: E402
from scrfd_arcface_facerecognition_b200.gallery import Gallery  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--q", type=int, default=8192)
ap.add_argument("--g", type=int, default=125_000)
ap.add_argument("--reps", type=int, default=3)
a = ap.parse_args()
gen = torch.Generator(device="cuda").manual_seed(1)
G = Gallery()
G.set_shard(torch.randn((a.g, 512), generator=gen, device="cuda"), 0)
qs = torch.randn((a.q, 512), generator=gen, device="cuda")
for g in (1, 0):
    _lib.call("b2f_set_tuning", 17, g)
    for _ in range(a.reps):
        G.match_local(qs, 1, 0.4, strict=True)
torch.cuda.synchronize()
print("ok")
