"""One batched detect + embed pass (64 x 1080p frames, 16 faces each) for ncu captures of the stage kernels:
python tools/stage_one.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("B2F_SYNTHETIC_WEIGHTS", "1")
from models import SCRFD, ArcFace  # noqa: E402
from scrfd_arcface_facerecognition_b200.pipeline import FacePipeline  # noqa: E402

det, rec = SCRFD("weights/det_10g.onnx"), ArcFace("weights/w600k_r50.onnx")
pipe = FacePipeline(det, rec, None, max_num=16)
frames = torch.from_numpy(np.random.default_rng(0).integers(0, 256, (64, 1080, 1920, 3), dtype=np.uint8)).cuda()
for _ in range(3):
    pipe.process(frames)
torch.cuda.synchronize()
print("ok")
