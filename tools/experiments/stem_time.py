"""Times the two ways into ArcFace's first layer: patches + 1x1 GEMM vs 8-channel image + stem-form conv."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from scrfd_arcface_facerecognition_b200 import ArcFace, _lib
from scrfd_arcface_facerecognition_b200.engine import stream_ptr
from tests.golden import inputs

n = 1024
rec = ArcFace("weights/w600k_r50.onnx")
eng, lib = rec._engine, rec._lib
frames = torch.from_numpy(np.stack([inputs.frame(80 + i, 1080, 1920) for i in range(8)])).cuda()
kps = torch.from_numpy(inputs.landmarks(81, 1080, 1920, n).reshape(n, 10)).cuda()
fidx = (torch.arange(n, device="cuda") % 8).to(torch.int32)
rec.embed_batch(frames, fidx, kps)
patches = eng.patch_buffer(n)
img8, launch8 = eng.stem8(n)
bound = eng._bound[n][0]
sp = stream_ptr()

def crop_patches():
    _lib.check(lib.b2f_norm_crop_patches(frames.data_ptr(), 1080, 1920, fidx.data_ptr(), kps.data_ptr(), n, 112, float(rec.input_mean),
                                         rec._scale, patches[0].data_ptr(), eng.dtype, sp))
def crop_img8():
    _lib.check(lib.b2f_norm_crop(frames.data_ptr(), 1080, 1920, fidx.data_ptr(), kps.data_ptr(), n, 112, float(rec.input_mean),
                                 rec._scale, img8.data_ptr(), 8, eng.dtype, None, None, sp))
def crop_img8_fast():
    _lib.check(lib.b2f_norm_crop_image8(frames.data_ptr(), 1080, 1920, fidx.data_ptr(), kps.data_ptr(), n, 112, float(rec.input_mean),
                                        rec._scale, img8.data_ptr(), eng.dtype, sp))
def gemm1x1():
    bound[1].fn(*bound[1].args, sp)
for name, fn in (("norm_crop_patches (32-ch patches)", crop_patches), ("1x1 GEMM over patches", gemm1x1),
                 ("norm_crop -> 8-ch image", crop_img8), ("norm_crop_image8 (CTA per face)", crop_img8_fast), ("stem-form conv", launch8)):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(200):
        fn()
    e1.record()
    torch.cuda.synchronize()
    print(f"{name:40s} {e0.elapsed_time(e1) * 5:8.1f} us")
