"""Device-time of the memory-bound stages around the nets (CUDA events): stem-input kernels and pooling."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from scrfd_arcface_facerecognition_b200 import _lib
from tests.golden import inputs

lib = _lib.lib()
sp = lambda: torch.cuda.current_stream().cuda_stream


def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


B = 64
frames = torch.randint(0, 256, (B, 1080, 1920, 3), dtype=torch.uint8, device="cuda")
x = torch.empty((B, 640, 640, 4), dtype=torch.float16, device="cuda")
pat = torch.empty((B, 320, 320, 32), dtype=torch.float16, device="cuda")
t1 = timeit(lambda: _lib.check(lib.b2f_preprocess(frames.data_ptr(), B, 1080, 1920, 640, 360, 640, 640, 127.5, 1 / 128.0, x.data_ptr(), 4, 0, sp())))
t2 = timeit(lambda: _lib.check(lib.b2f_im2col3x3(x.data_ptr(), B, 640, 640, 2, 320, 320, 0, pat.data_ptr(), sp())))
t3 = timeit(lambda: _lib.check(lib.b2f_preprocess_patches(frames.data_ptr(), B, 1080, 1920, 640, 360, 640, 640, 2, 127.5, 1 / 128.0, pat.data_ptr(), 0, sp())))
print(f"det stem input: preprocess {t1:.1f} us + im2col {t2:.1f} us = {t1 + t2:.1f} us; fused {t3:.1f} us")

F = 1024
lm = inputs.landmarks(5, 1080, 1920, F)
kps = torch.from_numpy(lm.reshape(F, 10)).cuda()
fidx = (torch.arange(F, device="cuda") // 16).to(torch.int32)
xc = torch.empty((F, 112, 112, 4), dtype=torch.float16, device="cuda")
pc = torch.empty((F, 112, 112, 32), dtype=torch.float16, device="cuda")
sc = float(np.float32(1 / 127.5))
t1 = timeit(lambda: _lib.check(lib.b2f_norm_crop(frames.data_ptr(), 1080, 1920, fidx.data_ptr(), kps.data_ptr(), F, 112, 127.5, sc, xc.data_ptr(), 4, 0, None, None, sp())))
t2 = timeit(lambda: _lib.check(lib.b2f_im2col3x3(xc.data_ptr(), F, 112, 112, 1, 112, 112, 0, pc.data_ptr(), sp())))
t3 = timeit(lambda: _lib.check(lib.b2f_norm_crop_patches(frames.data_ptr(), 1080, 1920, fidx.data_ptr(), kps.data_ptr(), F, 112, 127.5, sc, pc.data_ptr(), 0, sp())))
print(f"rec stem input: norm_crop {t1:.1f} us + im2col {t2:.1f} us = {t1 + t2:.1f} us; fused {t3:.1f} us")

a = torch.randn((B, 320, 320, 64), device="cuda").half()
o = torch.empty((B, 160, 160, 64), dtype=torch.float16, device="cuda")
t = timeit(lambda: _lib.check(lib.b2f_pool(a.data_ptr(), B, 320, 320, 64, 3, 2, 1, 0, 160, 160, 0, o.data_ptr(), sp())))
print(f"maxpool 3x3 s2 64x320x320x64: {t:.1f} us  ({(a.numel() + o.numel()) * 2 / t / 1e3:.0f} GB/s)")
a2 = torch.randn((B, 160, 160, 64), device="cuda").half()
o2 = torch.empty((B, 80, 80, 64), dtype=torch.float16, device="cuda")
t = timeit(lambda: _lib.check(lib.b2f_pool(a2.data_ptr(), B, 160, 160, 64, 2, 2, 0, 1, 80, 80, 0, o2.data_ptr(), sp())))
print(f"avgpool 2x2 s2 64x160x160x64: {t:.1f} us  ({(a2.numel() + o2.numel()) * 2 / t / 1e3:.0f} GB/s)")
