"""Micro-benchmark of b2f_conv2d on one layer shape (device-resident, CUDA events).
usage: python tools/conv_bench.py N H W CIN COUT K STRIDE [act] [res] [bias9] [reps]
   tuning via env: B2F_PERSISTENT=0/1 B2F_VHALO=0/1 B2F_MAXN=..."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from scrfd_arcface_facerecognition_b200 import _lib

def main():
    a = [int(v) for v in sys.argv[1:]]
    n, h, w, cin, cout, k, stride = a[:7]
    act = a[7] if len(a) > 7 else 0
    res = a[8] if len(a) > 8 else 0
    bias9 = a[9] if len(a) > 9 else 0
    reps = a[10] if len(a) > 10 else 20
    lib = _lib.lib()
    for key, env in ((2, "B2F_PERSISTENT"), (3, "B2F_VHALO"), (1, "B2F_MAXN"), (4, "B2F_DEBUG"), (5, "B2F_GROUPS"), (6, "B2F_MT"), (7, "B2F_AMODE"), (8, "B2F_EPI")):
        if env in os.environ:
            _lib.check(lib.b2f_set_tuning(key, int(os.environ[env])))
    pad = k // 2 if k == 3 else 0
    ho, wo = (h + 2 * pad - k) // stride + 1, (w + 2 * pad - k) // stride + 1
    pc = lambda c: 16 if c <= 16 else (c + 31) // 32 * 32
    cin_p, cout_p = pc(cin), pc(cout)
    x = torch.randn((n, h, w, cin_p), device="cuda").half()
    wt = (torch.randn((k * k, cout_p, cin_p), device="cuda") * 0.05).half()
    bias = torch.randn((9 if bias9 else 1, cout_p), device="cuda")
    slope = torch.rand(cout_p, device="cuda")
    r = torch.randn((n, ho, wo, cout_p), device="cuda").half()
    pool = int(os.environ.get("B2F_POOL", "0"))          # fused 3x3 / s2 / p1 max-pool: `out` is the pooled map
    out = torch.empty((n, (ho - 1) // 2 + 1, (wo - 1) // 2 + 1, cout_p) if pool else (n, ho, wo, cout_p), device="cuda", dtype=torch.float16)
    d = _lib.ConvDesc()
    d.n, d.h, d.w, d.cin_p, d.ho, d.wo, d.cout_p = n, h, w, cin_p, ho, wo, cout_p
    d.kh, d.kw, d.stride, d.pad = k, k, stride, pad
    d.dtype, d.out_dtype, d.act, d.bias_classes = 0, 0, act, 9 if bias9 else 1
    d.in_, d.weight, d.bias, d.slope, d.out = x.data_ptr(), wt.data_ptr(), bias.data_ptr(), slope.data_ptr(), out.data_ptr()
    d.pool = pool
    if res:
        d.residual, d.res_mode = r.data_ptr(), 1
    sp = torch.cuda.current_stream().cuda_stream
    for _ in range(3):
        _lib.check(lib.b2f_conv2d(C.byref(d), sp))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        _lib.check(lib.b2f_conv2d(C.byref(d), sp))
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    fl = 2.0 * n * ho * wo * cout * cin * k * k
    byts = (x.numel() + out.numel() + (r.numel() if res else 0)) * 2
    print(f"conv n{n} {h}x{w} {cin}->{cout} k{k} s{stride} act{act} res{res} b9{bias9} "
          f"[P{os.environ.get('B2F_PERSISTENT','1')} V{os.environ.get('B2F_VHALO','1')} D{os.environ.get('B2F_DEBUG','0')}]: {ms*1e3:8.1f} us  "
          f"{fl/ms/1e9:7.1f} TFLOP/s  {byts/ms/1e6:7.1f} GB/s(act)")

main()
