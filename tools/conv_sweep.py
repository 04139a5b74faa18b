"""Sweep b2f_conv2d over layer shapes x tuning variants in ONE process (device-resident, CUDA events).
usage: python tools/conv_sweep.py [shape-set] [variants]      shapes: N H W CIN COUT K STRIDE ACT RES BIAS9 [SC_CIN]
RES: 0 none, 1 separate residual tensor, 2 in place (residual == out).  B2F_SWEEP_REPS=1500 holds each case long enough
for the power cap to settle (the default 30 repetitions measure the first, un-capped milliseconds)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from scrfd_arcface_facerecognition_b200 import _lib

SHAPES = {
    "rec": [(1024, 112, 112, 27, 64, 1, 1, 2, 0, 0), (1024, 112, 112, 64, 64, 3, 1, 2, 0, 1), (1024, 112, 112, 64, 64, 3, 2, 0, 1, 0),
            (1024, 112, 112, 64, 64, 1, 2, 0, 0, 0), (1024, 56, 56, 64, 64, 3, 1, 2, 0, 1), (1024, 56, 56, 64, 64, 3, 1, 0, 1, 0),
            (1024, 56, 56, 64, 128, 3, 1, 2, 0, 1), (1024, 56, 56, 128, 128, 3, 2, 0, 1, 0), (1024, 28, 28, 128, 128, 3, 1, 2, 0, 1),
            (1024, 28, 28, 128, 128, 3, 1, 0, 1, 0), (1024, 28, 28, 128, 256, 3, 1, 2, 0, 1), (1024, 28, 28, 256, 256, 3, 2, 0, 1, 0),
            (1024, 14, 14, 256, 256, 3, 1, 2, 0, 1), (1024, 14, 14, 256, 256, 3, 1, 0, 1, 0), (1024, 14, 14, 256, 512, 3, 1, 2, 0, 1),
            (1024, 14, 14, 512, 512, 3, 2, 0, 1, 0), (1024, 7, 7, 512, 512, 3, 1, 2, 0, 1), (1024, 7, 7, 512, 512, 7, 1, 0, 0, 0)],
    "inplace": [(1024, 56, 56, 64, 64, 3, 1, 0, 1, 0), (1024, 56, 56, 64, 64, 3, 1, 0, 2, 0), (1024, 28, 28, 128, 128, 3, 1, 0, 1, 0),
                (1024, 28, 28, 128, 128, 3, 1, 0, 2, 0), (1024, 14, 14, 256, 256, 3, 1, 0, 1, 0), (1024, 14, 14, 256, 256, 3, 1, 0, 2, 0),
                (1024, 7, 7, 512, 512, 3, 1, 0, 1, 0), (1024, 7, 7, 512, 512, 3, 1, 0, 2, 0)],
    "fc": [(1024, 7, 7, 512, 512, 7, 1, 0, 0, 0), (64, 7, 7, 512, 512, 7, 1, 0, 0, 0), (1, 7, 7, 512, 512, 3, 1, 2, 0, 1)],
    "w14": [(1024, 14, 14, 256, 256, 3, 1, 2, 0, 1)],
    "w28": [(1024, 28, 28, 128, 128, 3, 1, 2, 0, 1), (1024, 28, 28, 128, 128, 3, 1, 0, 1, 0), (64, 28, 28, 128, 128, 3, 1, 0, 1, 0)],
    "r2": [(1024, 112, 112, 64, 64, 3, 1, 2, 0, 1), (1024, 56, 56, 64, 64, 3, 1, 2, 0, 1), (1024, 56, 56, 64, 64, 3, 1, 0, 1, 0)],
    "sc": [(1024, 112, 112, 64, 64, 3, 2, 0, 0, 0, 64), (1024, 56, 56, 128, 128, 3, 2, 0, 0, 0, 64),
           (1024, 28, 28, 256, 256, 3, 2, 0, 0, 0, 128), (1024, 14, 14, 512, 512, 3, 2, 0, 0, 0, 256)],
    "det": [(64, 320, 320, 27, 28, 1, 1, 1, 0, 0), (64, 320, 320, 28, 28, 3, 1, 1, 0, 0), (64, 320, 320, 28, 56, 3, 1, 1, 0, 0),
            (64, 160, 160, 56, 56, 3, 1, 1, 0, 0), (64, 160, 160, 56, 56, 3, 1, 1, 1, 0), (64, 160, 160, 56, 88, 3, 2, 1, 0, 0),
            (64, 80, 80, 88, 88, 3, 1, 1, 1, 0), (64, 80, 80, 56, 88, 1, 1, 0, 0, 0), (64, 80, 80, 88, 88, 3, 2, 1, 0, 0),
            (64, 40, 40, 88, 88, 3, 1, 1, 1, 0), (64, 40, 40, 88, 224, 3, 2, 1, 0, 0), (64, 20, 20, 224, 224, 3, 1, 1, 1, 0),
            (64, 80, 80, 56, 80, 3, 1, 1, 0, 0), (64, 80, 80, 80, 80, 3, 1, 1, 0, 0), (64, 40, 40, 80, 80, 3, 1, 1, 0, 0),
            (64, 20, 20, 80, 80, 3, 1, 1, 0, 0), (64, 80, 80, 56, 56, 3, 1, 0, 0, 0)],
}
# name -> {tuning key: value}; keys: 2 kernel generation, 3 halo on/off, 5 groups, 6 mt, 7 a_mode, 8 tma epilogue
VARIANTS = {
    "old": {2: 1}, "auto": {}, "new": {2: 3}, "m0": {2: 3, 7: 0}, "m1": {2: 3, 7: 1}, "m2": {2: 3, 7: 2}, "mt1": {2: 3, 6: 1},
    "mt2": {2: 3, 6: 2}, "g2": {2: 3, 5: 2}, "g4": {2: 3, 5: 4}, "direct": {2: 3, 8: 0}, "m2g2": {2: 3, 7: 2, 5: 2},
    "m2g4": {2: 3, 7: 2, 5: 4}, "m1mt2": {2: 3, 7: 1, 6: 2}, "m2mt2": {2: 3, 7: 2, 6: 2}, "m0mt2": {2: 3, 7: 0, 6: 2},
    "cg0": {2: 3, 11: 0}, "cg2": {2: 3, 11: 2}, "m0cg0": {2: 3, 7: 0, 11: 0}, "m0cg2": {2: 3, 7: 0, 11: 2},
    "br0": {15: 0}, "br1": {15: 1}, "m1mt1": {2: 3, 7: 1, 6: 1}, "m2mt1": {2: 3, 7: 2, 6: 1}, "m0mt1": {2: 3, 7: 0, 6: 1},
}
DEFAULTS = {2: 2, 3: 1, 4: 0, 5: 0, 6: 0, 7: -1, 8: -1, 11: 1, 15: 1}


def bench(lib, shape, reps=int(os.environ.get('B2F_SWEEP_REPS', '30'))):
    n, h, w, cin, cout, k, stride, act, res, bias9 = shape[:10]
    sc_cin = shape[10] if len(shape) > 10 else 0
    pad = k // 2 if k == 3 else 0
    ho, wo = (h + 2 * pad - k) // stride + 1, (w + 2 * pad - k) // stride + 1
    pc = lambda c: (c + 15) // 16 * 16
    cin_p, cout_p = pc(cin), pc(cout)
    x = torch.randn((n, h, w, cin_p), device="cuda").half()
    wt = (torch.randn((k * k, cout_p, cin_p), device="cuda") * 0.05).half()
    bias = torch.randn((9 if bias9 else 1, cout_p), device="cuda")
    slope = torch.rand(cout_p, device="cuda")
    r = torch.randn((n, ho, wo, cout_p), device="cuda").half()
    out = torch.empty((n, ho, wo, cout_p), device="cuda", dtype=torch.float16)
    d = _lib.ConvDesc()
    d.n, d.h, d.w, d.cin_p, d.ho, d.wo, d.cout_p = n, h, w, cin_p, ho, wo, cout_p
    d.kh, d.kw, d.stride, d.pad = k, k, stride, pad
    d.dtype, d.out_dtype, d.act, d.bias_classes = 0, 0, act, 9 if bias9 else 1
    d.in_, d.weight, d.bias, d.slope, d.out = x.data_ptr(), wt.data_ptr(), bias.data_ptr(), slope.data_ptr(), out.data_ptr()
    if res == 2:                                        # in-place block output: the output buffer holds the residual
        d.residual, d.res_mode = out.data_ptr(), 1
    elif res:
        d.residual, d.res_mode = r.data_ptr(), 1
    if sc_cin:                                          # fused projection shortcut: 1x1, same stride, from a second tensor
        sc_p = pc(sc_cin)
        xs = torch.randn((n, h, w, sc_p), device="cuda").half()
        ws = (torch.randn((1, cout_p, sc_p), device="cuda") * 0.05).half()
        d.sc_in, d.sc_weight, d.sc_cin_p, d.sc_stride, d.sc_h, d.sc_w = xs.data_ptr(), ws.data_ptr(), sc_p, stride, h, w
    sp = torch.cuda.current_stream().cuda_stream
    for _ in range(6):
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
    return ms, fl / ms / 1e9, byts / ms / 1e6


def main():
    sets = sys.argv[1].split(",") if len(sys.argv) > 1 else ["rec", "det"]
    names = sys.argv[2].split(",") if len(sys.argv) > 2 else list(VARIANTS)
    dbg = int(os.environ.get("B2F_DEBUG", "0"))
    lib = _lib.lib()
    print("shape".ljust(44) + "".join(v.rjust(9) for v in names) + "   (us; best TFLOP/s, GB/s)")
    for st in sets:
        for shape in SHAPES[st]:
            row, best = [], None
            for v in names:
                tune = dict(DEFAULTS)
                tune.update(VARIANTS[v])
                tune[4] = dbg
                for key, val in tune.items():
                    _lib.check(lib.b2f_set_tuning(key, val))
                try:
                    if os.environ.get("B2F_TRACE"):
                        print("  ..", shape, v, flush=True)
                    ms, tf, gb = bench(lib, shape)
                    row.append(f"{ms * 1e3:9.1f}")
                    if best is None or ms < best[0]:
                        best = (ms, tf, gb, v)
                except Exception as e:  # a variant that cannot be planned for this shape
                    row.append("      n/a")
                    torch.cuda.synchronize()
            n, h, w, cin, cout, k, s, act, res, b9 = shape[:10]
            tag = f"{st} n{n} {h}x{w} {cin}->{cout} k{k}s{s} a{act}r{res}b{b9}"
            print(tag.ljust(44) + "".join(row) + (f"   {best[3]} {best[1]:.0f} TF {best[2]:.0f} GB/s" if best else ""), flush=True)


main()
