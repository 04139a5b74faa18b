"""GPU parity tests, kernel by kernel, all through the C-ABI (ctypes).

Integer / float32-exact kernels are compared bit-for-bit with the oracle (oracle/restate.py), real cv2
and the committed reference outputs; tensor-core kernels against a plain PyTorch fp32 reference with
the tolerance written next to each assert (fp16 operands, fp32 accumulation)."""
import ctypes as C

import cv2
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import restate
from scrfd_arcface_facerecognition_b200 import _lib
from tests.golden import inputs
from tests.golden.make_golden import postprocess_cases

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lib():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return _lib.lib()


def sp():
    return torch.cuda.current_stream().cuda_stream


def dev(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a))
    if dtype is not None:
        t = t.to(dtype)
    return t.cuda()


# =============================================================================================
# TMA conventions
# =============================================================================================

def _swizzle_expected(dense: np.ndarray, row_bytes: int) -> np.ndarray:
    """dense byte image with rows of row_bytes -> physical image under the 32/64/128-byte swizzle."""
    bits = {128: 3, 64: 2, 32: 1}[row_bytes]
    o = np.arange(dense.size)
    phys = o ^ (((o >> 7) & ((1 << bits) - 1)) << 4)
    out = np.zeros_like(dense)
    out[phys] = dense
    return out


@pytest.mark.parametrize("cfg", [
    # (C, W, H, N, box(c,w,h,n), estr, coords)
    (64, 20, 12, 3, (64, 8, 4, 2), (1, 1, 1, 1), (0, 3, 2, 1)),
    (64, 20, 12, 3, (64, 8, 4, 2), (1, 1, 1, 1), (0, -1, -1, 0)),       # negative coords: zero fill = conv padding
    (128, 20, 12, 3, (64, 16, 8, 1), (1, 1, 1, 1), (64, 10, 8, 2)),     # runs off the right/bottom edge
    (64, 21, 13, 2, (64, 16, 8, 1), (1, 2, 2, 1), (0, -1, -1, 1)),      # elementStrides 2 == stride-2 convolution
    (32, 20, 12, 3, (32, 8, 4, 4), (1, 1, 1, 1), (0, 2, 1, 0)),         # 64-byte rows / SWIZZLE_64B, n overflow
    (16, 20, 12, 3, (16, 16, 8, 1), (1, 1, 1, 1), (0, -1, 0, 2)),       # 32-byte rows / SWIZZLE_32B
], ids=["plain", "negative", "edge", "stride2", "sw64", "sw32"])
def test_tma_box_layout(lib, cfg):
    c, w, h, n, box, es, coords = cfg
    rng = np.random.default_rng(0)
    src = rng.integers(1, 60000, (n, h, w, c)).astype(np.uint16)       # never zero, so fill is detectable
    cnt = [(box[i] + es[i] - 1) // es[i] for i in range(4)]
    nbytes = 2 * cnt[0] * cnt[1] * cnt[2] * cnt[3]
    d_src = dev(src.view(np.int16))
    out = torch.zeros(nbytes, dtype=torch.uint8, device="cuda")
    dims = (C.c_longlong * 4)(c, w, h, n)
    bx, est, crd = (C.c_int * 4)(*box), (C.c_int * 4)(*es), (C.c_int * 4)(*coords)
    _lib.check(lib.b2f_debug_tma_probe(d_src.data_ptr(), dims, bx, est, box[0] * 2, crd, out.data_ptr(), nbytes, sp()))
    torch.cuda.synchronize()
    got = out.cpu().numpy()
    dense = np.zeros((cnt[3], cnt[2], cnt[1], cnt[0]), np.uint16)
    for a in range(cnt[3]):
        for b in range(cnt[2]):
            for cc in range(cnt[1]):
                nn, yy, xx = coords[3] + a * es[3], coords[2] + b * es[2], coords[1] + cc * es[1]
                if 0 <= nn < n and 0 <= yy < h and 0 <= xx < w:
                    dense[a, b, cc] = src[nn, yy, xx, coords[0]:coords[0] + cnt[0]]
    exp = _swizzle_expected(dense.reshape(-1).view(np.uint8), box[0] * 2)
    bad = np.nonzero(got != exp)[0]
    assert bad.size == 0, f"{bad.size} of {nbytes} bytes differ; first at {bad[:8]}"


# =============================================================================================
# tcgen05 convolution
# =============================================================================================

def _pad16(c):
    return (c + 15) // 16 * 16


def run_conv(lib, x_nchw, w_oihw, bias, stride, pad, act=0, slope=None, residual=None, res_mode=0,
             out_f32=False, bias_tab=None, dtype=0, force_kchunk=0, in_place=False, splitk=False):
    """x (N,C,H,W) f32 torch cpu; returns (N,Cout,Ho,Wo) f32 cpu computed by b2f_conv2d."""
    tdt = torch.bfloat16 if dtype == 1 else torch.float16
    n, cin, h, w = x_nchw.shape
    cout, _, kh, kw = w_oihw.shape
    cin_p, cout_p = _pad16(cin), _pad16(cout)
    ho, wo = (h + 2 * pad - kh) // stride + 1, (w + 2 * pad - kw) // stride + 1
    xin = torch.zeros((n, h, w, cin_p), dtype=tdt)
    xin[..., :cin] = x_nchw.permute(0, 2, 3, 1).to(tdt)
    wk = torch.zeros((kh * kw, cout_p, cin_p), dtype=tdt)
    wk[:, :cout, :cin] = w_oihw.permute(2, 3, 0, 1).reshape(kh * kw, cout, cin).to(tdt)
    classes = 1 if bias_tab is None else 9
    bt = torch.zeros((classes, cout_p), dtype=torch.float32)
    if bias_tab is None:
        bt[0, :cout] = bias
    else:
        bt[:, :cout] = bias_tab
    d = _lib.ConvDesc()
    d.n, d.h, d.w, d.cin_p, d.ho, d.wo, d.cout_p = n, h, w, cin_p, ho, wo, cout_p
    d.kh, d.kw, d.stride, d.pad = kh, kw, stride, pad
    d.dtype, d.out_dtype, d.act, d.bias_classes = dtype, 2 if out_f32 else dtype, act, classes
    d.force_kchunk = force_kchunk
    keep = [xin.cuda(), wk.cuda(), bt.cuda()]
    d.in_, d.weight, d.bias = (t.data_ptr() for t in keep)
    if slope is not None:
        sl = torch.zeros(cout_p)
        sl[:cout] = slope
        keep.append(sl.cuda())
        d.slope = keep[-1].data_ptr()
    if residual is not None:
        rn, rc, rh, rw = residual.shape
        r = torch.zeros((rn, rh, rw, cout_p), dtype=tdt)
        r[..., :cout] = residual.permute(0, 2, 3, 1).to(tdt)
        keep.append(r.cuda())
        d.residual, d.res_mode, d.res_h, d.res_w = keep[-1].data_ptr(), res_mode, rh, rw
    out = torch.full((n, ho, wo, cout_p), float("nan"), dtype=torch.float32 if out_f32 else tdt, device="cuda")
    if in_place:                                  # the output buffer holds the residual and is overwritten by the sum
        out = keep[-1]
        d.residual = out.data_ptr()
    d.out = out.data_ptr()
    before = lib.b2f_launch_count()
    if splitk:                                    # the caller's workspace lets long-K fp32 layers split along K
        keep.append(torch.full((8 * out.numel(),), float("nan"), dtype=torch.float32, device="cuda"))
        d.splitk_ws, d.splitk_ws_bytes = keep[-1].data_ptr(), keep[-1].numel() * 4
    _lib.check(lib.b2f_conv2d(C.byref(d), sp()), "b2f_conv2d")
    torch.cuda.synchronize()
    run_conv.launches = lib.b2f_launch_count() - before
    res = out.float().cpu()
    assert torch.isfinite(res).all(), "conv left unwritten / non-finite outputs"
    assert (res[..., cout:] == (0.5 if act == 3 else 0.0)).all(), "padding channels must stay neutral"
    return res[..., :cout].permute(0, 3, 1, 2).contiguous()


def _q(t, dtype=0):
    return t.to(torch.bfloat16 if dtype == 1 else torch.float16).float()


CONV_CASES = [
    # n, h, w, cin, cout, k, stride, pad
    (2, 16, 16, 64, 64, 1, 1, 0),
    (2, 16, 16, 64, 64, 3, 1, 1),
    (1, 112, 112, 64, 64, 3, 1, 1),
    (3, 56, 56, 64, 128, 3, 2, 1),
    (3, 56, 56, 64, 128, 1, 2, 0),
    (20, 14, 14, 256, 256, 3, 1, 1),
    (40, 7, 7, 512, 512, 3, 1, 1),
    (5, 28, 28, 128, 256, 3, 2, 1),
    (2, 40, 40, 96, 32, 3, 1, 1),        # kchunk 32 (SWIZZLE_64B)
    (2, 20, 20, 80, 80, 3, 1, 1),        # kchunk 16 (SWIZZLE_32B)
    (2, 33, 47, 32, 48, 3, 1, 1),        # ragged spatial size
    (2, 33, 47, 32, 48, 3, 2, 1),
    (2, 80, 80, 56, 24, 3, 1, 1),        # channels that need padding (56 -> 64, 24 -> 32)
    (130, 7, 7, 64, 32, 7, 1, 0),        # fully-connected as a 7x7 valid conv; M tile spans images
    (300, 1, 1, 512, 512, 1, 1, 0),      # plain GEMM, two N tiles
]


@pytest.mark.parametrize("case", CONV_CASES, ids=lambda c: "x".join(map(str, c)))
def test_conv2d_matches_fp32_reference(lib, case):
    n, h, w, cin, cout, k, stride, pad = case
    g = torch.Generator().manual_seed(sum(case))
    x = _q(torch.randn((n, cin, h, w), generator=g))
    wt = _q(torch.randn((cout, cin, k, k), generator=g) * (2.0 / (cin * k * k)) ** 0.5)
    b = torch.randn(cout, generator=g) * 0.1
    ref = F.conv2d(x, wt, b, stride, pad)
    out = run_conv(lib, x, wt, b, stride, pad, out_f32=True)
    err = (out - ref).abs().max().item()
    # identical fp16 operands, fp32 accumulation on both sides: only summation order differs
    assert err <= 2e-3 * max(1.0, ref.abs().max().item()), f"max abs err {err}"
    out16 = run_conv(lib, x, wt, b, stride, pad, act=1)
    assert (out16 - _q(torch.relu(ref))).abs().max().item() <= 4e-3 * max(1.0, ref.abs().max().item())


def test_conv2d_epilogue_variants(lib):
    g = torch.Generator().manual_seed(7)
    x = _q(torch.randn((3, 64, 28, 28), generator=g))
    wt = _q(torch.randn((96, 64, 3, 3), generator=g) * 0.06)
    b = torch.randn(96, generator=g) * 0.1
    ref = F.conv2d(x, wt, b, 1, 1)
    scale = max(1.0, ref.abs().max().item())
    # PReLU + same-size residual (add before the activation)
    res = _q(torch.randn((3, 96, 28, 28), generator=g))
    slope = torch.rand(96, generator=g) * 0.3 + 0.1
    y = ref + res
    want = torch.where(y >= 0, y, y * slope.view(1, -1, 1, 1))
    got = run_conv(lib, x, wt, b, 1, 1, act=2, slope=slope, residual=res, res_mode=1, out_f32=True)
    assert (got - want).abs().max().item() <= 2e-3 * scale
    # nearest-2x upsampled residual (PAFPN top-down path)
    res2 = _q(torch.randn((3, 96, 14, 14), generator=g))
    want = ref + res2.repeat_interleave(2, 2).repeat_interleave(2, 3)
    got = run_conv(lib, x, wt, b, 1, 1, residual=res2, res_mode=2, out_f32=True)
    assert (got - want).abs().max().item() <= 2e-3 * scale
    # sigmoid head, fp32 out
    got = run_conv(lib, x, wt, b, 1, 1, act=3, out_f32=True)
    assert (got - torch.sigmoid(ref)).abs().max().item() <= 1e-3
    # nine-class border bias table (a BatchNorm shift folded through zero padding)
    tab = torch.randn((9, 96), generator=g)
    iy = torch.arange(28)
    cls = torch.where(iy - 1 < 0, 0, torch.where(iy + 1 >= 28, 2, 1))
    cmap = cls[:, None] * 3 + cls[None, :]
    want = F.conv2d(x, wt, None, 1, 1) + tab[cmap].permute(2, 0, 1).unsqueeze(0)
    got = run_conv(lib, x, wt, None, 1, 1, bias_tab=tab, out_f32=True)
    assert (got - want).abs().max().item() <= 2e-3 * scale
    # bf16 operands
    xb, wb = _q(x, 1), _q(wt, 1)
    got = run_conv(lib, xb, wb, b, 1, 1, out_f32=True, dtype=1)
    assert (got - F.conv2d(xb, wb, b, 1, 1)).abs().max().item() <= 2e-3 * scale
    # forcing a smaller K chunk must not change the result
    got = run_conv(lib, x, wt, b, 1, 1, out_f32=True, force_kchunk=16)
    assert (got - ref).abs().max().item() <= 2e-3 * scale


@pytest.mark.parametrize("case", [(3, 56, 56, 64, 64, 64, 2), (3, 28, 28, 128, 128, 64, 2), (5, 14, 14, 256, 256, 128, 2),
                                  (9, 14, 14, 256, 512, 256, 2), (2, 33, 47, 64, 96, 64, 2), (2, 20, 20, 64, 64, 128, 1), (2, 24, 24, 32, 48, 96, 2)],
                         ids=lambda c: "x".join(map(str, c)))
def test_conv2d_fused_projection_shortcut(lib, case):
    """out = conv3x3_stride_s(y) + conv1x1_stride_s(x): the shortcut rides as extra K chunks of the main convolution
    (ResNet down-sampling blocks, reference models/arcface.py:51 graph); single CTAs and CTA pairs"""
    n, h, w, cin, cout, sc_cin, stride = case
    g = torch.Generator().manual_seed(sum(case))
    y = _q(torch.randn((n, cin, h, w), generator=g))
    x = _q(torch.randn((n, sc_cin, h, w), generator=g))
    wt = _q(torch.randn((cout, cin, 3, 3), generator=g) * (2.0 / (cin * 9)) ** 0.5)
    ws = _q(torch.randn((cout, sc_cin, 1, 1), generator=g) * (2.0 / sc_cin) ** 0.5)
    b = torch.randn(cout, generator=g) * 0.1
    ref = F.conv2d(y, wt, b, stride, 1) + F.conv2d(x, ws, None, stride, 0)
    ho, wo = ref.shape[2], ref.shape[3]
    cin_p, cout_p, sc_p = _pad16(cin), _pad16(cout), _pad16(sc_cin)
    keep = {}
    keep["y"] = torch.zeros((n, h, w, cin_p), dtype=torch.float16); keep["y"][..., :cin] = y.permute(0, 2, 3, 1).half()
    keep["x"] = torch.zeros((n, h, w, sc_p), dtype=torch.float16); keep["x"][..., :sc_cin] = x.permute(0, 2, 3, 1).half()
    keep["w"] = torch.zeros((9, cout_p, cin_p), dtype=torch.float16)
    keep["w"][:, :cout, :cin] = wt.permute(2, 3, 0, 1).reshape(9, cout, cin).half()
    keep["ws"] = torch.zeros((1, cout_p, sc_p), dtype=torch.float16); keep["ws"][0, :cout, :sc_cin] = ws[:, :, 0, 0].half()
    keep["b"] = torch.zeros((1, cout_p)); keep["b"][0, :cout] = b
    dev_t = {k: v.cuda() for k, v in keep.items()}
    try:
        for pairs in (0, 1, 2):
            _lib.check(lib.b2f_set_tuning(11, pairs))
            for f32 in (True, False):
                out = torch.full((n, ho, wo, cout_p), float("nan"), dtype=torch.float32 if f32 else torch.float16, device="cuda")
                d = _lib.ConvDesc()
                d.n, d.h, d.w, d.cin_p, d.ho, d.wo, d.cout_p = n, h, w, cin_p, ho, wo, cout_p
                d.kh, d.kw, d.stride, d.pad = 3, 3, stride, 1
                d.dtype, d.out_dtype, d.act, d.bias_classes = 0, 2 if f32 else 0, 0, 1
                d.in_, d.weight, d.bias, d.out = dev_t["y"].data_ptr(), dev_t["w"].data_ptr(), dev_t["b"].data_ptr(), out.data_ptr()
                d.sc_in, d.sc_weight = dev_t["x"].data_ptr(), dev_t["ws"].data_ptr()
                d.sc_cin_p, d.sc_stride, d.sc_h, d.sc_w = sc_p, stride, h, w
                _lib.check(lib.b2f_conv2d(C.byref(d), sp()), "b2f_conv2d")
                torch.cuda.synchronize()
                got = out.float().cpu()[..., :cout].permute(0, 3, 1, 2)
                err = (got - ref).abs().max().item()
                tol = (2e-3 if f32 else 6e-3) * max(1.0, ref.abs().max().item())
                assert err <= tol, f"pairs={pairs} f32={f32}: {err}"
    finally:
        _lib.check(lib.b2f_set_tuning(11, 1))


@pytest.mark.parametrize("case", [(2, 64, 64, 32, 64), (3, 320, 320, 32, 64), (2, 33, 47, 32, 32), (2, 50, 24, 64, 96), (1, 16, 8, 32, 128)],
                         ids=lambda c: "x".join(map(str, c)))
def test_conv2d_fused_maxpool(lib, case):
    """3x3 ReLU convolution with the 3x3 / s2 / p1 max-pool fused into its epilogue (SCRFD stem, b2f.h `pool`): BIT-EXACT
    against the same convolution followed by b2f_pool -- max is order-independent, so the TMA max-reduce of the tiles'
    partial window maxima equals the separate pass, also for maps that tiles overhang (odd sizes) and for bf16"""
    n, h, w, cin, cout = case
    g = torch.Generator().manual_seed(sum(case))
    for dtype, tdt in ((0, torch.float16), (1, torch.bfloat16)):
        x = (torch.randn((n, h, w, cin), generator=g)).to(tdt).cuda()
        wt = (torch.randn((9, cout, cin), generator=g) * (2.0 / (cin * 9)) ** 0.5).to(tdt).cuda()
        bias = (torch.randn((1, cout), generator=g) * 0.1).cuda()
        hp, wp = (h - 1) // 2 + 1, (w - 1) // 2 + 1
        full = torch.empty((n, h, w, cout), dtype=tdt, device="cuda")
        pooled = torch.full((n, hp, wp, cout), float("nan"), dtype=tdt, device="cuda")      # the entry must zero it itself
        want = torch.empty((n, hp, wp, cout), dtype=tdt, device="cuda")

        def desc(out, pool):
            d = _lib.ConvDesc()
            d.n, d.h, d.w, d.cin_p, d.ho, d.wo, d.cout_p = n, h, w, cin, h, w, cout
            d.kh, d.kw, d.stride, d.pad = 3, 3, 1, 1
            d.dtype, d.out_dtype, d.act, d.bias_classes = dtype, dtype, 1, 1
            d.in_, d.weight, d.bias, d.out, d.pool = x.data_ptr(), wt.data_ptr(), bias.data_ptr(), out.data_ptr(), pool
            return d
        _lib.check(lib.b2f_conv2d(C.byref(desc(full, 0)), sp()), "b2f_conv2d")
        _lib.check(lib.b2f_pool(full.data_ptr(), n, h, w, cout, 3, 2, 1, 0, hp, wp, dtype, want.data_ptr(), sp()), "b2f_pool")
        for _ in range(2):                                       # twice: the second launch finds stale maxima in `out`
            _lib.check(lib.b2f_conv2d(C.byref(desc(pooled, 1)), sp()), "b2f_conv2d")
        torch.cuda.synchronize()
        assert torch.equal(pooled.view(torch.int16), want.view(torch.int16)), f"dtype {dtype}"
        ref = F.max_pool2d(torch.relu(F.conv2d(x.float().permute(0, 3, 1, 2), wt.float().permute(1, 2, 0).reshape(cout, cin, 3, 3),
                                               bias[0], 1, 1)), 3, 2, 1).permute(0, 2, 3, 1)
        assert (pooled.float() - ref).abs().max().item() <= (6e-3 if dtype == 0 else 4e-2) * max(1.0, ref.abs().max().item())


@pytest.mark.parametrize("case", [(2, 64, 64, 64, 64), (2, 48, 80, 96, 96), (1, 32, 32, 128, 128), (3, 24, 40, 32, 32),
                                  (2, 56, 56, 64, 128), (2, 20, 20, 224, 224), (1, 33, 47, 64, 32), (2, 40, 24, 80, 80)],
                         ids=lambda c: "x".join(map(str, c)))
def test_conv_kernel_variants_agree(lib, case):
    """per-tile kernel, persistent kernel, vertical-halo and full-halo kernels (resident / streamed weights) on 3x3 s1 convs"""
    n, h, w, cin, cout = case
    g = torch.Generator().manual_seed(sum(case))
    x = _q(torch.randn((n, cin, h, w), generator=g))
    wt = _q(torch.randn((cout, cin, 3, 3), generator=g) * (2.0 / (cin * 9)) ** 0.5)
    b = torch.randn(cout, generator=g) * 0.1
    res = _q(torch.randn((n, cout, h, w), generator=g))
    ref = torch.relu(F.conv2d(x, wt, b, 1, 1) + res)
    # (generation, vhalo, a_mode, mt, groups, tma_epilogue): 0 = CTA per tile, 1 = first persistent kernels,
    # 2 = default dispatch, 3 = conv_tile_kernel for every layer; last field: CTA pairs 0 never / 1 wide tiles / 2 forced
    variants = [(0, 0, -1, 0, 0, -1, 1), (1, 0, -1, 0, 0, -1, 1), (1, 1, -1, 0, 0, -1, 1), (1, 2, -1, 0, 0, -1, 1),
                (2, 1, -1, 0, 0, -1, 1), (3, 1, -1, 0, 0, -1, 1), (3, 1, -1, 0, 0, -1, 0)]
    variants += [(3, 1, am, mt, gr, 1, 0) for am in (0, 1, 2) for mt in (1, 2) for gr in (2, 4)]
    variants += [(3, 1, 2, 1, 4, 0, 0), (3, 1, 0, 2, 2, 0, 0)]
    variants += [(3, 1, 0, 0, 0, -1, 2), (3, 1, 0, 0, 2, 1, 2), (3, 1, 0, 0, 4, 0, 2)]      # CTA pairs (cta_group::2)
    variants += [(3, 1, am, mt, 2, 1, 2) for am in (1, 2) for mt in (1, 2)] + [(3, 1, -1, 0, 0, -1, 2), (3, 1, 2, 1, 4, 0, 2)]
    try:
        for persistent, vhalo, amode, mt, groups, epi, cg2 in variants:
            for key, val in ((2, persistent), (3, vhalo), (7, amode), (6, mt), (5, groups), (8, epi), (11, cg2)):
                _lib.check(lib.b2f_set_tuning(key, val))
            for f32 in (True, False):
                out = run_conv(lib, x, wt, b, 1, 1, act=1, residual=res, res_mode=1, out_f32=f32)
                err = (out - ref).abs().max().item()
                tol = (2e-3 if f32 else 4e-3) * max(1.0, ref.abs().max().item())
                assert err <= tol, f"variant {(persistent, vhalo, amode, mt, groups, epi, cg2)} f32={f32}: {err}"
    finally:
        for key, val in ((2, 2), (3, 1), (7, -1), (6, 0), (5, 0), (8, -1), (11, 1)):
            _lib.check(lib.b2f_set_tuning(key, val))


# =============================================================================================
# CUDA-core layers
# =============================================================================================

def _nhwc16(x, cp):
    n, c, h, w = x.shape
    t = torch.zeros((n, h, w, cp), dtype=torch.float16)
    t[..., :c] = x.permute(0, 2, 3, 1).half()
    return t.cuda()


@pytest.mark.parametrize("stride,cout", [(1, 64), (2, 28), (2, 128)])
def test_stem_conv(lib, stride, cout):
    g = torch.Generator().manual_seed(stride + cout)
    x = _q(torch.rand((2, 3, 37, 45), generator=g) * 2 - 1)
    wt = torch.randn((cout, 3, 3, 3), generator=g) * 0.3
    b = torch.randn(cout, generator=g) * 0.1
    slope = torch.rand(cout, generator=g) * 0.3
    cp = _pad16(cout)
    wk = torch.zeros((9, 4, cp))
    wk[:, :3, :cout] = wt.permute(2, 3, 1, 0).reshape(9, 3, cout)
    bk, sk = torch.zeros(cp), torch.zeros(cp)
    bk[:cout], sk[:cout] = b, slope
    xin = _nhwc16(x, 4)
    ho, wo = (37 + 2 - 3) // stride + 1, (45 + 2 - 3) // stride + 1
    out = torch.empty((2, ho, wo, cp), dtype=torch.float16, device="cuda")
    keep = [wk.cuda(), bk.cuda(), sk.cuda()]
    _lib.check(lib.b2f_stem_conv3x3(xin.data_ptr(), 2, 37, 45, 4, stride, keep[0].data_ptr(), keep[1].data_ptr(),
                                    keep[2].data_ptr(), 2, cp, 0, out.data_ptr(), sp()))
    y = F.conv2d(x, wt, b, stride, 1)
    want = torch.where(y >= 0, y, y * slope.view(1, -1, 1, 1))
    got = out.float().cpu()[..., :cout].permute(0, 3, 1, 2)
    assert (got - want).abs().max().item() <= 3e-3 * max(1.0, want.abs().max().item())   # fp16 output rounding


@pytest.mark.parametrize("k,stride,pad", [(3, 1, 1), (3, 2, 1), (7, 1, 0)])
def test_depthwise_conv(lib, k, stride, pad):
    g = torch.Generator().manual_seed(k * 10 + stride)
    c, h, w = 40, (7 if k == 7 else 19), (7 if k == 7 else 23)
    x = _q(torch.randn((3, c, h, w), generator=g))
    wt = torch.randn((c, 1, k, k), generator=g) * 0.3
    b = torch.randn(c, generator=g) * 0.1
    cp = _pad16(c)
    wk, bk = torch.zeros((k * k, cp)), torch.zeros(cp)
    wk[:, :c] = wt.reshape(c, k * k).t()
    bk[:c] = b
    ho, wo = (h + 2 * pad - k) // stride + 1, (w + 2 * pad - k) // stride + 1
    out = torch.empty((3, ho, wo, cp), dtype=torch.float16, device="cuda")
    keep = [_nhwc16(x, cp), wk.cuda(), bk.cuda()]
    _lib.check(lib.b2f_dwconv(keep[0].data_ptr(), 3, h, w, cp, k, stride, pad, keep[1].data_ptr(), keep[2].data_ptr(),
                              None, 1, 0, out.data_ptr(), sp()))
    want = torch.relu(F.conv2d(x, wt, b, stride, pad, groups=c))
    got = out.float().cpu()[..., :c].permute(0, 3, 1, 2)
    assert (got - want).abs().max().item() <= 3e-3 * max(1.0, want.abs().max().item())


def test_im2col3x3(lib):
    g = torch.Generator().manual_seed(11)
    for stride in (1, 2):
        x = _q(torch.rand((2, 3, 21, 30), generator=g) * 2 - 1)
        xin = _nhwc16(x, 4)
        ho, wo = (21 + 2 - 3) // stride + 1, (30 + 2 - 3) // stride + 1
        out = torch.empty((2, ho, wo, 32), dtype=torch.float16, device="cuda")
        _lib.check(lib.b2f_im2col3x3(xin.data_ptr(), 2, 21, 30, stride, ho, wo, 0, out.data_ptr(), sp()))
        cols = F.unfold(x, 3, padding=1, stride=stride).view(2, 3, 9, ho, wo).permute(0, 3, 4, 2, 1).reshape(2, ho, wo, 27)
        got = out.float().cpu()
        assert torch.equal(got[..., :27], cols) and (got[..., 27:] == 0).all()


@pytest.mark.parametrize("geom", [(1080, 1920, 640, 360, 2), (720, 1280, 640, 360, 2), (480, 640, 640, 480, 2),
                                  (640, 640, 640, 640, 1), (333, 517, 640, 412, 2)],
                         ids=lambda g: "x".join(map(str, g)))
def test_fused_preprocess_patches_equal_preprocess_plus_im2col(lib, geom):
    """letterbox + blob + 3x3 patches in one kernel == the two-kernel path, bit for bit"""
    h, w, new_w, new_h, stride = geom
    frames = torch.from_numpy(np.stack([inputs.frame(40 + i, h, w) for i in range(2)])).cuda()
    ho = wo = (640 + 2 - 3) // stride + 1
    x = torch.empty((2, 640, 640, 4), dtype=torch.float16, device="cuda")
    _lib.check(lib.b2f_preprocess(frames.data_ptr(), 2, h, w, new_w, new_h, 640, 640, 127.5, 1 / 128.0, x.data_ptr(), 4, 0, sp()))
    want = torch.empty((2, ho, wo, 32), dtype=torch.float16, device="cuda")
    _lib.check(lib.b2f_im2col3x3(x.data_ptr(), 2, 640, 640, stride, ho, wo, 0, want.data_ptr(), sp()))
    got = torch.full((2, ho, wo, 32), float("nan"), dtype=torch.float16, device="cuda")
    _lib.check(lib.b2f_preprocess_patches(frames.data_ptr(), 2, h, w, new_w, new_h, 640, 640, stride, 127.5, 1 / 128.0,
                                          got.data_ptr(), 0, sp()))
    torch.cuda.synchronize()
    assert torch.equal(got.view(torch.int16), want.view(torch.int16))


@pytest.mark.parametrize("geom", [(1080, 1920, 640, 360), (720, 1280, 640, 360), (480, 640, 640, 480), (333, 517, 640, 412)],
                         ids=lambda g: "x".join(map(str, g)))
@pytest.mark.parametrize("cout", [28, 16])
def test_fused_preprocess_conv1_matches_patches_plus_conv(lib, geom, cout):
    """letterbox + blob + FIRST CONVOLUTION in one kernel (b2f_preprocess_conv1) against the two-kernel path it
    replaces (b2f_preprocess_patches, bit-exact vs cv2, then the tcgen05 1x1 GEMM over the patches): same exact fp16
    products, fp32 accumulation in a different order -> equal after the 16-bit rounding except for isolated 1-ulp
    flips; and both within 2e-3 of the fp32 convolution of the exact blob"""
    h, w, new_w, new_h = geom
    n = 2
    frames = torch.from_numpy(np.stack([inputs.frame(40 + i, h, w) for i in range(n)])).cuda()
    gen = torch.Generator().manual_seed(cout)
    cout_p = 32 if cout > 16 else 16
    for dtype, tdt in ((0, torch.float16), (1, torch.bfloat16)):
        wt = torch.zeros((1, cout_p, 32), dtype=tdt)
        wt[0, :cout, :27] = (torch.randn((cout, 27), generator=gen) * (2.0 / 27) ** 0.5).to(tdt)
        bias = torch.zeros((1, cout_p))
        bias[0, :cout] = torch.randn(cout, generator=gen) * 0.1
        wt_d, bias_d = wt.cuda(), bias.cuda()
        patches = torch.empty((n, 320, 320, 32), dtype=tdt, device="cuda")
        _lib.check(lib.b2f_preprocess_patches(frames.data_ptr(), n, h, w, new_w, new_h, 640, 640, 2, 127.5, 1 / 128.0,
                                              patches.data_ptr(), dtype, sp()))
        want = torch.empty((n, 320, 320, cout_p), dtype=tdt, device="cuda")
        d = _lib.ConvDesc()
        d.n, d.h, d.w, d.cin_p, d.ho, d.wo, d.cout_p = n, 320, 320, 32, 320, 320, cout_p
        d.kh, d.kw, d.stride, d.pad = 1, 1, 1, 0
        d.dtype, d.out_dtype, d.act, d.bias_classes = dtype, dtype, 1, 1
        d.in_, d.weight, d.bias, d.out = patches.data_ptr(), wt_d.data_ptr(), bias_d.data_ptr(), want.data_ptr()
        _lib.check(lib.b2f_conv2d(C.byref(d), sp()), "b2f_conv2d")
        got = torch.full((n, 320, 320, cout_p), float("nan"), dtype=tdt, device="cuda")
        _lib.check(lib.b2f_preprocess_conv1(frames.data_ptr(), n, h, w, new_w, new_h, 640, 640, 127.5, 1 / 128.0,
                                            wt_d.data_ptr(), bias_d.data_ptr(), cout_p, 1, got.data_ptr(), dtype, sp()))
        torch.cuda.synchronize()
        differ = (got.view(torch.int16) != want.view(torch.int16))
        assert differ.float().mean().item() <= 2e-3, f"{differ.float().mean().item():.2e} of the outputs differ"
        ulp = 2.0 ** -10 if dtype == 0 else 2.0 ** -7
        assert ((got.float() - want.float()).abs() <= ulp * want.float().abs().clamp_min(2.0 ** -14) * 1.01).all()
        ref = torch.relu(patches.float().reshape(-1, 32) @ wt_d[0].float().T + bias_d[0]).reshape(n, 320, 320, cout_p)
        assert (got.float() - ref).abs().max().item() <= (2e-3 if dtype == 0 else 1.6e-2) * max(1.0, ref.abs().max().item())
        assert (got[..., cout:] == 0).all()


def test_fused_norm_crop_patches_equal_norm_crop_plus_im2col(lib):
    h, w, n = 360, 480, 40
    frames = torch.from_numpy(np.stack([inputs.smooth_frame(50 + i, h, w) for i in range(2)])).cuda()
    lm = inputs.landmarks(51, h, w, n)
    lm[:6] += np.array([[-300.0, 0.0]], np.float32)                     # crops that hang over the frame edge
    kps = torch.from_numpy(lm.reshape(n, 10)).cuda()
    fidx = (torch.arange(n, device="cuda") % 2).to(torch.int32)
    scale = float(np.float32(1.0 / 127.5))
    for dtype, tdt in ((0, torch.float16), (1, torch.bfloat16)):
        x = torch.empty((n, 112, 112, 4), dtype=tdt, device="cuda")
        _lib.check(lib.b2f_norm_crop(frames.data_ptr(), h, w, fidx.data_ptr(), kps.data_ptr(), n, 112, 127.5, scale,
                                     x.data_ptr(), 4, dtype, None, None, sp()))
        want = torch.empty((n, 112, 112, 32), dtype=tdt, device="cuda")
        _lib.check(lib.b2f_im2col3x3(x.data_ptr(), n, 112, 112, 1, 112, 112, dtype, want.data_ptr(), sp()))
        got = torch.full((n, 112, 112, 32), float("nan"), dtype=tdt, device="cuda")
        _lib.check(lib.b2f_norm_crop_patches(frames.data_ptr(), h, w, fidx.data_ptr(), kps.data_ptr(), n, 112, 127.5, scale,
                                             got.data_ptr(), dtype, sp()))
        torch.cuda.synchronize()
        assert torch.equal(got.view(torch.int16), want.view(torch.int16))
        # the same crop as 16-byte pixels for the stem-form convolution: bit-identical to norm_crop with c_pad = 8
        want8 = torch.full((n, 112, 112, 8), float("nan"), dtype=tdt, device="cuda")
        _lib.check(lib.b2f_norm_crop(frames.data_ptr(), h, w, fidx.data_ptr(), kps.data_ptr(), n, 112, 127.5, scale,
                                     want8.data_ptr(), 8, dtype, None, None, sp()))
        got8 = torch.full((n, 112, 112, 8), float("nan"), dtype=tdt, device="cuda")
        _lib.check(lib.b2f_norm_crop_image8(frames.data_ptr(), h, w, fidx.data_ptr(), kps.data_ptr(), n, 112, 127.5, scale,
                                            got8.data_ptr(), dtype, sp()))
        torch.cuda.synchronize()
        assert torch.equal(got8.view(torch.int16), want8.view(torch.int16))
        assert torch.equal(got8[..., :3].view(torch.int16), x[..., :3].view(torch.int16)) and (got8[..., 3:] == 0).all()


def test_pool_and_eltwise(lib):
    g = torch.Generator().manual_seed(3)
    x = _q(torch.randn((2, 24, 21, 27), generator=g))
    xin = _nhwc16(x, 32)
    out = torch.empty((2, 11, 14, 32), dtype=torch.float16, device="cuda")
    _lib.check(lib.b2f_pool(xin.data_ptr(), 2, 21, 27, 32, 3, 2, 1, 0, 11, 14, 0, out.data_ptr(), sp()))
    want = F.max_pool2d(x, 3, 2, 1)
    assert torch.equal(out.float().cpu()[..., :24].permute(0, 3, 1, 2), want)
    _lib.check(lib.b2f_pool(xin.data_ptr(), 2, 21, 27, 32, 2, 2, 0, 1, 11, 14, 0, out.data_ptr(), sp()))
    want = F.avg_pool2d(x, 2, 2, 0, ceil_mode=True, count_include_pad=False)
    assert (out.float().cpu()[..., :24].permute(0, 3, 1, 2) - want).abs().max().item() <= 2e-3
    # eltwise: relu(a*scale + shift + b)
    y = _q(torch.randn((2, 24, 21, 27), generator=g))
    sc, sh = torch.zeros(32), torch.zeros(32)
    sc[:24], sh[:24] = torch.rand(24, generator=g) + 0.5, torch.randn(24, generator=g)
    out2 = torch.empty_like(xin)
    keep = [_nhwc16(y, 32), sc.cuda(), sh.cuda()]
    _lib.check(lib.b2f_eltwise(xin.data_ptr(), keep[0].data_ptr(), 2 * 21 * 27, 32, keep[1].data_ptr(),
                               keep[2].data_ptr(), None, 1, 0, out2.data_ptr(), sp()))
    want = torch.relu(x * sc[:24].view(1, -1, 1, 1) + sh[:24].view(1, -1, 1, 1) + y)
    assert (out2.float().cpu()[..., :24].permute(0, 3, 1, 2) - want).abs().max().item() <= 4e-3


# =============================================================================================
# letterbox / blob / align (bit-exact vs cv2)
# =============================================================================================

@pytest.mark.parametrize("size", [(1080, 1920, 640, 640), (720, 1280, 640, 640), (640, 640, 640, 640),
                                  (480, 640, 640, 640), (500, 375, 640, 640), (777, 1333, 640, 640),
                                  (240, 320, 320, 320)])
def test_letterbox_bit_exact(lib, size):
    h, w, in_w, in_h = size
    new_w, new_h, _ = restate.letterbox_geometry(h, w, in_w, in_h)
    frames = np.stack([inputs.frame(1, h, w), inputs.smooth_frame(2, h, w)])
    d = dev(frames)
    out = torch.empty((2, in_h, in_w, 3), dtype=torch.uint8, device="cuda")
    _lib.check(lib.b2f_letterbox_u8(d.data_ptr(), 2, h, w, new_w, new_h, in_w, in_h, out.data_ptr(), sp()))
    got = out.cpu().numpy()
    for i in range(2):
        canvas = np.zeros((in_h, in_w, 3), np.uint8)
        canvas[:new_h, :new_w] = cv2.resize(frames[i], (new_w, new_h))          # reference models/scrfd.py:135-138
        np.testing.assert_array_equal(got[i], canvas)
    # fused normalise + layout: fp16 rounding of the exact float32 blob
    x = torch.empty((2, in_h, in_w, 4), dtype=torch.float16, device="cuda")
    _lib.check(lib.b2f_preprocess(d.data_ptr(), 2, h, w, new_w, new_h, in_w, in_h, 127.5, 1 / 128, x.data_ptr(), 4, 0, sp()))
    blob = restate.blob_from_bgr(got, 1 / 128, 127.5)                            # (2,3,H,W) f32 RGB
    want = torch.from_numpy(blob).permute(0, 2, 3, 1).half()
    assert torch.equal(x[..., :3].cpu(), want) and (x[..., 3] == 0).all()


def test_blob_bit_exact(lib):
    img = np.stack([inputs.frame(3, 112, 112), inputs.frame(4, 112, 112)])
    d = dev(img)
    out = torch.empty((2, 3, 112, 112), dtype=torch.float32, device="cuda")
    _lib.check(lib.b2f_blob_nchw_f32(d.data_ptr(), 2, 112, 112, 127.5, float(np.float32(1 / 127.5)), out.data_ptr(), sp()))
    want = cv2.dnn.blobFromImages(list(img), 1 / 127.5, (112, 112), (127.5,) * 3, swapRB=True)
    np.testing.assert_array_equal(out.cpu().numpy(), want)


def test_warp_affine_bit_exact_and_estimate(lib, golden):
    for tag, (h, w) in (("1080p", (1080, 1920)), ("vga", (480, 640))):
        img = inputs.smooth_frame(7, h, w)
        lms = inputs.landmarks(8, h, w, 6)
        frames = dev(img[None])
        idx = torch.zeros(6, dtype=torch.int32, device="cuda")
        # (1) given the reference's own M: bit-exact crop
        M = dev(golden[f"align_{tag}_M"].reshape(6, 6))
        out = torch.empty((6, 112, 112, 3), dtype=torch.uint8, device="cuda")
        _lib.check(lib.b2f_warp_affine_u8(frames.data_ptr(), h, w, idx.data_ptr(), M.data_ptr(), 6, 112, out.data_ptr(), sp()))
        np.testing.assert_array_equal(out.cpu().numpy(), golden[f"align_{tag}_crop"])
        # (2) closed-form estimate: T1 tolerance on M (<= 1e-6 relative)
        m = torch.empty((6, 6), dtype=torch.float64, device="cuda")
        lm = dev(lms.reshape(6, 10))
        _lib.check(lib.b2f_estimate_norm(lm.data_ptr(), 6, 112, m.data_ptr(), sp()))
        np.testing.assert_allclose(m.cpu().numpy().reshape(6, 2, 3), golden[f"align_{tag}_M"], rtol=1e-9, atol=1e-8)
        # (3) fused landmarks -> crop + normalised NHWC: crop equals cv2 on the kernel's own M
        x = torch.empty((6, 112, 112, 4), dtype=torch.float16, device="cuda")
        crop = torch.empty((6, 112, 112, 3), dtype=torch.uint8, device="cuda")
        m2 = torch.empty((6, 6), dtype=torch.float64, device="cuda")
        _lib.check(lib.b2f_norm_crop(frames.data_ptr(), h, w, idx.data_ptr(), lm.data_ptr(), 6, 112, 127.5,
                                     float(np.float32(1 / 127.5)), x.data_ptr(), 4, 0, crop.data_ptr(), m2.data_ptr(), sp()))
        cr = crop.cpu().numpy()
        for i in range(6):
            np.testing.assert_array_equal(cr[i], cv2.warpAffine(img, m2[i].cpu().numpy().reshape(2, 3), (112, 112), borderValue=0.0))
        # against the reference's crops the only slack is M's last-bit difference: allow <= 0.1 % of pixels off by one
        diff = np.abs(cr.astype(int) - golden[f"align_{tag}_crop"].astype(int))
        assert diff.max() <= 1 and (diff > 0).mean() <= 1e-3
        want = torch.from_numpy(restate.blob_from_bgr(cr, 1 / 127.5, 127.5)).permute(0, 2, 3, 1).half()
        assert torch.equal(x[..., :3].cpu(), want)


# =============================================================================================
# decode + threshold + sort + NMS + max_num (bit-exact)
# =============================================================================================

def run_decode(lib, heads, in_h, in_w, det_scale, image_hw, conf, iou, max_num, metric, max_cand=None, pad_to=None,
               batch=1):
    """heads: 9 arrays in reference layout (for batch>1 a list of such lists).  Returns per-frame results."""
    frames = [heads] if batch == 1 else heads
    lv = _lib.DetLevels()
    keep = []
    for i in range(3):
        for j, (field, psf, width) in enumerate((("score", "score_ps", 2), ("bbox", "bbox_ps", 8), ("kps", "kps_ps", 20))):
            arr = np.stack([f[i + 3 * j].reshape(-1, width) for f in frames])          # [B, pix, width]
            ps = width if pad_to is None else pad_to[j]
            buf = np.full((arr.shape[0], arr.shape[1], ps), 0.25, np.float32)
            buf[..., :width] = arr
            t = dev(buf)
            keep.append(t)
            getattr(lv, field)[i] = t.data_ptr()
            getattr(lv, psf)[i] = ps
    total = sum((in_h // s) * (in_w // s) * 2 for s in (8, 16, 32))
    max_cand = max_cand or total
    max_det = max_cand
    b = len(frames)
    det = torch.zeros((b, max_det, 5), device="cuda")
    kps = torch.zeros((b, max_det, 10), device="cuda")
    kidx = torch.zeros((b, max_det), dtype=torch.int32, device="cuda")
    counts = torch.zeros((b, 4), dtype=torch.int32, device="cuda")
    ws = torch.empty(int(lib.b2f_decode_nms_workspace(b, max_cand)), dtype=torch.uint8, device="cuda")
    scale = dev(np.full(b, det_scale, np.float32))
    hw = dev(np.tile(np.asarray(image_hw, np.int32), (b, 1)))
    _lib.check(lib.b2f_decode_nms(C.byref(lv), b, in_h, in_w, scale.data_ptr(), hw.data_ptr(), conf, iou, max_num,
                                  0 if metric == "max" else 1, max_cand, max_det, det.data_ptr(), kps.data_ptr(),
                                  kidx.data_ptr(), counts.data_ptr(), ws.data_ptr(), ws.numel(), sp()))
    torch.cuda.synchronize()
    cnt = counts.cpu().numpy()
    res = []
    for f in range(b):
        n = cnt[f, 0]
        res.append((det[f, :n].cpu().numpy(), kps[f, :n].cpu().numpy().reshape(-1, 5, 2), kidx[f, :n].cpu().numpy(), cnt[f]))
    return res[0] if batch == 1 else res


@pytest.mark.parametrize("case", postprocess_cases(), ids=lambda c: c[0])
def test_decode_nms_matches_reference_outputs(lib, golden, case):
    name, (ih, iw), in_size, seed, ties, max_num, metric = case
    heads = inputs.head_tensors(seed, in_size[1], in_size[0], ties)
    _, _, ds = restate.letterbox_geometry(ih, iw, in_size[0], in_size[1])
    det, kps, kidx, cnt = run_decode(lib, heads, in_size[1], in_size[0], np.float32(ds), (ih, iw), 0.5,
                                     float(np.float32(0.4)), max_num, metric)
    np.testing.assert_array_equal(det, golden[name + "_det"])                  # bit-exact boxes + scores
    np.testing.assert_array_equal(kps, golden[name + "_kps"])                  # bit-exact landmarks
    assert cnt[1] == int(golden[name + "_ncand"]) and cnt[2] == len(golden[name + "_keep"]) and cnt[3] == 0
    if max_num == 0:
        np.testing.assert_array_equal(kidx, golden[name + "_keep"])            # NMS keep indices, bit-exact


@pytest.mark.parametrize("seed,mu,ties,pad", [(100, -4.0, 0.0, None), (101, -2.5, 0.3, (16, 16, 32)),
                                              (102, -1.0, 0.5, (2, 8, 20)), (103, 1.0, 0.0, (16, 16, 32))])
def test_decode_nms_matches_oracle_with_ties_and_padding(lib, seed, mu, ties, pad):
    """heavier candidate loads (hundreds .. ~12k), duplicated scores, channel-padded head layout"""
    heads = inputs.head_tensors(seed, 640, 640, ties, score_mu=mu)
    for max_num, metric in ((0, "max"), (7, "max"), (9, "default")):
        want_d, want_k = restate.scrfd_postprocess(heads, 640, 640, 1 / 3, 0.5, 0.4, max_num, metric, (1080, 1920))
        det, kps, kidx, cnt = run_decode(lib, heads, 640, 640, np.float32(1 / 3), (1080, 1920), 0.5,
                                         float(np.float32(0.4)), max_num, metric, pad_to=pad)
        np.testing.assert_array_equal(det, want_d)
        np.testing.assert_array_equal(kps, want_k)


def test_decode_nms_edge_cases(lib):
    # nothing over the threshold -> empty, no error
    heads = inputs.head_tensors(0, 320, 320, score_mu=-30.0)
    det, kps, kidx, cnt = run_decode(lib, heads, 320, 320, np.float32(1.0), (320, 320), 0.5, 0.4, 0, "max")
    assert det.shape == (0, 5) and kps.shape == (0, 5, 2) and list(cnt) == [0, 0, 0, 0]
    # candidate overflow is flagged, never silent
    heads = inputs.head_tensors(1, 320, 320, score_mu=0.0)
    det, kps, kidx, cnt = run_decode(lib, heads, 320, 320, np.float32(1.0), (320, 320), 0.5, 0.4, 0, "max", max_cand=64)
    assert cnt[1] > 64 and (cnt[3] & 1)
    # ... and deterministic: the candidates kept are the first 64 in anchor order (ordered ballot compaction), i.e. the
    # result equals the oracle run on the head tensors with every later candidate pushed under the threshold
    clipped = [h.copy() for h in heads]
    seen = 0
    for lvl in range(3):
        over = np.nonzero(clipped[lvl][:, 0] >= 0.5)[0]
        drop = over[max(0, 64 - seen):]
        clipped[lvl][drop, 0] = 0.0
        seen += len(over)
    want_d, want_k = restate.scrfd_postprocess(clipped, 320, 320, 1.0, 0.5, 0.4)
    np.testing.assert_array_equal(det, want_d)
    np.testing.assert_array_equal(kps, want_k)
    for _ in range(3):
        det2, kps2, _, _ = run_decode(lib, heads, 320, 320, np.float32(1.0), (320, 320), 0.5, 0.4, 0, "max", max_cand=64)
        np.testing.assert_array_equal(det2, det)
    # degenerate boxes: 0/0 overlap is NaN and must suppress, as np.where(ovr <= thr) does
    heads = inputs.head_tensors(2, 320, 320, score_mu=-2.0)
    for i in (3, 4, 5):
        heads[i][:, 0] = -heads[i][:, 2] - 1.0 / (8 << (i - 3))      # x2 - x1 + 1 == 0 before det_scale
    want_d, want_k = restate.scrfd_postprocess(heads, 320, 320, 1.0, 0.5, 0.4)
    det, kps, kidx, cnt = run_decode(lib, heads, 320, 320, np.float32(1.0), (320, 320), 0.5, float(np.float32(0.4)), 0, "max")
    np.testing.assert_array_equal(det, want_d)
    # batched launch == per-frame launches
    frames = [inputs.head_tensors(10 + i, 320, 320, score_mu=-3.0) for i in range(5)]
    res = run_decode(lib, frames, 320, 320, np.float32(0.5), (640, 640), 0.5, float(np.float32(0.4)), 4, "max", batch=5)
    for f, r in zip(frames, res):
        want_d, want_k = restate.scrfd_postprocess(f, 320, 320, 0.5, 0.5, 0.4, 4, "max", (640, 640))
        np.testing.assert_array_equal(r[0], want_d)
        np.testing.assert_array_equal(r[1], want_k)


def test_forward_view_and_standalone_nms(lib):
    heads = inputs.head_tensors(5, 320, 320, score_mu=-3.0)
    det, kps, anchor, cnt = run_decode(lib, heads, 320, 320, np.float32(1.0), (320, 320), 0.5, -1.0, 0, "max")
    sl, bl, kl = [], [], []
    for i, s in enumerate((8, 16, 32)):
        a, b, c = restate.decode_level(heads[i], heads[i + 3], heads[i + 6], s, 320, 320, 0.5)
        sl.append(a), bl.append(b), kl.append(c)
    np.testing.assert_array_equal(det[:, :4], np.vstack(bl))
    np.testing.assert_array_equal(det[:, 4:], np.vstack(sl))
    np.testing.assert_array_equal(kps, np.vstack(kl))
    assert (np.diff(anchor) > 0).all()
    # stand-alone NMS on an arbitrary array with duplicate scores
    rng = np.random.default_rng(0)
    d = np.zeros((700, 5), np.float32)
    d[:, :2] = rng.uniform(0, 300, (700, 2))
    d[:, 2:4] = d[:, :2] + rng.uniform(5, 80, (700, 2))
    d[:, 4] = rng.choice(np.linspace(0.1, 0.9, 50).astype(np.float32), 700)
    t = dev(d)
    keep = torch.empty(700, dtype=torch.int32, device="cuda")
    n_keep = torch.zeros(1, dtype=torch.int32, device="cuda")
    ws = torch.empty(1024 * 8, dtype=torch.uint8, device="cuda")
    _lib.check(lib.b2f_nms(t.data_ptr(), 700, float(np.float32(0.4)), keep.data_ptr(), n_keep.data_ptr(), ws.data_ptr(), ws.numel(), sp()))
    got = keep[:int(n_keep.item())].cpu().numpy()
    np.testing.assert_array_equal(got, np.asarray(restate.nms(d, 0.4)))


# =============================================================================================
# matching / clustering
# =============================================================================================

def test_l2norm_and_cosine(lib):
    x = inputs.embeddings(1, 37) * 7
    d = dev(x)
    f32 = torch.empty_like(d)
    h16 = torch.empty((37, 512), dtype=torch.float16, device="cuda")
    nrm = torch.empty(37, device="cuda")
    _lib.check(lib.b2f_l2norm_rows(d.data_ptr(), 37, 512, f32.data_ptr(), h16.data_ptr(), 0, nrm.data_ptr(), sp()))
    want = restate.normalize_rows(x)
    np.testing.assert_allclose(f32.cpu().numpy(), want, rtol=0, atol=1e-6)      # fp32, different summation order
    np.testing.assert_allclose(nrm.cpu().numpy(), np.linalg.norm(x, axis=1), rtol=1e-6)
    a, b = inputs.embeddings(2, 16), inputs.embeddings(3, 16)
    out = torch.empty(16, device="cuda")
    da, db = dev(a), dev(b)
    _lib.check(lib.b2f_cosine_pairs(da.data_ptr(), db.data_ptr(), 16, 512, out.data_ptr(), sp()))
    want = np.array([restate.compute_similarity(u, v) for u, v in zip(a, b)])
    np.testing.assert_allclose(out.cpu().numpy(), want, rtol=0, atol=1e-6)       # fp32, different summation order


@pytest.mark.parametrize("q,g,k", [(5, 64, 1), (200, 3000, 5), (130, 70001, 8)])
def test_match_topk_identities_bit_exact(lib, q, g, k):
    from scrfd_arcface_facerecognition_b200.gallery import Gallery
    gal = inputs.embeddings(20, g)
    qs, ids = inputs.planted_queries(gal, 21, q)
    G = Gallery()
    G.add(gal)
    s, i = G.match(torch.from_numpy(qs).cuda(), k)
    s, i = s.cpu().numpy(), i.cpu().numpy()
    gn, qn = restate.normalize_rows(gal).astype(np.float64), restate.normalize_rows(qs).astype(np.float64)
    sims = qn @ gn.T
    order = np.argsort(-sims, axis=1, kind="stable")[:, :k]
    np.testing.assert_array_equal(i[:, 0], ids)                                   # planted identities
    np.testing.assert_array_equal(i[:, 0], order[:, 0])                           # top-1 bit-exact
    top = np.take_along_axis(sims, order, 1)
    # ranks 2..k: exact wherever neighbouring scores are separated by >= 1e-3 (SURVEY 8c tier T0)
    gap_ok = np.ones_like(order, bool)
    full = np.sort(sims, axis=1)[:, ::-1][:, :k + 1]
    gap_ok[:, :] = (full[:, :-1] - full[:, 1:] >= 1e-3)
    gap_ok[:, 1:] &= gap_ok[:, :-1]
    assert (i == order)[gap_ok].all()
    np.testing.assert_allclose(s[gap_ok], top[gap_ok], rtol=0, atol=2e-6)
    # threshold semantics of QdrantManager.search_similar: score >= thr, descending
    s2, i2 = G.match(torch.from_numpy(qs).cuda(), k, threshold=0.5)
    s2, i2 = s2.cpu().numpy(), i2.cpu().numpy()
    assert ((i2 >= 0) == (s >= 0.5)).all() and (np.diff(np.where(i2 >= 0, s2, -1), axis=1) <= 0).all()


def test_best_match_and_search_similar_semantics(lib, golden):
    from scrfd_arcface_facerecognition_b200.gallery import Gallery
    gal = inputs.embeddings(11, 64)
    qs, ids = inputs.planted_queries(gal, 12, 16)
    G = Gallery()
    G.add(gal, payloads=[{"person_id": 1000 + i, "name": f"p{i}"} for i in range(64)])
    best = [G.best_match(q, 0.4)[0] for q in qs]
    np.testing.assert_array_equal(best, golden["best_match"])                    # reference main.py:136-142 scan
    assert G.best_match(inputs.embeddings(99, 1)[0], 0.4) == (-1, 0.0)            # "Unknown"
    res = G.search_similar(qs[0], k=5, threshold=0.35)
    assert res and res[0]["person_id"] == 1000 + ids[0] and res[0]["name"] == f"p{ids[0]}"
    idx, sc = restate.search_similar(qs[0], gal, 5, 0.35)
    assert [r["person_id"] - 1000 for r in res] == list(idx)
    np.testing.assert_allclose([r["similarity"] for r in res], sc, atol=2e-6)


@pytest.mark.parametrize("q,g,k", [(129, 5000, 1), (700, 70001, 5), (2048, 300_000, 1)])
def test_pair_kernel_matches_first_generation_kernel(lib, q, g, k):
    """match_pair_kernel (persistent CTA pairs, cta_group::2) against umma_conv_kernel<EPI_TOPK> on the same inputs:
    identical identities AND identical scores (the exact fp32 re-score makes them bit-equal), ragged last tiles included"""
    from scrfd_arcface_facerecognition_b200.gallery import Gallery
    gen = torch.Generator(device="cuda").manual_seed(q)
    gal = torch.randn((g, 512), generator=gen, device="cuda")
    ids = torch.randint(0, g, (q,), generator=gen, device="cuda")
    qs = gal[ids] + 0.9 * torch.randn((q, 512), generator=gen, device="cuda")
    G = Gallery()
    G.add(gal)
    try:
        _lib.call("b2f_set_tuning", 17, 0)
        s0, i0 = (t.clone() for t in G.match(qs, k))
        _lib.call("b2f_set_tuning", 17, 1)
        s1, i1 = (t.clone() for t in G.match(qs, k))
    finally:
        _lib.call("b2f_set_tuning", 17, 1)
    assert torch.equal(i1[:, 0], ids) and torch.equal(i0, i1) and torch.equal(s0, s1)


def test_pair_kernel_pairs_match_first_generation_kernel(lib):
    """thresholded all-pairs: same sorted pair list from both kernels, whole matrix and an unaligned row block"""
    from scrfd_arcface_facerecognition_b200.gallery import Gallery
    emb = inputs.clustered(31, 900, 3, noise=0.45)
    G = Gallery()
    G.add(emb)
    out = {}
    try:
        for gen in (0, 1):
            _lib.call("b2f_set_tuning", 17, gen)
            out[gen] = (G.duplicate_pairs(0.8).clone(), G.duplicate_pairs(0.8, 333, 1900).clone())
    finally:
        _lib.call("b2f_set_tuning", 17, 1)
    assert out[0][0].numel() > 1000 and torch.equal(out[0][0], out[1][0]) and torch.equal(out[0][1], out[1][1])


@pytest.mark.parametrize("n", [2, 130, 700, 5000])
def test_causal_match_is_a_prefix_search(lib, n):
    """b2f_match_partial_causal: row i against the rows BEFORE it only (the online loop's best earlier person,
    reference duplicate.py:1853-1855, for all rows in one pass) == the sequential numpy prefix scan"""
    from scrfd_arcface_facerecognition_b200.gallery import Gallery
    emb = inputs.clustered(40 + n, n // 3 + 1, 3, noise=0.5)[:n]
    assert len(emb) == n
    G = Gallery()
    G.add(emb)
    s, i = G.match_local(G.f32, 1, 0.3, causal_base=0)
    s, i = s[:, 0].cpu().numpy(), i[:, 0].cpu().numpy()
    u = restate.normalize_rows(emb).astype(np.float64)
    sims = np.tril(u @ u.T, -1) + np.triu(np.full((n, n), -2.0))             # only j < i
    want_i = sims.argmax(1)
    want_s = sims[np.arange(n), want_i]
    ok = want_s >= 0.3
    ok[0] = False
    np.testing.assert_array_equal(i >= 0, ok)
    gap = np.sort(sims, axis=1)[:, -1] - np.sort(sims, axis=1)[:, -2] if n > 2 else np.ones(n)
    sure = ok & (gap >= 1e-3)
    np.testing.assert_array_equal(i[sure], want_i[sure])
    np.testing.assert_allclose(s[ok], want_s[ok], rtol=0, atol=2e-6)


def test_exhaustive_search_uses_the_rows_dot_kernel(lib):
    """QdrantManager-shaped search with k beyond the running top-8 (duplicate.py:2757-2766 asks for every row): exact
    fp32 scores from b2f_rows_dot, ordered like the reference (score desc), thresholded with >="""
    from scrfd_arcface_facerecognition_b200.vector_store import GalleryManager
    emb = inputs.clustered(61, 30, 4, noise=0.4)
    mgr = GalleryManager({"vector_database": {}})
    for j in range(len(emb)):
        assert mgr.add_embedding(j, emb[j], {"name": f"p{j}"})
    before = _lib.launch_count()
    res = mgr.search_similar(emb[7], k=len(emb), threshold=0.5)
    assert _lib.launch_count() > before
    idx, sc = restate.search_similar(emb[7], emb, len(emb), 0.5)
    assert [r["person_id"] for r in res] == list(idx) and res[0]["person_id"] == 7
    np.testing.assert_allclose([r["similarity"] for r in res], sc, rtol=0, atol=2e-6)


def test_top1_key_kernels_match_host_arithmetic(lib):
    """b2f_topk_pack_keys / b2f_topk_unpack_keys == the torch restatement used by the gloo tests, bit for bit"""
    from scrfd_arcface_facerecognition_b200.gallery import pack_top1_keys, unpack_top1_keys
    g = torch.Generator().manual_seed(9)
    s = torch.rand(5000, generator=g) * 2 - 1
    i = torch.randint(0, 2 ** 32 - 1, (5000,), generator=g)
    i[torch.rand(5000, generator=g) < 0.1] = -1
    s[:3] = torch.tensor([0.0, 1.0, -1.0])
    host = pack_top1_keys(s, i)
    devk = pack_top1_keys(s.cuda(), i.cuda())
    assert torch.equal(devk.cpu(), host)
    ds, di = unpack_top1_keys(devk)
    hs, hi = unpack_top1_keys(host)
    assert torch.equal(di.cpu(), hi) and torch.equal(ds.cpu().view(torch.int32), hs.view(torch.int32))
    assert torch.equal(hi, i) and torch.equal(hs[i >= 0], s[i >= 0]) and (hs[i < 0] == 0).all()
    order = torch.argsort(host)                        # key order == (score asc, index desc)
    so, io = s[order][i[order] >= 0], i[order][i[order] >= 0]
    assert ((so[1:] > so[:-1]) | ((so[1:] == so[:-1]) & (io[1:] < io[:-1]))).all()


@pytest.mark.parametrize("centres,members", [(50, 4), (700, 3)])
def test_duplicate_merge_matches_oracle(lib, centres, members):
    from scrfd_arcface_facerecognition_b200.gallery import Gallery
    emb = inputs.clustered(30, centres, members)
    G = Gallery()
    G.add(emb)
    leader = G.merge_duplicates(0.8)
    np.testing.assert_array_equal(leader, restate.merge_duplicates(emb, 0.8))
    assert len(np.unique(leader)) == centres
    # chain case: one-hop only (no transitive closure)
    a = np.zeros(512, np.float32); a[0] = 1
    b = np.zeros(512, np.float32); b[0], b[1] = 0.8, 0.6
    c = np.zeros(512, np.float32); c[0], c[1] = 0.28, 0.96
    G2 = Gallery()
    G2.add(np.stack([a, b, c]))
    assert list(G2.merge_duplicates(0.8)) == [0, 0, 2]


@pytest.mark.parametrize("centres,members,noise", [(60, 4, 0.35), (300, 3, 0.5), (5, 40, 0.62)])
def test_online_clusters_match_oracle(lib, centres, members, noise):
    """row a20: the online per-visit decision (reference duplicate.py:1853-1949) == sequential oracle, bit for bit"""
    from scrfd_arcface_facerecognition_b200.gallery import Gallery
    emb = inputs.clustered(33, centres, members, noise=noise)
    G = Gallery()
    G.add(emb)
    got = G.online_clusters(0.8 if noise < 0.6 else 0.7)
    want = restate.online_clusters(emb, 0.8 if noise < 0.6 else 0.7)
    np.testing.assert_array_equal(got, want)
    assert Gallery().online_clusters(0.8).shape == (0,)


@pytest.mark.parametrize("case", inputs.cluster_cases(), ids=lambda c: c[0])
def test_clustering_matches_reference_run(lib, case):
    """rows a19-a21 against the REFERENCE's own code run verbatim (tests/golden/make_cluster_golden.py: duplicate.py's
    online loop, find_and_merge_duplicates and QdrantManager.search_similar over a restated qdrant local mode).
    Identities / labels / leaders bit-exact; similarities within fp32 summation-order noise (2e-6)."""
    import os
    from scrfd_arcface_facerecognition_b200.gallery import Gallery
    name, rows = case
    g = dict(np.load(os.path.join(os.path.dirname(__file__), "golden", "cluster_outputs.npz")))
    grouping, search, duplicate, merge = (float(v) for v in g["thresholds"])
    G = Gallery()
    G.add(rows)
    label = G.online_clusters(grouping, search_threshold=search)
    np.testing.assert_array_equal(label, g[f"online_{name}_label"])
    np.testing.assert_allclose(G.online_similarities(label, search), g[f"online_{name}_sim"], rtol=0, atol=2e-6)
    label = G.online_clusters(grouping, duplicate_threshold=duplicate, search_threshold=search)
    np.testing.assert_array_equal(label, g[f"online_dup_{name}_label"])
    np.testing.assert_allclose(G.online_similarities(label, search), g[f"online_dup_{name}_sim"], rtol=0, atol=2e-6)
    np.testing.assert_array_equal(G.merge_duplicates(merge), g[f"merge_{name}_leader"])
    qs, _ = inputs.planted_queries(rows, 90, 12, noise=0.8)
    for qi, q in enumerate(qs):
        res = G.search_similar(q, k=5, threshold=search)
        gi, gs = g[f"search_{name}_idx"][qi], g[f"search_{name}_score"][qi]
        k = int((gi >= 0).sum())
        assert [r["person_id"] for r in res] == list(gi[:k])
        np.testing.assert_allclose([r["similarity"] for r in res], gs[:k], rtol=0, atol=2e-6)


@pytest.mark.gpu
def test_clustering_results_written_from_gpu_labels(lib, tmp_path):
    """GPU labels -> the reference's SQLite rows and clustering_results JSON; identical to the same
    writer fed by the sequential CPU oracle, similarities within fp32 dot-order noise (1e-6)."""
    import json
    import sqlite3
    from scrfd_arcface_facerecognition_b200 import result_store as rs
    from scrfd_arcface_facerecognition_b200.vector_store import GalleryManager
    emb = inputs.clustered(35, 40, 5, noise=0.4)
    n = len(emb)
    visits = [{"id": f"v{i}", "customerId": f"c{i}", "image": f"http://example.invalid/{i}.jpg", "entryTime": f"t{i}"} for i in range(n)]
    mgr = GalleryManager({"vector_database": {}})
    for i in range(n):
        assert mgr.add_embedding(i, emb[i], {"name": f"p{i}"})
    want_label = restate.online_clusters(emb, 0.75)
    want_sim = restate.online_similarities(emb, want_label)
    got_sim = mgr.gallery.online_similarities(mgr.gallery.online_clusters(0.75))
    np.testing.assert_allclose(got_sim, want_sim, rtol=0, atol=2e-6)
    conn_g, conn_o = sqlite3.connect(":memory:"), sqlite3.connect(":memory:")
    clock = lambda: 1767261600.0
    out_g = mgr.write_clustering_results(visits, 0.75, rs.PersonDatabase(connection=conn_g), str(tmp_path / "gpu"), clock=clock)
    out_o = rs.write_online_clustering(visits, want_label, want_sim, rs.PersonDatabase(connection=conn_o), str(tmp_path / "cpu"), clock=clock)
    assert out_g["results"] == out_o["results"] and out_g["person_ids"] == out_o["person_ids"]
    assert out_g["results"]["new_persons"] == 40 and out_g["results"]["recognized"] == n - 40
    q = "SELECT person_id, visit_id, customer_id, entry_time, image_url FROM person_visits ORDER BY id"
    assert conn_g.execute(q).fetchall() == conn_o.execute(q).fetchall()
    q = "SELECT id, name, image_path, match_count FROM persons ORDER BY id"
    assert conn_g.execute(q).fetchall() == conn_o.execute(q).fetchall()
    pg, po = (json.load(open(o["json_path"])) for o in (out_g, out_o))
    assert len(pg["groups"]) == n
    for a, b in zip(pg["groups"], po["groups"]):
        assert {k: v for k, v in a.items() if k not in ("group_score", "visits")} == {k: v for k, v in b.items() if k not in ("group_score", "visits")}
        assert abs(a["group_score"] - b["group_score"]) <= 1e-3 + 1e-9
        assert abs(a["visits"][0]["similarity"] - b["visits"][0]["similarity"]) <= 2e-6


@pytest.mark.parametrize("n,cin,k", [(4, 512, 7), (300, 512, 7), (1024, 512, 7), (130, 64, 5), (5, 256, 4)], ids=lambda v: str(v))
def test_conv2d_split_k(lib, n, cin, k):
    """the embedding layer (Flatten + Gemm as a k x k valid convolution over a k x k map) with a split-K workspace: the
    filter taps are dealt to several work items per tile and a second kernel adds the fp32 partial sums; same result as
    the unsplit launch up to the order of the fp32 additions, and independent of the batch"""
    g = torch.Generator().manual_seed(n + cin + k)
    x = _q(torch.randn((n, cin, k, k), generator=g))
    w = _q(torch.randn((512, cin, k, k), generator=g) * (1.0 / (cin * k * k)) ** 0.5)
    b = torch.randn(512, generator=g) * 0.1
    ref = F.conv2d(x, w, b)
    plain = run_conv(lib, x, w, b, 1, 0, out_f32=True)
    assert run_conv.launches == 1
    split = run_conv(lib, x, w, b, 1, 0, out_f32=True, splitk=True)
    assert run_conv.launches == (2 if k * k * (cin // 64) >= 128 else 1)       # split only when K is long
    scale = max(1.0, ref.abs().max().item())
    assert (split - ref).abs().max().item() <= 2e-3 * scale
    assert (split - plain).abs().max().item() <= 1e-4 * scale                # fp32 summation order only (K = 25 088)
    first = run_conv(lib, x[:1], w, b, 1, 0, out_f32=True, splitk=True)      # an image's bits do not depend on its batch mates
    assert torch.equal(first[0], split[0])


@pytest.mark.parametrize("n,cin,k,f32", [(4, 64, 7, True), (4, 64, 7, False), (3, 512, 3, True), (1, 128, 1, True)],
                         ids=lambda v: str(v))
def test_tile_kernel_problem_smaller_than_one_tile(lib, n, cin, k, f32):
    """fewer output pixels than the 128 rows of one MMA tile (the embedding layer at batch 4): the operand slot must
    still span 128 rows -- a short slot let the last stage's read run off the end of shared memory"""
    g = torch.Generator().manual_seed(n * 100 + cin + k)
    x = _q(torch.randn((n, cin, k, k), generator=g))
    w = _q(torch.randn((512, cin, k, k), generator=g) * (1.0 / (cin * k * k)) ** 0.5)
    b = torch.randn(512, generator=g) * 0.1
    ref = F.conv2d(x, w, b)
    try:
        for gen in (3, 2):                                   # conv_tile_kernel forced, then default dispatch
            _lib.check(lib.b2f_set_tuning(2, gen))
            out = run_conv(lib, x, w, b, 1, 0, out_f32=f32)
            assert (out - ref).abs().max().item() <= (2e-3 if f32 else 4e-3) * max(1.0, ref.abs().max().item())
    finally:
        _lib.check(lib.b2f_set_tuning(2, 2))


@pytest.mark.parametrize("case", [(2, 56, 56, 64, 64, 0), (3, 28, 28, 128, 128, 0), (2, 14, 14, 256, 256, 0), (2, 33, 21, 96, 80, 0),
                                  (2, 28, 28, 128, 128, 1), (1, 7, 7, 512, 512, 0), (130, 7, 7, 64, 64, 0)],
                         ids=lambda c: "x".join(map(str, c)))
def test_conv2d_in_place_block_output(lib, case):
    """residual == out (the engine's in-place block outputs): narrow tiles without activation add through the TMA
    reduce-store (16-bit add of the rounded conv result), everything else reads and overwrites its own elements;
    both must equal conv + residual within fp16 rounding of the sum (tolerance 4e-3 * max|ref|)."""
    n, h, w, cin, cout, act = case
    g = torch.Generator().manual_seed(sum(case))
    x = _q(torch.randn((n, cin, h, w), generator=g))
    wt = _q(torch.randn((cout, cin, 3, 3), generator=g) * (2.0 / (cin * 9)) ** 0.5)
    b = torch.randn(cout, generator=g) * 0.1
    res = _q(torch.randn((n, cout, h, w), generator=g))
    ref = F.conv2d(x, wt, b, 1, 1) + res
    if act:
        ref = torch.relu(ref)
    tol = 4e-3 * max(1.0, ref.abs().max().item())
    try:
        outs = {}
        for key16, gen in ((1, 2), (0, 2), (1, 3), (1, 1), (1, 0)):   # reduce-store on / off, forced tile kernel, first-generation kernels
            _lib.check(lib.b2f_set_tuning(16, key16))
            _lib.check(lib.b2f_set_tuning(2, gen))
            out = run_conv(lib, x, wt, b, 1, 1, act=act, residual=res, res_mode=1, in_place=True)
            assert (out - ref).abs().max().item() <= tol, (key16, gen)
            outs[(key16, gen)] = out
        # wide layers go to the first persistent kernel when the batch is small: whichever kernel takes the layer, the
        # in-place sum must have the same bits (narrow layers always run in conv_tile_kernel, whose halo modes
        # accumulate the taps in another order than the first-generation kernels)
        assert torch.equal(outs[(1, 2)], outs[(1, 3)])
        if cout > 128:
            for k in ((1, 1), (1, 0)):
                assert torch.equal(outs[(1, 2)], outs[k]), k
    finally:
        _lib.check(lib.b2f_set_tuning(16, 1))
        _lib.check(lib.b2f_set_tuning(2, 2))


@pytest.mark.parametrize("n,h,w,cout,stride,act", [(3, 112, 112, 64, 1, 2), (2, 64, 48, 32, 2, 1), (130, 16, 16, 64, 1, 0), (1, 33, 21, 128, 1, 2)],
                         ids=lambda v: str(v))
def test_stem8_conv_matches_fp32_reference(lib, n, h, w, cout, stride, act):
    """the 8-channel stem form of b2f_conv2d (16-byte pixels, one TMA box per filter tap into no-swizzle slots, K = 16
    steps spanning two taps): equals conv3x3(pad 1) of the three real channels; fp16 operands, fp32 accumulation,
    tolerance 4e-3 * max|ref|"""
    g = torch.Generator().manual_seed(n + h + cout)
    x = _q(torch.randn((n, 3, h, w), generator=g))
    wt = _q(torch.randn((cout, 3, 3, 3), generator=g) * 0.3)
    b = torch.randn(cout, generator=g) * 0.1
    slope = torch.rand(cout, generator=g) * 0.5
    ref = F.conv2d(x, wt, b, stride, 1)
    ref = torch.relu(ref) if act == 1 else (torch.where(ref >= 0, ref, ref * slope[None, :, None, None]) if act == 2 else ref)
    ho, wo = ref.shape[2], ref.shape[3]
    xin = torch.zeros((n, h, w, 8), dtype=torch.float16)
    xin[..., :3] = x.permute(0, 2, 3, 1).half()
    wk = torch.zeros((10, cout, 8), dtype=torch.float16)                    # [slot = tap][cout][8], slot 9 zero
    wk[:9, :, :3] = wt.permute(2, 3, 0, 1).reshape(9, cout, 3).half()
    d = _lib.ConvDesc()
    d.n, d.h, d.w, d.cin_p, d.ho, d.wo, d.cout_p = n, h, w, 8, ho, wo, cout
    d.kh, d.kw, d.stride, d.pad = 3, 3, stride, 1
    d.dtype, d.out_dtype, d.act, d.bias_classes = 0, 0, act, 1
    keep = [xin.cuda(), wk.cuda(), b.reshape(1, cout).float().cuda(), slope.float().cuda()]
    d.in_, d.weight, d.bias, d.slope = (t.data_ptr() for t in keep)
    out = torch.full((n, ho, wo, cout), float("nan"), dtype=torch.float16, device="cuda")
    d.out = out.data_ptr()
    _lib.check(lib.b2f_conv2d(C.byref(d), sp()), "b2f_conv2d")
    torch.cuda.synchronize()
    got = out.float().cpu().permute(0, 3, 1, 2)
    assert torch.isfinite(got).all()
    assert (got - ref).abs().max().item() <= 4e-3 * max(1.0, ref.abs().max().item())
