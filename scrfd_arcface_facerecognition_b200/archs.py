"""Architecture specs -> ONNX graphs with seeded random weights.

The reference ships no weights (weights/.gitkeep only; download.sh:12-16 fetches
det_500m / det_2.5g / det_10g / w600k_mbf / w600k_r50 .onnx).  Offline, the engine
instantiates the same architectures with deterministic random initialisers and hands them
to the rest of the stack *in ONNX form*, so the loader / graph compiler see exactly what a
real file would give them.  Real files, when present, bypass this module.

What the reference code pins about the graphs (models/scrfd.py:39-45,89-94; arcface.py:22-37):
  * SCRFD: one NCHW float input, 9 outputs ordered [score8,score16,score32,bbox8,..,kps8,..],
    each 2-D (H/s*W/s*2, {1,4,10}); scores already sigmoid-ed; bbox/kps in stride units.
  * ArcFace: input [N,3,112,112], a single [N,512] output.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from .onnx_wire import Graph, Node, ValueInfo

__all__ = ["build_arch", "KNOWN_WEIGHTS", "arch_for_path", "count_macs"]

# file name (reference download.sh:12-16) -> architecture key
KNOWN_WEIGHTS = {
    "det_500m.onnx": "scrfd_500m",
    "det_2.5g.onnx": "scrfd_2.5g",
    "det_10g.onnx": "scrfd_10g",
    "w600k_mbf.onnx": "arcface_mbf",
    "w600k_r50.onnx": "arcface_r50",
}


def arch_for_path(path: str) -> Optional[str]:
    import os
    return KNOWN_WEIGHTS.get(os.path.basename(path))


class _GB:
    """Tiny ONNX graph builder with He-normal conv init and randomised BN statistics."""

    def __init__(self, seed: int, name: str):
        self.rng = np.random.default_rng(seed)
        self.nodes: List[Node] = []
        self.init: Dict[str, np.ndarray] = {}
        self.n = 0
        self.name = name
        # running estimate of each tensor's mean square, so BN statistics and head gains can be
        # set the way a trained network's would be (keeps activations O(1) without a data pass)
        self.m2: Dict[str, float] = {"input.1": 1.0 / 3.0}

    def _t(self, prefix: str) -> str:
        self.n += 1
        return f"{prefix}_{self.n}"

    def _add_init(self, prefix: str, arr: np.ndarray) -> str:
        name = self._t(prefix)
        self.init[name] = np.ascontiguousarray(arr)
        return name

    def conv(self, x: str, cin: int, cout: int, k: int, s: int = 1, p: Optional[int] = None,
             groups: int = 1, bias: bool = False, gain: float = 1.0,
             bias_value: Optional[np.ndarray] = None, out_std: Optional[float] = None) -> str:
        if p is None:
            p = k // 2
        fan_in = (cin // groups) * k * k
        if out_std is not None:
            gain = out_std / float(np.sqrt(2.0 * self.m2[x]))
        w = self.rng.standard_normal((cout, cin // groups, k, k), dtype=np.float32)
        w *= np.float32(gain * np.sqrt(2.0 / fan_in))
        ins = [x, self._add_init("w", w)]
        if bias or bias_value is not None:
            b = (bias_value.astype(np.float32) if bias_value is not None
                 else (0.05 * self.rng.standard_normal(cout)).astype(np.float32))
            ins.append(self._add_init("b", b))
        out = self._t("conv")
        self.m2[out] = 2.0 * gain * gain * self.m2[x]
        self.nodes.append(Node("Conv", ins, [out], {
            "dilations": [1, 1], "group": groups, "kernel_shape": [k, k],
            "pads": [p, p, p, p], "strides": [s, s]}))
        return out

    def bn(self, x: str, c: int, gamma_scale: float = 1.0) -> str:
        g = (gamma_scale * self.rng.uniform(0.8, 1.2, c)).astype(np.float32)
        b = (0.1 * self.rng.standard_normal(c)).astype(np.float32)
        sd = float(np.sqrt(self.m2[x]))
        m = (0.1 * sd * self.rng.standard_normal(c)).astype(np.float32)
        v = (self.m2[x] * self.rng.uniform(0.8, 1.25, c)).astype(np.float32)
        out = self._t("bn")
        self.m2[out] = gamma_scale * gamma_scale + 0.01
        self.nodes.append(Node("BatchNormalization",
                               [x, self._add_init("bn_g", g), self._add_init("bn_b", b),
                                self._add_init("bn_m", m), self._add_init("bn_v", v)],
                               [out], {"epsilon": 1e-5, "momentum": 0.9}))
        return out

    def relu(self, x: str) -> str:
        out = self._t("relu")
        # a residual sum carries a non-negative identity path, so ReLU removes little of it
        self.m2[out] = (0.9 if x.startswith("add") else 0.5) * self.m2[x]
        self.nodes.append(Node("Relu", [x], [out]))
        return out

    def prelu(self, x: str, c: int) -> str:
        slope = self.rng.uniform(0.1, 0.4, (c, 1, 1)).astype(np.float32)
        out = self._t("prelu")
        self.m2[out] = 0.535 * self.m2[x]
        self.nodes.append(Node("PRelu", [x, self._add_init("slope", slope)], [out]))
        return out

    def add(self, a: str, b: str) -> str:
        out = self._t("add")
        self.m2[out] = self.m2[a] + self.m2[b]
        self.nodes.append(Node("Add", [a, b], [out]))
        return out

    def mul_scalar(self, x: str, v: float) -> str:
        out = self._t("mul")
        self.m2[out] = self.m2[x] * v * v
        self.nodes.append(Node("Mul", [x, self._add_init("scale", np.asarray(v, np.float32))], [out]))
        return out

    def sigmoid(self, x: str) -> str:
        out = self._t("sigmoid")
        self.nodes.append(Node("Sigmoid", [x], [out]))
        return out

    def maxpool(self, x: str, k: int, s: int, p: int) -> str:
        out = self._t("maxpool")
        self.m2[out] = 2.0 * self.m2[x]
        self.nodes.append(Node("MaxPool", [x], [out], {
            "ceil_mode": 0, "kernel_shape": [k, k], "pads": [p, p, p, p], "strides": [s, s]}))
        return out

    def avgpool(self, x: str, k: int, s: int) -> str:
        out = self._t("avgpool")
        self.m2[out] = 0.7 * self.m2[x]
        self.nodes.append(Node("AveragePool", [x], [out], {
            "ceil_mode": 1, "count_include_pad": 0, "kernel_shape": [k, k],
            "pads": [0, 0, 0, 0], "strides": [s, s]}))
        return out

    def upsample2x(self, x: str) -> str:
        out = self._t("resize")
        self.m2[out] = self.m2[x]
        roi = self._add_init("roi", np.zeros((0,), np.float32))
        scales = self._add_init("scales", np.asarray([1, 1, 2, 2], np.float32))
        self.nodes.append(Node("Resize", [x, roi, scales], [out], {
            "coordinate_transformation_mode": "asymmetric", "mode": "nearest",
            "nearest_mode": "floor"}))
        return out

    def head_reshape(self, x: str, last: int) -> str:
        t = self._t("transpose")
        self.nodes.append(Node("Transpose", [x], [t], {"perm": [0, 2, 3, 1]}))
        out = self._t("reshape")
        shape = self._add_init("shape", np.asarray([-1, last], np.int64))
        self.nodes.append(Node("Reshape", [t, shape], [out]))
        return out

    def flatten(self, x: str) -> str:
        out = self._t("flatten")
        self.m2[out] = self.m2[x]
        self.nodes.append(Node("Flatten", [x], [out], {"axis": 1}))
        return out

    def gemm(self, x: str, cin: int, cout: int, gain: float = 1.0) -> str:
        w = self.rng.standard_normal((cout, cin), dtype=np.float32) * np.float32(gain / np.sqrt(cin))
        b = (0.05 * self.rng.standard_normal(cout)).astype(np.float32)
        out = self._t("gemm")
        self.m2[out] = gain * gain * self.m2[x]
        self.nodes.append(Node("Gemm", [x, self._add_init("fc_w", w), self._add_init("fc_b", b)],
                               [out], {"alpha": 1.0, "beta": 1.0, "transB": 1}))
        return out

    def finish(self, inputs: List[ValueInfo], outputs: List[ValueInfo]) -> Graph:
        return Graph(self.nodes, self.init, inputs, outputs, self.name)


# ---------------------------------------------------------------------------------------------
# SCRFD family
# ---------------------------------------------------------------------------------------------

_SCRFD_CFG = {
    # ResNetV1e-style backbones (deep stem, avg_down shortcuts) -- SURVEY.md section 7.3
    "scrfd_10g": dict(kind="resnet", stem=(28, 28, 56), planes=(56, 88, 88, 224), blocks=(3, 4, 2, 3),
                      neck=56, head_feat=80, head_stack=3, head_dw=False),
    "scrfd_2.5g": dict(kind="resnet", stem=(12, 12, 24), planes=(24, 48, 48, 80), blocks=(3, 5, 3, 2),
                       neck=24, head_feat=64, head_stack=2, head_dw=False),
    # MobileNetV1-style depthwise-separable backbone
    "scrfd_500m": dict(kind="mobilenet", planes=(16, 16, 40, 72, 152, 288), blocks=(2, 3, 2, 6),
                       neck=16, head_feat=64, head_stack=2, head_dw=True),
}

# cls-head bias: chosen so random-noise frames give a realistic number of candidates over the
# reference's default conf_thres 0.5 (hand-calibrated on seed-0 noise frames).
_SCRFD_CLS_BIAS = {"scrfd_10g": -3.3, "scrfd_2.5g": -2.0, "scrfd_500m": -1.2}


def _basic_block(g: _GB, x: str, cin: int, planes: int, stride: int) -> str:
    out = g.relu(g.bn(g.conv(x, cin, planes, 3, stride), planes))
    out = g.bn(g.conv(out, planes, planes, 3, 1), planes, gamma_scale=0.5)
    if stride != 1 or cin != planes:
        sc = x
        if stride != 1:
            sc = g.avgpool(sc, stride, stride)
        sc = g.bn(g.conv(sc, cin, planes, 1, 1, 0), planes)
    else:
        sc = x
    return g.relu(g.add(out, sc))


def _conv_dw(g: _GB, x: str, cin: int, cout: int, stride: int) -> str:
    x = g.relu(g.bn(g.conv(x, cin, cin, 3, stride, 1, groups=cin, gain=1.0), cin))
    return g.relu(g.bn(g.conv(x, cin, cout, 1, 1, 0), cout))


def _build_scrfd(key: str, seed: int, size: Tuple[int, int]) -> Graph:
    cfg = _SCRFD_CFG[key]
    g = _GB(seed, key)
    x = "input.1"
    feats: List[Tuple[str, int]] = []
    if cfg["kind"] == "resnet":
        s0, s1, s2 = cfg["stem"]
        x = g.relu(g.bn(g.conv(x, 3, s0, 3, 2), s0))
        x = g.relu(g.bn(g.conv(x, s0, s1, 3, 1), s1))
        x = g.relu(g.bn(g.conv(x, s1, s2, 3, 1), s2))
        x = g.maxpool(x, 3, 2, 1)
        cin = s2
        for si, (planes, nb) in enumerate(zip(cfg["planes"], cfg["blocks"])):
            for bi in range(nb):
                stride = 2 if (bi == 0 and si > 0) else 1
                x = _basic_block(g, x, cin, planes, stride)
                cin = planes
            feats.append((x, planes))
    else:
        pl = cfg["planes"]
        x = g.relu(g.bn(g.conv(x, 3, pl[0], 3, 2), pl[0]))
        x = _conv_dw(g, x, pl[0], pl[1], 1)
        for si, nb in enumerate(cfg["blocks"]):
            for bi in range(nb):
                if bi == 0:
                    x = _conv_dw(g, x, pl[si + 1], pl[si + 2], 2)
                else:
                    x = _conv_dw(g, x, pl[si + 2], pl[si + 2], 1)
            feats.append((x, pl[si + 2]))
    feats = feats[1:]                                   # strides 8, 16, 32
    nc = cfg["neck"]
    # PAFPN: laterals, top-down, fpn convs, bottom-up, pafpn convs (conv bias only, no norm/act)
    lat = [g.conv(f, c, nc, 1, 1, 0, bias=True, out_std=1.0) for f, c in feats]
    lat[1] = g.add(lat[1], g.upsample2x(lat[2]))
    lat[0] = g.add(lat[0], g.upsample2x(lat[1]))
    inter = [g.conv(l, nc, nc, 3, 1, 1, bias=True, out_std=1.0) for l in lat]
    for i in range(2):
        inter[i + 1] = g.add(inter[i + 1], g.conv(inter[i], nc, nc, 3, 2, 1, bias=True, out_std=0.7))
    outs = [inter[0]] + [g.conv(inter[i], nc, nc, 3, 1, 1, bias=True, out_std=1.0) for i in (1, 2)]

    hf = cfg["head_feat"]
    scores, bboxes, kpss = [], [], []
    for lvl, f in enumerate(outs):
        h = f
        cin = nc
        for _ in range(cfg["head_stack"]):
            if cfg["head_dw"]:
                h = g.relu(g.bn(g.conv(h, cin, cin, 3, 1, 1, groups=cin), cin))
                h = g.relu(g.bn(g.conv(h, cin, hf, 1, 1, 0), hf))
            else:
                h = g.relu(g.bn(g.conv(h, cin, hf, 3, 1, 1), hf))
            cin = hf
        cls_b = np.full(2, _SCRFD_CLS_BIAS[key], np.float32)
        cls = g.sigmoid(g.conv(h, hf, 2, 3, 1, 1, out_std=1.5, bias_value=cls_b))
        # distances ~ 1.5..6 stride units so decoded boxes are well-formed; kps centred on the anchor
        box_b = np.tile(np.asarray([2.5, 3.0, 2.5, 3.0], np.float32), 2)
        box = g.mul_scalar(g.conv(h, hf, 8, 3, 1, 1, out_std=0.8, bias_value=box_b), 1.0 + 0.1 * lvl)
        kps_b = np.tile(np.asarray([-1.0, -0.9, 1.0, -0.9, 0.0, 0.1, -0.8, 1.1, 0.8, 1.1], np.float32), 2)
        kps = g.conv(h, hf, 20, 3, 1, 1, out_std=0.5, bias_value=kps_b)
        scores.append(g.head_reshape(cls, 1))
        bboxes.append(g.head_reshape(box, 4))
        kpss.append(g.head_reshape(kps, 10))
    w, h = size
    out_infos = []
    for names, last in ((scores, 1), (bboxes, 4), (kpss, 10)):
        for n, s in zip(names, (8, 16, 32)):
            out_infos.append(ValueInfo(n, [(h // s) * (w // s) * 2, last]))
    return g.finish([ValueInfo("input.1", [1, 3, h, w])], out_infos)


# ---------------------------------------------------------------------------------------------
# ArcFace family
# ---------------------------------------------------------------------------------------------

def _ibasic_block(g: _GB, x: str, cin: int, planes: int, stride: int) -> str:
    out = g.bn(x, cin)
    out = g.prelu(g.bn(g.conv(out, cin, planes, 3, 1), planes), planes)
    out = g.bn(g.conv(out, planes, planes, 3, stride), planes, gamma_scale=0.45)
    sc = x
    if stride != 1 or cin != planes:
        sc = g.bn(g.conv(x, cin, planes, 1, stride, 0), planes)
    return g.add(out, sc)


def _build_iresnet50(seed: int) -> Graph:
    g = _GB(seed, "arcface_r50")
    x = g.prelu(g.bn(g.conv("input.1", 3, 64, 3, 1), 64), 64)
    cin = 64
    for planes, nb in zip((64, 128, 256, 512), (3, 4, 14, 3)):
        for bi in range(nb):
            x = _ibasic_block(g, x, cin, planes, 2 if bi == 0 else 1)
            cin = planes
    x = g.bn(x, 512)
    x = g.flatten(x)
    x = g.gemm(x, 512 * 7 * 7, 512)
    x = g.bn(x, 512)
    return g.finish([ValueInfo("input.1", ["N", 3, 112, 112])], [ValueInfo(x, ["N", 512])])


def _mbf_conv(g: _GB, x, cin, cout, k, s, p, groups=1, act=True):
    x = g.bn(g.conv(x, cin, cout, k, s, p, groups=groups), cout)
    return g.prelu(x, cout) if act else x


def _mbf_depthwise(g: _GB, x, cin, cout, groups, stride, residual):
    y = _mbf_conv(g, x, cin, groups, 1, 1, 0)
    y = _mbf_conv(g, y, groups, groups, 3, stride, 1, groups=groups)
    y = _mbf_conv(g, y, groups, cout, 1, 1, 0, act=False)
    return g.add(x, y) if residual else y


def _build_mbf(seed: int) -> Graph:
    """arcface_torch MobileFaceNet(blocks=(1,4,6,2), scale=2) -- SURVEY.md section 7.3."""
    g = _GB(seed, "arcface_mbf")
    sc = 2
    x = _mbf_conv(g, "input.1", 3, 64 * sc, 3, 2, 1)
    x = _mbf_conv(g, x, 64 * sc, 64 * sc, 3, 1, 1, groups=64 * sc)              # blocks[0] == 1
    x = _mbf_depthwise(g, x, 64 * sc, 64 * sc, 128, 2, False)
    for _ in range(4):
        x = _mbf_depthwise(g, x, 64 * sc, 64 * sc, 128, 1, True)
    x = _mbf_depthwise(g, x, 64 * sc, 128 * sc, 256, 2, False)
    for _ in range(6):
        x = _mbf_depthwise(g, x, 128 * sc, 128 * sc, 256, 1, True)
    x = _mbf_depthwise(g, x, 128 * sc, 128 * sc, 512, 2, False)
    for _ in range(2):
        x = _mbf_depthwise(g, x, 128 * sc, 128 * sc, 256, 1, True)
    x = _mbf_conv(g, x, 128 * sc, 512, 1, 1, 0)
    x = _mbf_conv(g, x, 512, 512, 7, 1, 0, groups=512, act=False)               # GDC linear 7x7 dw
    x = g.flatten(x)
    x = g.gemm(x, 512, 512)
    x = g.bn(x, 512)
    return g.finish([ValueInfo("input.1", ["N", 3, 112, 112])], [ValueInfo(x, ["N", 512])])


_SEEDS = {"scrfd_500m": 1231, "scrfd_2.5g": 1232, "scrfd_10g": 1234, "arcface_mbf": 1235, "arcface_r50": 1236}


def build_arch(key: str, seed: Optional[int] = None, det_size: Tuple[int, int] = (640, 640)) -> Graph:
    """Return an ONNX Graph for `key` with deterministic random initialisers."""
    if seed is None:
        seed = _SEEDS[key]
    if key in _SCRFD_CFG:
        return _build_scrfd(key, seed, det_size)
    if key == "arcface_r50":
        return _build_iresnet50(seed)
    if key == "arcface_mbf":
        return _build_mbf(seed)
    raise KeyError(key)


def count_macs(graph: Graph, input_shape: Sequence[int]) -> int:
    """Multiply-accumulates of Conv/Gemm nodes for one NCHW input (architecture sanity check)."""
    shapes = {graph.real_inputs()[0].name: tuple(input_shape)}
    macs = 0
    for n in graph.nodes:
        a = n.attrs
        if n.op_type == "Conv":
            _, cin, h, w = shapes[n.inputs[0]]
            wt = graph.initializers[n.inputs[1]]
            cout, cpg, kh, kw = wt.shape
            p, s = a["pads"], a["strides"]
            ho = (h + p[0] + p[2] - kh) // s[0] + 1
            wo = (w + p[1] + p[3] - kw) // s[1] + 1
            macs += cout * cpg * kh * kw * ho * wo
            shapes[n.outputs[0]] = (1, cout, ho, wo)
        elif n.op_type in ("MaxPool", "AveragePool"):
            _, c, h, w = shapes[n.inputs[0]]
            k, s, p = a["kernel_shape"], a["strides"], a["pads"]
            if a.get("ceil_mode", 0):
                ho = -(-(h + p[0] + p[2] - k[0]) // s[0]) + 1
                wo = -(-(w + p[1] + p[3] - k[1]) // s[1]) + 1
            else:
                ho = (h + p[0] + p[2] - k[0]) // s[0] + 1
                wo = (w + p[1] + p[3] - k[1]) // s[1] + 1
            shapes[n.outputs[0]] = (1, c, ho, wo)
        elif n.op_type == "Resize":
            _, c, h, w = shapes[n.inputs[0]]
            shapes[n.outputs[0]] = (1, c, h * 2, w * 2)
        elif n.op_type == "Gemm":
            wt = graph.initializers[n.inputs[1]]
            macs += wt.shape[0] * wt.shape[1]
            shapes[n.outputs[0]] = (1, wt.shape[0])
        elif n.op_type == "Flatten":
            s = shapes[n.inputs[0]]
            shapes[n.outputs[0]] = (1, int(np.prod(s[1:])))
        elif n.op_type in ("Transpose", "Reshape"):
            shapes[n.outputs[0]] = shapes[n.inputs[0]]
        else:
            shapes[n.outputs[0]] = shapes[n.inputs[0]]
    return macs
