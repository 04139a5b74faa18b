"""ONNX graph -> fused layer plan for the sm_100a kernels (host side, numpy only).

The reference hands its .onnx files to onnxruntime (reference models/scrfd.py:59-62,83;
models/arcface.py:18-21,51).  Here the same graph is compiled into a short list of fused ops:

  * every Conv absorbs a preceding BatchNormalization (scale folded into the weights, shift turned
    into a per-border-class bias table so zero padding stays exact), the following
    BatchNormalization / Mul-by-constant, one residual Add (optionally of a nearest-2x Resize) and
    the activation (Relu / PRelu / Sigmoid);
  * Flatten + Gemm become a kxk "valid" convolution over the kxk map (same tensor-core kernel);
  * Transpose(0,2,3,1) + Reshape(-1,k) on the detector heads are no-ops in NHWC;
  * whatever is left (pools, stray BN / Add / activations) maps to the small CUDA-core kernels.

Nothing here touches the GPU: `compile_graph` returns numpy weights in kernel layout, so the
folding arithmetic is unit-tested on the CPU box (tests/test_graph_compile.py).
"""
from __future__ import annotations

import os

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np

from .onnx_wire import Graph, Node

ACT_NONE, ACT_RELU, ACT_PRELU, ACT_SIGMOID = 0, 1, 2, 3
_ACT = {"Relu": ACT_RELU, "PRelu": ACT_PRELU, "Sigmoid": ACT_SIGMOID}


def pad_ch(c: int) -> int:
    """Stored channel count: 16 for tiny tensors, else a multiple of 32 so the K pipeline never drops to
    16-element (32-byte) chunks -- 28->32, 56->64, 80->96, 88->96, 224 stays."""
    return 16 if c <= 16 else (c + 31) // 32 * 32


pad16 = pad_ch      # historical name


@dataclass
class TensorSpec:
    name: str
    c: int
    h: int
    w: int
    cp: int
    f32: bool = False


@dataclass
class FusedOp:
    kind: str                       # stem | conv | dwconv | pool | eltwise
    src: str
    dst: str
    residual: Optional[str] = None
    res_mode: int = 0
    act: int = ACT_NONE
    attrs: Dict[str, int] = field(default_factory=dict)
    arrays: Dict[str, np.ndarray] = field(default_factory=dict)   # kernel-layout host arrays
    order: int = 0
    sc_src: Optional[str] = None    # fused projection shortcut: second input tensor (arrays['sc_weight'], attrs sc_*)


@dataclass
class Plan:
    input_name: str
    in_hw: Tuple[int, int]
    ops: List[FusedOp]
    tensors: Dict[str, TensorSpec]
    outputs: List[Tuple[str, str, int, int]]  # (graph output name, tensor name, channels used, channel offset)

    def conv_flops(self, n: int = 1) -> int:
        total = 0
        for op in self.ops:
            if op.kind in ("conv", "stem", "dwconv"):
                total += 2 * op.attrs["macs_per_image"] * n
        return total


def _bn_affine(g: Graph, node: Node) -> Tuple[np.ndarray, np.ndarray]:
    gamma, beta, mean, var = (g.initializers[k].astype(np.float64) for k in node.inputs[1:5])
    s = gamma / np.sqrt(var + float(node.attrs.get("epsilon", 1e-5)))
    return s, beta - mean * s


def _infer_one(init: Dict[str, np.ndarray], n: Node, shapes: Dict[str, Tuple[int, ...]]) -> None:
    """(C, H, W) of node n's output from its inputs' shapes (batch dimension implicit)."""
    a = n.attrs
    t = n.op_type
    if t == "Conv":
        c, h, w = shapes[n.inputs[0]]
        wt = init[n.inputs[1]]
        p, s = a.get("pads", [0, 0, 0, 0]), a.get("strides", [1, 1])
        shapes[n.outputs[0]] = (wt.shape[0], (h + p[0] + p[2] - wt.shape[2]) // s[0] + 1,
                                (w + p[1] + p[3] - wt.shape[3]) // s[1] + 1)
    elif t in ("MaxPool", "AveragePool"):
        c, h, w = shapes[n.inputs[0]]
        k, s, p = a["kernel_shape"], a.get("strides", [1, 1]), a.get("pads", [0, 0, 0, 0])
        if a.get("ceil_mode", 0):
            ho = -(-(h + p[0] + p[2] - k[0]) // s[0]) + 1
            wo = -(-(w + p[1] + p[3] - k[1]) // s[1]) + 1
        else:
            ho = (h + p[0] + p[2] - k[0]) // s[0] + 1
            wo = (w + p[1] + p[3] - k[1]) // s[1] + 1
        shapes[n.outputs[0]] = (c, ho, wo)
    elif t in ("Resize", "Upsample"):
        c, h, w = shapes[n.inputs[0]]
        sc = None
        for nm in n.inputs[1:]:
            if nm and nm in init and init[nm].size == 4:
                sc = init[nm]
        if sc is None:
            raise NotImplementedError("Resize whose scales / sizes do not fold to constants is not supported")
        if sc.dtype.kind == "f":
            shapes[n.outputs[0]] = (c, int(h * float(sc.reshape(-1)[2])), int(w * float(sc.reshape(-1)[3])))
        else:
            shapes[n.outputs[0]] = (c, int(sc.reshape(-1)[2]), int(sc.reshape(-1)[3]))
    elif t == "Flatten":
        shapes[n.outputs[0]] = (int(np.prod(shapes[n.inputs[0]])),)
    elif t == "Gemm":
        wt = init[n.inputs[1]]
        shapes[n.outputs[0]] = (wt.shape[0] if a.get("transB", 0) else wt.shape[1],)
    elif t in ("Transpose", "Reshape"):
        shapes[n.outputs[0]] = shapes[n.inputs[0]]
    else:
        src = [i for i in n.inputs if i in shapes]
        if not src:
            raise NotImplementedError(f"cannot infer shape for {t}")
        shapes[n.outputs[0]] = shapes[src[0]]


def _infer_shapes(g: Graph, in_hw: Tuple[int, int]) -> Dict[str, Tuple[int, ...]]:
    inp = g.real_inputs()[0]
    shapes: Dict[str, Tuple[int, ...]] = {inp.name: (3, in_hw[0], in_hw[1])}
    for n in g.nodes:
        _infer_one(g.initializers, n, shapes)
    return shapes


def compile_graph(g: Graph, in_hw: Tuple[int, int], stem_im2col: bool = True, merge_heads: bool = True,
                  fuse_shortcuts: bool = True) -> Plan:
    # exporter glue first: Constant nodes and the Shape / Gather / Concat arithmetic around Resize and Reshape fold to
    # initializers for the static input size (onnx_fold.py); what is left is the network
    from .onnx_fold import fold_shape_glue
    g = fold_shape_glue(g, in_hw)
    nodes = g.nodes
    init = g.initializers
    shapes = _infer_shapes(g, in_hw)
    prod: Dict[str, int] = {}
    cons: Dict[str, List[int]] = {}
    for i, n in enumerate(nodes):
        for o in n.outputs:
            prod[o] = i
        for x in n.inputs:
            if x and x not in init:
                cons.setdefault(x, []).append(i)
    graph_out = {o.name for o in g.outputs}

    def single(name: str) -> bool:
        return len(cons.get(name, [])) == 1 and name not in graph_out

    def const_operand(node: Node, other_than: str) -> Optional[np.ndarray]:
        for x in node.inputs:
            if x != other_than and x in init:
                return init[x].astype(np.float64)
        return None

    def conv_can_residual(i: int) -> bool:
        n = nodes[i]
        wt = init[n.inputs[1]]
        return n.op_type == "Gemm" or (n.attrs.get("group", 1) == 1 and wt.shape[1] > 4)

    def root_conv_of(name: str) -> Optional[int]:
        """conv whose pre-Add chain ends at `name` (walk back through single-consumer BN / Mul-const)."""
        while name in prod and single(name):
            i = prod[name]
            n = nodes[i]
            if n.op_type in ("Conv", "Gemm"):
                return i if conv_can_residual(i) else None
            if n.op_type == "BatchNormalization" or (n.op_type == "Mul" and const_operand(n, n.inputs[0]) is not None):
                name = n.inputs[0]
                continue
            return None
        return None

    absorbed: set = set()
    alias: Dict[str, str] = {}           # Flatten / Transpose / Reshape outputs -> underlying tensor
    ops: List[FusedOp] = []

    def resolve(name: str) -> str:
        while name in alias:
            name = alias[name]
        return name

    def follow_chain(idx: int, cout: int, bias: np.ndarray, allow_add: bool):
        """absorb BN / Mul / Add / activation after node idx.  Returns (scale, shift, residual, res_mode, act, slope, out)."""
        cur = nodes[idx].outputs[0]
        scale = np.ones(cout, np.float64)
        shift = bias.astype(np.float64).copy()
        residual, res_mode, act, slope = None, 0, ACT_NONE, None
        stage = 0
        while single(cur):
            ni = cons[cur][0]
            nx = nodes[ni]
            t = nx.op_type
            if t == "BatchNormalization" and stage == 0:
                s, sh = _bn_affine(g, nx)
                scale, shift = scale * s, shift * s + sh
            elif t == "Mul" and stage == 0 and const_operand(nx, cur) is not None:
                m = const_operand(nx, cur).reshape(-1)
                if m.size not in (1, cout):
                    break
                scale, shift = scale * m, shift * m
            elif t == "Add" and stage == 0 and const_operand(nx, cur) is not None:
                m = const_operand(nx, cur).reshape(-1)
                if m.size not in (1, cout):
                    break
                shift = shift + m
            elif t == "Add" and stage == 0 and allow_add:
                other = [x for x in nx.inputs if x != cur]
                if len(other) != 1:
                    break
                other = other[0]
                root = root_conv_of(other)
                if root is not None and root > idx:
                    break                                   # the later convolution absorbs this Add
                on = nodes[prod[other]] if other in prod else None
                if on is not None and on.op_type in ("Resize", "Upsample") and single(other):
                    ci, hi, wi = shapes[on.inputs[0]]
                    co, ho, wo = shapes[other]
                    if on.attrs.get("mode", "nearest") != "nearest" or (ho, wo) != (2 * hi, 2 * wi):
                        break
                    residual, res_mode = on.inputs[0], 2
                    absorbed.add(prod[other])
                else:
                    residual, res_mode = other, 1
                stage = 1
            elif t in _ACT and stage <= 1:
                act = _ACT[t]
                if t == "PRelu":
                    sl = init[nx.inputs[1]].astype(np.float64).reshape(-1)
                    slope = np.broadcast_to(sl, (cout,)).copy() if sl.size in (1, cout) else None
                    if slope is None:
                        break
                stage = 2
            else:
                break
            absorbed.add(ni)
            cur = nx.outputs[0]
            if stage == 2:
                break
        return scale, shift, residual, res_mode, act, slope, cur

    for idx, n in enumerate(nodes):
        t = n.op_type
        if t in ("Flatten", "Transpose", "Reshape"):
            if t == "Transpose" and list(n.attrs.get("perm", [])) != [0, 2, 3, 1]:
                raise NotImplementedError("only the NCHW->NHWC head transpose is supported")
            alias[n.outputs[0]] = n.inputs[0]
            absorbed.add(idx)
            continue
        if t not in ("Conv", "Gemm"):
            continue
        a = n.attrs
        x = n.inputs[0]
        if t == "Conv":
            W = init[n.inputs[1]].astype(np.float64)
            group = a.get("group", 1)
            pads, strides = a.get("pads", [0, 0, 0, 0]), a.get("strides", [1, 1])
            if len(set(pads)) != 1 or strides[0] != strides[1] or list(a.get("dilations", [1, 1])) != [1, 1]:
                raise NotImplementedError(f"conv {n.name}: asymmetric pads/strides or dilation unsupported")
            pad, stride = pads[0], strides[0]
            cin, hi, wi = shapes[resolve(x)] if resolve(x) in shapes else shapes[x]
        else:
            if not a.get("transB", 0) or a.get("transA", 0) or a.get("alpha", 1.0) != 1.0 or a.get("beta", 1.0) != 1.0:
                raise NotImplementedError("Gemm must be y = x @ W.T + b")
            base = resolve(x)
            cin, hi, wi = shapes[base]
            W = init[n.inputs[1]].astype(np.float64).reshape(-1, cin, hi, wi)
            group, pad, stride = 1, 0, 1
        cout, cpg, kh, kw = W.shape
        b = init[n.inputs[2]].astype(np.float64) if len(n.inputs) > 2 and n.inputs[2] else np.zeros(cout)
        depthwise = group > 1
        if depthwise and not (group == cin == cout and cpg == 1):
            raise NotImplementedError("grouped convolution other than depthwise is unsupported")
        stem = (not depthwise) and cin <= 4
        if stem and not (kh == kw == 3 and pad == 1):
            raise NotImplementedError("first-layer convolution must be 3x3 pad 1")

        # ---- preceding BatchNormalization (input affine) ----------------------------------------
        src = resolve(x)
        pre = None
        if src in prod and not depthwise and not stem:
            pn = nodes[prod[src]]
            if pn.op_type == "BatchNormalization" and prod[src] not in absorbed \
                    and len(cons.get(src, [])) == 1 and src not in graph_out \
                    and (pad == 0 or (kh <= 3 and kw <= 3 and pad == 1 and hi >= 2 and wi >= 2)):
                pre = _bn_affine(g, pn)
                absorbed.add(prod[src])
                src = resolve(pn.inputs[0])
        scale, shift, residual, res_mode, act, slope, out = follow_chain(
            idx, cout, b, allow_add=(not depthwise and not stem))
        absorbed.add(idx)

        ho = (hi + 2 * pad - kh) // stride + 1
        wo = (wi + 2 * pad - kw) // stride + 1
        attrs = dict(cin=cin, cout=cout, kh=kh, kw=kw, stride=stride, pad=pad, h=hi, w=wi, ho=ho, wo=wo,
                     macs_per_image=cout * cpg * kh * kw * ho * wo)
        arrays: Dict[str, np.ndarray] = {}
        cout_p = pad16(cout)
        if slope is not None:
            sl = np.zeros(cout_p, np.float32)
            sl[:cout] = slope
            arrays["slope"] = sl
        if depthwise:
            Wf = W[:, 0] * scale[:, None, None]                      # (C, kh, kw)
            wk = np.zeros((kh * kw, cout_p), np.float32)
            wk[:, :cout] = Wf.reshape(cout, kh * kw).T
            bias = np.zeros(cout_p, np.float32)
            bias[:cout] = shift
            arrays.update(weight=wk, bias=bias)
            kind = "dwconv"
        elif stem:
            Wf = W * scale[:, None, None, None]                      # (Cout, Cin, 3, 3)
            wk = np.zeros((9, 4, cout_p), np.float32)
            wk[:, :cin, :cout] = Wf.transpose(2, 3, 1, 0).reshape(9, cin, cout)
            bias = np.zeros(cout_p, np.float32)
            bias[:cout] = shift
            arrays.update(weight=wk, bias=bias)
            kind = "stem"
        else:
            cin_p = pad16(cin)
            classes = 1
            if pre is not None:
                s_in, t_in = pre
                T = np.einsum("oirs,i->rso", W, t_in)                # (kh, kw, Cout): shift seen through each tap
                W = W * s_in[None, :, None, None]
                if pad > 0:
                    classes = 9
                    table = np.zeros((9, cout), np.float64)
                    for cy in range(3):
                        rs = [r for r in range(kh) if not ((cy == 0 and r < pad) or (cy == 2 and r >= kh - pad))]
                        for cx in range(3):
                            ss = [s for s in range(kw) if not ((cx == 0 and s < pad) or (cx == 2 and s >= kw - pad))]
                            table[cy * 3 + cx] = T[np.ix_(rs, ss)].sum(axis=(0, 1))
                    bias_tab = shift[None, :] + scale[None, :] * table
                else:
                    bias_tab = (shift + scale * T.sum(axis=(0, 1)))[None, :]
            else:
                bias_tab = shift[None, :]
            Wf = W * scale[:, None, None, None]
            wk = np.zeros((kh * kw, cout_p, cin_p), np.float32)
            wk[:, :cout, :cin] = Wf.transpose(2, 3, 0, 1).reshape(kh * kw, cout, cin)
            bias = np.zeros((classes, cout_p), np.float32)
            bias[:, :cout] = bias_tab
            arrays.update(weight=wk, bias=bias)
            attrs["bias_classes"] = classes
            kind = "conv"
        ops.append(FusedOp(kind, src, out, resolve(residual) if residual else None, res_mode, act, attrs, arrays, idx))

    # ---- everything no convolution absorbed ------------------------------------------------------
    for idx, n in enumerate(nodes):
        if idx in absorbed:
            continue
        t, a = n.op_type, n.attrs
        src = resolve(n.inputs[0])
        c, h, w = shapes[src]
        if t in ("MaxPool", "AveragePool"):
            k, s, p = a["kernel_shape"], a.get("strides", [1, 1]), a.get("pads", [0, 0, 0, 0])
            if k[0] != k[1] or s[0] != s[1] or len(set(p)) != 1:
                raise NotImplementedError("pooling must be square and symmetric")
            if t == "AveragePool" and a.get("count_include_pad", 0) and p[0] > 0:
                raise NotImplementedError("AveragePool with count_include_pad=1 and padding is unsupported")
            co, ho, wo = shapes[n.outputs[0]]
            ops.append(FusedOp("pool", src, n.outputs[0], attrs=dict(
                k=k[0], stride=s[0], pad=p[0], mode=0 if t == "MaxPool" else 1, h=h, w=w, ho=ho, wo=wo, c=c), order=idx))
        elif t == "BatchNormalization":
            s, sh = _bn_affine(g, n)
            cp = pad16(c)
            sc = np.zeros(cp, np.float32)
            sf = np.zeros(cp, np.float32)
            sc[:c], sf[:c] = s, sh
            ops.append(FusedOp("eltwise", src, n.outputs[0], arrays=dict(scale=sc, shift=sf),
                               attrs=dict(h=h, w=w, c=c), order=idx))
        elif t in _ACT:
            arrays = {}
            if t == "PRelu":
                sl = np.zeros(pad16(c), np.float32)
                sl[:c] = np.broadcast_to(init[n.inputs[1]].reshape(-1), (c,))
                arrays["slope"] = sl
            ops.append(FusedOp("eltwise", src, n.outputs[0], act=_ACT[t], arrays=arrays,
                               attrs=dict(h=h, w=w, c=c), order=idx))
        elif t == "Add":
            ops.append(FusedOp("eltwise", src, n.outputs[0], residual=resolve(n.inputs[1]), res_mode=1,
                               attrs=dict(h=h, w=w, c=c), order=idx))
        else:
            raise NotImplementedError(f"ONNX op {t} ({n.name}) is not supported by the B200 engine")

    out_names = {resolve(o.name) for o in g.outputs}
    view: Dict[str, Tuple[str, int]] = {}          # output tensor -> (merged tensor, channel offset)
    extra_specs: Dict[str, TensorSpec] = {}

    # ---- first layer on the tensor cores: 3x3 patches -> 27(+5) channels, then a 1x1 convolution -------------
    if stem_im2col:
        new_ops: List[FusedOp] = []
        for op in ops:
            if op.kind != "stem":
                new_ops.append(op)
                continue
            a = op.attrs
            mid = op.dst + "__patches"
            new_ops.append(FusedOp("im2col", op.src, mid, attrs=dict(h=a["h"], w=a["w"], ho=a["ho"], wo=a["wo"],
                                                                     stride=a["stride"]), order=op.order))
            extra_specs[mid] = TensorSpec(mid, 27, a["ho"], a["wo"], 32)
            cout, cout_p = a["cout"], pad_ch(a["cout"])
            wk = op.arrays["weight"]                                    # (9, 4, cout_p) fp32
            w2 = np.zeros((1, cout_p, 32), np.float32)
            w2[0, :, :27] = wk[:, :3, :].reshape(27, cout_p).T           # k = tap*3 + ci
            arrays = dict(weight=w2, bias=op.arrays["bias"][None, :].copy())
            if "slope" in op.arrays:
                arrays["slope"] = op.arrays["slope"]
            attrs = dict(cin=27, cout=cout, kh=1, kw=1, stride=1, pad=0, h=a["ho"], w=a["wo"], ho=a["ho"], wo=a["wo"],
                         macs_per_image=a["macs_per_image"], bias_classes=1)
            new_ops.append(FusedOp("conv", mid, op.dst, None, 0, op.act, attrs, arrays, op.order))
        ops = new_ops

    # ---- sibling head convolutions (same input, same geometry, all graph outputs) become one launch ----------
    if merge_heads:
        groups: Dict[Tuple, List[FusedOp]] = {}
        for op in ops:
            a = op.attrs
            if op.kind == "conv" and op.dst in out_names and op.residual is None and a["bias_classes"] == 1 \
                    and op.act in (ACT_NONE, ACT_SIGMOID):
                groups.setdefault((op.src, a["kh"], a["kw"], a["stride"], a["pad"]), []).append(op)
        for key, members in groups.items():
            if len(members) < 2:
                continue
            members.sort(key=lambda o: (o.act != ACT_SIGMOID, o.order))   # sigmoid channels first
            couts = [m.attrs["cout"] for m in members]
            total = sum(couts)
            cin = members[0].attrs["cin"]
            cin_p = members[0].arrays["weight"].shape[2]
            taps = members[0].arrays["weight"].shape[0]
            wk = np.zeros((taps, pad_ch(total), cin_p), np.float32)
            bias = np.zeros((1, pad_ch(total)), np.float32)
            name = "+".join(m.dst for m in members)
            off = 0
            sig_hi = 0
            for m_op, c in zip(members, couts):
                wk[:, off:off + c, :] = m_op.arrays["weight"][:, :c, :]
                bias[0, off:off + c] = m_op.arrays["bias"][0, :c]
                view[m_op.dst] = (name, off)
                if m_op.act == ACT_SIGMOID:
                    sig_hi = off + c
                off += c
            a0 = dict(members[0].attrs)
            a0.update(cout=total, macs_per_image=sum(m.attrs["macs_per_image"] for m in members), sig_hi=sig_hi)
            merged = FusedOp("conv", members[0].src, name, None, 0, ACT_SIGMOID if sig_hi else ACT_NONE, a0,
                             dict(weight=wk, bias=bias), min(m.order for m in members))
            ops = [o for o in ops if o not in members] + [merged]

    # ---- projection shortcuts ride as extra K of the convolution they are added to ----------------------------
    # ResNet down-sampling block: out = conv_kxk(y) + conv_1x1_stride_s(x).  The 1x1 launch, its output tensor and the
    # residual read disappear: the kernel appends the shortcut's K chunks (activation boxes from x) after the taps.
    if fuse_shortcuts:
        def kchunk_of(cp: int) -> int:
            return 64 if cp % 64 == 0 else (32 if cp % 32 == 0 else 16)
        readers: Dict[str, int] = {}
        for op in ops:
            for t in (op.src, op.residual):
                if t:
                    readers[t] = readers.get(t, 0) + 1
        by_dst = {op.dst: op for op in ops}

        def is_projection(o: FusedOp) -> bool:
            b = o.attrs
            return o.kind == "conv" and b.get("kh") == 1 and b.get("kw") == 1 and b.get("pad") == 0 \
                and b.get("bias_classes") == 1 and o.sc_src is None

        for op in list(ops):
            if op not in ops or op.kind != "conv" or op.res_mode != 1 or not op.residual or op.residual not in by_dst:
                continue
            other = by_dst[op.residual]
            if other is op or other.kind != "conv" or other.residual or other.act != ACT_NONE or other.sc_src \
                    or readers.get(other.dst, 0) != 1 or other.dst in out_names:
                continue
            a, b = op.attrs, other.attrs
            if (b["ho"], b["wo"]) != (a["ho"], a["wo"]) or b["cout"] != a["cout"]:
                continue
            # the Add may have been absorbed by either convolution: `main` keeps its taps, `proj` becomes extra K
            if is_projection(other) and not (is_projection(op) and a["cin"] < b["cin"]):
                main, proj = op, other
            elif is_projection(op):
                main, proj = other, op
            else:
                continue
            ma = main.attrs
            if ma["kh"] == 3 and ma["kw"] == 3 and ma["stride"] == 1 and ma["pad"] == 1:
                continue                            # stride-1 3x3 layers run on halo boxes, which extra per-tap chunks would forfeit
            cin_p, sc_cin_p = main.arrays["weight"].shape[2], proj.arrays["weight"].shape[2]
            if sc_cin_p % kchunk_of(cin_p) != 0:
                continue
            pa = proj.attrs
            main.arrays["sc_weight"] = proj.arrays["weight"]
            main.arrays["bias"] = main.arrays["bias"] + proj.arrays["bias"][0][None, :]
            main.attrs.update(sc_cin=pa["cin"], sc_stride=pa["stride"], sc_h=pa["h"], sc_w=pa["w"],
                              macs_per_image=main.attrs["macs_per_image"] + pa["macs_per_image"])
            main.sc_src = proj.src
            if main is other:                       # the fused op takes over the Add's output, activation and position
                main.dst, main.act, main.order = op.dst, op.act, max(op.order, other.order)
                if "slope" in op.arrays:
                    main.arrays["slope"] = op.arrays["slope"]
                by_dst[main.dst] = main
            main.residual, main.res_mode = None, 0
            ops.remove(proj)

    # ---- fuse a 3x3 / stride 2 / pad 1 MaxPool into the ReLU convolution that feeds it -----------------------------
    # (the SCRFD stem: Conv-Relu-MaxPool).  `b2f_conv2d` then max-reduces the tiles' window maxima straight into the
    # pooled map (include/b2f.h, `pool`), so the full-resolution conv output -- the largest tensor of the detector --
    # never reaches HBM and the separate pooling pass disappears.  B2F_FUSE_POOL=0 keeps the two ops.
    if os.environ.get("B2F_FUSE_POOL", "1") != "0":
        readers = {}
        for op in ops:
            for t in (op.src, op.residual, op.sc_src):
                if t:
                    readers[t] = readers.get(t, 0) + 1
        by_dst = {op.dst: op for op in ops}
        for pool in [o for o in ops if o.kind == "pool"]:
            pa = pool.attrs
            conv = by_dst.get(pool.src)
            if (pa["mode"] != 0 or pa["k"] != 3 or pa["stride"] != 2 or pa["pad"] != 1 or conv is None or conv.kind != "conv"
                    or conv.act != ACT_RELU or conv.residual or conv.sc_src or readers.get(conv.dst, 0) != 1
                    or conv.dst in out_names or pool.dst in out_names):
                continue
            ca = conv.attrs
            cout_p = conv.arrays["weight"].shape[1]
            if (ca["kh"], ca["kw"], ca["stride"], ca["pad"]) != (3, 3, 1, 1) or cout_p % 32 != 0 or cout_p > 128:
                continue
            if (pa["ho"], pa["wo"]) != ((ca["ho"] - 1) // 2 + 1, (ca["wo"] - 1) // 2 + 1):
                continue
            conv.attrs.update(pool=1, pool_ho=pa["ho"], pool_wo=pa["wo"])
            conv.dst, conv.order = pool.dst, max(conv.order, pool.order)
            ops.remove(pool)

    # ---- a 2x2 / stride 2 AveragePool in front of a 1x1 convolution IS a 2x2 / stride 2 convolution ----------------
    # (the ResNet-D style down-sampling shortcut of the SCRFD backbones: AvgPool -> Conv1x1 -> BN, + main branch, ReLU).
    # conv1x1(avgpool(x)) = sum over the four taps of x_tap * (W / 4): the pooling pass, its launch and the pooled tensor
    # disappear, and the pooled value is no longer rounded to 16 bits before the convolution.  Only for even maps (no
    # partial windows, so ceil_mode / count_include_pad do not matter).  W / 4 is exact in fp16 / bf16 (a power of two).
    # `macs_per_image` stays the algorithmic count of the 1x1 layer.  B2F_FOLD_AVGPOOL=0 keeps the two ops.
    if os.environ.get("B2F_FOLD_AVGPOOL", "1") != "0":
        readers = {}
        for op in ops:
            for t in (op.src, op.residual, op.sc_src):
                if t:
                    readers[t] = readers.get(t, 0) + 1
        by_dst = {op.dst: op for op in ops}
        for conv in [o for o in ops if o.kind == "conv"]:
            pool = by_dst.get(conv.src)
            ca = conv.attrs
            if (pool is None or pool.kind != "pool" or pool.attrs["mode"] != 1 or pool.attrs["k"] != 2
                    or pool.attrs["stride"] != 2 or pool.attrs["pad"] != 0 or pool.attrs["h"] % 2 or pool.attrs["w"] % 2
                    or readers.get(pool.dst, 0) != 1 or pool.dst in out_names or conv.sc_src
                    or (ca["kh"], ca["kw"], ca["stride"], ca["pad"]) != (1, 1, 1, 0) or ca["bias_classes"] != 1
                    or conv.residual == pool.dst):
                continue
            w1 = conv.arrays["weight"]                                   # [1][cout_p][cin_p]
            conv.arrays["weight"] = np.ascontiguousarray(np.repeat(w1 * np.asarray(0.25, w1.dtype), 4, axis=0))
            conv.attrs.update(kh=2, kw=2, stride=2, pad=0, h=pool.attrs["h"], w=pool.attrs["w"])
            conv.src = pool.src
            ops.remove(pool)

    # ---- topological order over fused ops ------------------------------------------------------------
    inp = g.real_inputs()[0].name
    produced_by = {op.dst: op for op in ops}
    done = {inp}
    ordered: List[FusedOp] = []
    pending = sorted(ops, key=lambda o: o.order)
    while pending:
        progressed = False
        for op in list(pending):
            deps = [op.src] + ([op.residual] if op.residual else []) + ([op.sc_src] if op.sc_src else [])
            if all(d in done for d in deps):
                ordered.append(op)
                done.add(op.dst)
                pending.remove(op)
                progressed = True
        if not progressed:
            missing = {d for op in pending for d in [op.src, op.residual, op.sc_src] if d and d not in done and d not in produced_by}
            raise RuntimeError(f"graph compile: unresolved tensors {sorted(missing)[:5]}")

    # ---- tensor table ----------------------------------------------------------------------------------
    out_tensors = {view.get(t, (t, 0))[0] for t in out_names}
    tensors: Dict[str, TensorSpec] = {inp: TensorSpec(inp, 3, in_hw[0], in_hw[1], 4)}
    for op in ordered:
        if op.dst in extra_specs:
            tensors[op.dst] = extra_specs[op.dst]
            continue
        a = op.attrs
        if op.kind in ("conv", "dwconv", "stem"):
            c, h, w = a["cout"], a["ho"], a["wo"]
            if a.get("pool"):
                h, w = a["pool_ho"], a["pool_wo"]
        elif op.kind == "pool":
            c, h, w = a["c"], a["ho"], a["wo"]
        else:
            c, h, w = a["c"], a["h"], a["w"]
        tensors[op.dst] = TensorSpec(op.dst, c, h, w, pad_ch(c), f32=op.dst in out_tensors)
        if op.dst in out_tensors and op.kind != "conv":
            raise NotImplementedError("graph outputs must be produced by a tensor-core convolution / Gemm")
    outputs = []
    for o in g.outputs:
        tname = resolve(o.name)
        shp = shapes[tname]
        base, off = view.get(tname, (tname, 0))
        outputs.append((o.name, base, shp[0], off))
    return Plan(inp, in_hw, ordered, tensors, outputs)
