"""Executes a compiled Plan on one B200 through the C-ABI (ctypes) -- PyTorch only owns memory/streams.

Replaces `onnxruntime.InferenceSession.run` at reference models/scrfd.py:83 and
models/arcface.py:51.  Activations are NHWC fp16 (or bf16) with channels padded to 16; detector
heads and the embedding come back as fp32.  There is no CPU path: constructing a NetEngine without
CUDA raises.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

from . import _lib
from .graph import Plan, FusedOp

F16, BF16, F32 = 0, 1, 2


def default_dtype() -> int:
    return BF16 if os.environ.get("B2F_DTYPE", "f16").lower() in ("bf16", "bfloat16") else F16


def torch_dtype(code: int) -> torch.dtype:
    return torch.bfloat16 if code == BF16 else torch.float16


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


class _Bound:
    """One op bound to concrete buffers for a given batch size: a ready-to-call closure."""
    __slots__ = ("fn", "args", "keep")

    def __init__(self, fn, args, keep):
        self.fn, self.args, self.keep = fn, args, keep


def capacity_for(n: int) -> int:
    """Activation buffers are sized for the next power of two >= n, so a service that sees every face count 1..K binds
    log2(K) buffer sets (at most 2x the largest) instead of K of them."""
    cap = 1
    while cap < n:
        cap *= 2
    return cap


MAX_BOUND = 64      # per-batch-size launch lists kept (descriptors only, a few KB each); least recently used goes first


def stem8_weights(w: torch.Tensor) -> torch.Tensor:
    """[1][cout_p][32] weights of the first convolution as a 1x1 over 3x3 patches (k = tap * 3 + channel, 27 used) ->
    [10][cout_p][8] for `b2f_conv2d`'s stem form: slot = filter tap (ky * 3 + kx), 8 stored channels per pixel (3 used),
    slot 9 zero (the second half of the last K = 16 step)."""
    cout_p = w.shape[1]
    w8 = torch.zeros((10, cout_p, 8), dtype=w.dtype, device=w.device)
    w8[:9, :, :3] = w[0, :, :27].reshape(cout_p, 9, 3).permute(1, 0, 2)
    return w8.contiguous()


class NetEngine:
    def __init__(self, plan: Plan, device: Optional[torch.device] = None, dtype: Optional[int] = None):
        if not torch.cuda.is_available():
            raise _lib.B2FError("NetEngine needs a CUDA device: there is no CPU fallback")
        self.plan = plan
        self.device = device or torch.device("cuda", torch.cuda.current_device())
        self.dtype = default_dtype() if dtype is None else dtype
        self.lib = _lib.lib()
        self._weights: List[Dict[str, torch.Tensor]] = []
        tdt = torch_dtype(self.dtype)
        for op in plan.ops:
            dev: Dict[str, torch.Tensor] = {}
            for k, arr in op.arrays.items():
                t = torch.from_numpy(np.ascontiguousarray(arr))
                if op.kind == "conv" and k in ("weight", "sc_weight"):
                    t = t.to(tdt)
                dev[k] = t.to(self.device).contiguous()
            self._weights.append(dev)
        # launch lists per exact batch size n (descriptors + views) over buffer pools per capacity (capacity_for(n))
        self._bound: Dict[int, Tuple[List[_Bound], Dict[str, torch.Tensor], List[torch.Tensor]]] = {}
        self._pools: Dict[int, List[torch.Tensor]] = {}
        # liveness: last op index that reads each tensor
        self._last_use: Dict[str, int] = {}
        for i, op in enumerate(plan.ops):
            self._last_use[op.src] = i
            if op.residual:
                self._last_use[op.residual] = i
            if op.sc_src:
                self._last_use[op.sc_src] = i
        self._out_tensors = {o[1] for o in plan.outputs}
        self.in_place = os.environ.get("B2F_IN_PLACE", "1") != "0"
        self._stem8: Dict[int, tuple] = {}
        self._stem8_img: Dict[int, torch.Tensor] = {}
        self._stem8_w: Optional[torch.Tensor] = None

    # ------------------------------------------------------------------------------------------
    def _bound_for(self, n: int):
        if n in self._bound:
            self._bound[n] = self._bound.pop(n)              # most recently used last
            return self._bound[n]
        while len(self._bound) >= MAX_BOUND:
            old = next(iter(self._bound))
            self._bound.pop(old)
            self._stem8.pop(old, None)
        self._bound[n] = self._bind(n)
        return self._bound[n]

    def _bind(self, n: int):
        """Views and launch descriptors for exactly n images over the buffer pool of capacity_for(n).  Buffer sizes and
        the order they are requested in depend on the capacity only, so every n of one capacity replays the same
        allocation sequence and lands on the same buffers (NHWC with the image index outermost: n images are a prefix)."""
        plan = self.plan
        esz = 2
        cap = capacity_for(n)
        pool = self._pools.setdefault(cap, [])
        cursor = [0]
        free: List[torch.Tensor] = []
        tens: Dict[str, torch.Tensor] = {}
        bound: List[_Bound] = []

        def new_buffer(nbytes: int) -> torch.Tensor:
            j = cursor[0]
            cursor[0] += 1
            if j == len(pool):
                pool.append(torch.empty(nbytes, dtype=torch.uint8, device=self.device))
            assert pool[j].numel() >= nbytes
            return pool[j]

        h0, w0 = plan.in_hw
        in_raw = new_buffer(cap * h0 * w0 * 4 * esz)
        tens[plan.input_name] = in_raw[:n * h0 * w0 * 4 * esz].view(torch_dtype(self.dtype)).view(n, h0, w0, 4)

        def alloc(nbytes: int) -> torch.Tensor:
            best = -1
            for j, b in enumerate(free):
                if b.numel() >= nbytes and (best < 0 or b.numel() < free[best].numel()):
                    best = j
            if best >= 0:
                return free.pop(best)
            return new_buffer(nbytes)

        backing: Dict[str, torch.Tensor] = {}
        for i, op in enumerate(plan.ops):
            spec = plan.tensors[op.dst]
            per_image = spec.h * spec.w * spec.cp * (4 if spec.f32 else esz)
            if self._in_place(i, op, backing):
                raw = backing.pop(op.residual)          # the block output overwrites its identity input
            else:
                raw = alloc((cap * per_image + 255) // 256 * 256)
            backing[op.dst] = raw
            view = raw[:n * per_image].view(torch.float32 if spec.f32 else torch_dtype(self.dtype))
            tens[op.dst] = view.view(n, spec.h, spec.w, spec.cp)
            ws = alloc(8 * cap * per_image) if self._wants_splitk(op, spec) else None    # lives for this launch only
            bound.append(self._bind_op(i, op, n, tens, ws))
            if ws is not None:
                free.append(ws)
            for name in (op.src, op.residual, op.sc_src):
                if name and name in backing and self._last_use.get(name) == i and name not in self._out_tensors:
                    free.append(backing.pop(name))
        return bound, tens, pool

    def _in_place(self, i: int, op: FusedOp, backing: Dict[str, torch.Tensor]) -> bool:
        """A residual convolution without an activation after the add (every IResNet block output) may write over its
        residual when this op is the residual's last reader: `b2f_conv2d` then adds through a TMA reduce-store instead
        of loading the residual (include/b2f.h, `residual`).  B2F_IN_PLACE=0 keeps separate buffers."""
        if not self.in_place or op.kind != "conv" or not op.residual or op.res_mode != 1 or op.act != 0:
            return False
        r, d = self.plan.tensors[op.residual], self.plan.tensors[op.dst]
        return (op.residual in backing and op.residual not in (op.src, op.sc_src) and self._last_use.get(op.residual) == i
                and op.residual not in self._out_tensors and not d.f32 and not r.f32 and (r.h, r.w, r.cp) == (d.h, d.w, d.cp))

    @staticmethod
    def _wants_splitk(op: FusedOp, spec_out) -> bool:
        """Layers `b2f_conv2d` may split along K (include/b2f.h, `splitk_ws`): an fp32 output without activation or
        residual and a long reduction -- the embedding layer (Flatten + Gemm as a 7 x 7 valid convolution)."""
        a = op.attrs
        if os.environ.get("B2F_SPLITK", "1") == "0":          # A/B knob for bench.py
            return False
        return (op.kind == "conv" and spec_out.f32 and op.act == 0 and not op.residual and not op.sc_src
                and a["bias_classes"] == 1 and a["kh"] * a["kw"] >= 16)

    def _bind_op(self, i: int, op: FusedOp, n: int, tens: Dict[str, torch.Tensor],
                 ws: Optional[torch.Tensor] = None) -> _Bound:
        a, w = op.attrs, self._weights[i]
        lib = self.lib
        src, dst = tens[op.src], tens[op.dst]
        res = tens[op.residual] if op.residual else None
        spec_in, spec_out = self.plan.tensors[op.src], self.plan.tensors[op.dst]
        if op.kind == "conv":
            d = _lib.ConvDesc()
            d.n, d.h, d.w, d.cin_p = n, a["h"], a["w"], spec_in.cp
            d.ho, d.wo, d.cout_p = a["ho"], a["wo"], spec_out.cp
            d.kh, d.kw, d.stride, d.pad = a["kh"], a["kw"], a["stride"], a["pad"]
            d.dtype = self.dtype
            d.out_dtype = F32 if spec_out.f32 else self.dtype
            d.act = op.act
            d.sig_hi = a.get("sig_hi", 0)
            d.bias_classes = a["bias_classes"]
            d.res_mode = op.res_mode if res is not None else 0
            if res is not None:
                rs = self.plan.tensors[op.residual]
                d.res_h, d.res_w = rs.h, rs.w
            d.force_kchunk = 0
            d.pool = int(a.get("pool", 0))              # dst is then the 3x3 / s2 max-pooled map (b2f.h, `pool`)
            d.in_, d.weight, d.bias = src.data_ptr(), w["weight"].data_ptr(), w["bias"].data_ptr()
            d.slope = _ptr(w.get("slope"))
            d.residual = _ptr(res)
            d.out = dst.data_ptr()
            sc = tens[op.sc_src] if op.sc_src else None
            if sc is not None:                      # projection shortcut fused as extra K
                d.sc_in, d.sc_weight = sc.data_ptr(), w["sc_weight"].data_ptr()
                d.sc_cin_p, d.sc_stride, d.sc_h, d.sc_w = self.plan.tensors[op.sc_src].cp, a["sc_stride"], a["sc_h"], a["sc_w"]
            if ws is not None:
                d.splitk_ws, d.splitk_ws_bytes = ws.data_ptr(), ws.numel()
            return _Bound(lib.b2f_conv2d, (C.byref(d),), (d, src, dst, res, sc, ws))
        if op.kind == "im2col":
            return _Bound(lib.b2f_im2col3x3,
                          (src.data_ptr(), n, a["h"], a["w"], a["stride"], a["ho"], a["wo"], self.dtype, dst.data_ptr()),
                          (src, dst))
        if op.kind == "stem":
            return _Bound(lib.b2f_stem_conv3x3,
                          (src.data_ptr(), n, a["h"], a["w"], 4, a["stride"], w["weight"].data_ptr(),
                           w["bias"].data_ptr(), _ptr(w.get("slope")), op.act, spec_out.cp, self.dtype,
                           dst.data_ptr()), (src, dst))
        if op.kind == "dwconv":
            return _Bound(lib.b2f_dwconv,
                          (src.data_ptr(), n, a["h"], a["w"], spec_in.cp, a["kh"], a["stride"], a["pad"],
                           w["weight"].data_ptr(), w["bias"].data_ptr(), _ptr(w.get("slope")), op.act, self.dtype,
                           dst.data_ptr()), (src, dst))
        if op.kind == "pool":
            return _Bound(lib.b2f_pool,
                          (src.data_ptr(), n, a["h"], a["w"], spec_in.cp, a["k"], a["stride"], a["pad"], a["mode"],
                           a["ho"], a["wo"], self.dtype, dst.data_ptr()), (src, dst))
        if op.kind == "eltwise":
            return _Bound(lib.b2f_eltwise,
                          (src.data_ptr(), _ptr(res), n * a["h"] * a["w"], spec_in.cp, _ptr(w.get("scale")),
                           _ptr(w.get("shift")), _ptr(w.get("slope")), op.act, self.dtype, dst.data_ptr()),
                          (src, dst, res))
        raise AssertionError(op.kind)

    # ------------------------------------------------------------------------------------------
    def input_buffer(self, n: int) -> torch.Tensor:
        """[n, H, W, 4] 16-bit NHWC buffer the preprocess / norm_crop kernels write into."""
        return self._bound_for(n)[1][self.plan.input_name]

    def patch_buffer(self, n: int) -> Optional[Tuple[torch.Tensor, int]]:
        """When the plan starts with the 3x3 patch extraction of the first convolution, the ([n,ho,wo,32] tensor it
        writes, its stride): the fused preprocess / norm_crop kernels fill it directly and `run(start=1)` skips it."""
        op = self.plan.ops[0]
        if op.kind != "im2col":
            return None
        return self._bound_for(n)[1][op.dst], int(op.attrs["stride"])

    def stem_fused(self, n: int):
        """When the plan opens with the stride-2 3x3 patch extraction followed by the first convolution as a 1x1 over the
        27-channel patches (the SCRFD stem) and that convolution is small enough for `b2f_preprocess_conv1` (cout_p 16 or
        32, bias + optional ReLU, 16-bit output): (weight [cout_p][32], bias [cout_p], act, output tensor, cout_p) -- the
        caller runs letterbox + normalise + this convolution as ONE kernel and continues with run(start=2).  Else None."""
        ops = self.plan.ops
        if (len(ops) < 2 or ops[0].kind != "im2col" or ops[0].attrs["stride"] != 2 or ops[1].kind != "conv"
                or ops[1].src != ops[0].dst or ops[1].residual or ops[1].sc_src or ops[1].attrs["kh"] != 1
                or ops[1].act not in (0, 1) or ops[1].attrs["bias_classes"] != 1 or ops[1].attrs.get("pool")
                or self.plan.tensors[ops[1].dst].f32 or self.plan.tensors[ops[1].dst].cp not in (16, 32)):
            return None
        tens = self._bound_for(n)[1]
        w1 = self._weights[1]
        return w1["weight"][0], w1["bias"][0], int(ops[1].act), tens[ops[1].dst], int(self.plan.tensors[ops[1].dst].cp)

    def stem8(self, n: int):
        """When the plan opens with the 3x3 / pad 1 patch extraction followed by the first convolution as a 1x1 over the
        27-channel patches: (image buffer [n,H,W,8], launch) for the 8-channel stem form of `b2f_conv2d` instead -- the
        caller fills the image (RGB in channels 0..2, zeros above), calls launch() and continues with run(start=2).
        Same result as the two ops it replaces up to the order of the fp32 accumulation; 4x less input traffic.
        None when the plan has another shape."""
        ops = self.plan.ops
        if (len(ops) < 2 or ops[0].kind != "im2col" or ops[1].kind != "conv" or ops[1].src != ops[0].dst or ops[1].residual
                or ops[1].sc_src or ops[1].attrs["kh"] != 1 or self.plan.tensors[ops[1].dst].f32):
            return None
        if n not in self._stem8:
            tens = self._bound_for(n)[1]
            a0, w1 = ops[0].attrs, self._weights[1]
            spec_out = self.plan.tensors[ops[1].dst]
            if self._stem8_w is None:
                self._stem8_w = stem8_weights(w1["weight"])
            cap = capacity_for(n)
            if cap not in self._stem8_img:
                self._stem8_img[cap] = torch.zeros((cap, a0["h"], a0["w"], 8), dtype=torch_dtype(self.dtype), device=self.device)
            img = self._stem8_img[cap][:n]
            d = _lib.ConvDesc()
            d.n, d.h, d.w, d.cin_p = n, a0["h"], a0["w"], 8
            d.ho, d.wo, d.cout_p = a0["ho"], a0["wo"], spec_out.cp
            d.kh, d.kw, d.stride, d.pad = 3, 3, a0["stride"], 1
            d.dtype, d.out_dtype, d.act = self.dtype, self.dtype, ops[1].act
            d.bias_classes = ops[1].attrs["bias_classes"]
            d.in_, d.weight, d.bias = img.data_ptr(), self._stem8_w.data_ptr(), w1["bias"].data_ptr()
            d.slope = _ptr(w1.get("slope"))
            d.out = tens[ops[1].dst].data_ptr()
            lib = self.lib

            def launch(d=d, keep=(img, self._stem8_w)):
                rc = lib.b2f_conv2d(C.byref(d), stream_ptr())
                if rc != 0:
                    _lib.check(rc, "b2f_conv2d")
            self._stem8[n] = (img, launch)
        return self._stem8[n]

    def run(self, n: int, timings: Optional[list] = None, start: int = 0) -> Dict[str, torch.Tensor]:
        """Run the net on whatever `input_buffer(n)` holds (or, with start=1, on a filled `patch_buffer(n)`);
        returns {graph output name: [n,H,W,C] fp32 view}.
        `timings`, when given, receives (op index, kind, start event, end event) per launch (bench roofline)."""
        bound, tens, _ = self._bound_for(n)
        sp = stream_ptr()
        for i, b in enumerate(bound):
            if i < start:
                continue
            if timings is not None:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
            rc = b.fn(*b.args, sp)
            if rc != 0:
                _lib.check(rc, b.fn.__name__)
            if timings is not None:
                e1.record()
                timings.append((i, self.plan.ops[i].kind, e0, e1))
        return {name: tens[t][..., off:off + c] for name, t, c, off in self.plan.outputs}

    def op_flops(self, i: int, n: int) -> int:
        op = self.plan.ops[i]
        return 2 * op.attrs.get("macs_per_image", 0) * n

    def num_kernels(self) -> int:
        return len(self.plan.ops)

    def release(self, n: Optional[int] = None) -> None:
        """Drop the launch list of batch size n (all of them, with every buffer pool, when n is None)."""
        if n is None:
            self._bound.clear()
            self._stem8.clear()
            self._pools.clear()
            self._stem8_img.clear()
        else:
            self._bound.pop(n, None)
            self._stem8.pop(n, None)
            cap = capacity_for(n)
            if not any(capacity_for(m) == cap for m in self._bound):
                self._pools.pop(cap, None)
                self._stem8_img.pop(cap, None)

    def buffer_bytes(self) -> int:
        """Bytes of activation storage currently held (all capacities)."""
        return sum(b.numel() for pool in self._pools.values() for b in pool) + \
            sum(t.numel() * t.element_size() for t in self._stem8_img.values())
