"""TEST INFRASTRUCTURE: fp32 torch-CPU interpreter of a compiled Plan.

Executes the *fused* ops with the kernel-layout arrays produced by graph.compile_graph, using the
same semantics as the CUDA kernels (border-class bias tables, residual / upsampled residual,
activation order), so the folding arithmetic can be checked against the node-by-node oracle
without a GPU.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

from scrfd_arcface_facerecognition_b200.graph import Plan, ACT_PRELU, ACT_RELU, ACT_SIGMOID


def _act(y, act, slope):
    if act == ACT_RELU:
        return torch.relu(y)
    if act == ACT_PRELU:
        return torch.where(y >= 0, y, y * slope.view(1, -1, 1, 1))
    if act == ACT_SIGMOID:
        return torch.sigmoid(y)
    return y


def run_plan(plan: Plan, x_nchw: np.ndarray, quantize=None):
    """x_nchw: (N,3,H,W) float32 blob.  Returns {graph output name: (N,H,W,C) array}.
    `quantize` (e.g. torch.float16) rounds activations / weights like the GPU path does."""
    def qz(t):
        return t.to(quantize).to(torch.float32) if quantize is not None else t

    env = {plan.input_name: qz(torch.from_numpy(np.ascontiguousarray(x_nchw, dtype=np.float32)))}
    for op in plan.ops:
        a = op.attrs
        x = env[op.src]
        slope = torch.from_numpy(op.arrays["slope"]) if "slope" in op.arrays else None
        if op.kind == "im2col":
            n_, c_, h_, w_ = x.shape
            cols = F.unfold(x, 3, padding=1, stride=a["stride"])                   # (N, C*9, L), channel-major
            cols = cols.view(n_, c_, 9, a["ho"], a["wo"]).permute(0, 2, 1, 3, 4)     # -> tap-major: k = tap*3 + ci
            y = cols.reshape(n_, 9 * c_, a["ho"], a["wo"])
        elif op.kind in ("conv", "stem", "dwconv"):
            cout = a["cout"]
            if op.kind == "conv":
                wk = op.arrays["weight"]                                   # (taps, cout_p, cin_p)
                w = torch.from_numpy(wk[:, :cout, :a["cin"]]).permute(1, 2, 0).reshape(cout, a["cin"], a["kh"], a["kw"])
                y = F.conv2d(x, qz(w.contiguous()), None, a["stride"], a["pad"])
                if op.sc_src:                                               # fused projection shortcut (1x1, no padding)
                    ws = torch.from_numpy(op.arrays["sc_weight"][0, :cout, :a["sc_cin"]]).reshape(cout, a["sc_cin"], 1, 1)
                    y = y + F.conv2d(env[op.sc_src], qz(ws.contiguous()), None, a["sc_stride"], 0)
                bias = torch.from_numpy(op.arrays["bias"])[:, :cout]      # (classes, cout)
                if bias.shape[0] == 1:
                    y = y + bias[0].view(1, -1, 1, 1)
                else:
                    ho, wo = y.shape[2], y.shape[3]
                    iy = torch.arange(ho) * a["stride"] - a["pad"]
                    ix = torch.arange(wo) * a["stride"] - a["pad"]
                    cy = torch.where(iy < 0, 0, torch.where(iy + a["kh"] - 1 >= a["h"], 2, 1))
                    cx = torch.where(ix < 0, 0, torch.where(ix + a["kw"] - 1 >= a["w"], 2, 1))
                    cls = cy[:, None] * 3 + cx[None, :]
                    y = y + bias[cls].permute(2, 0, 1).unsqueeze(0)
            elif op.kind == "stem":
                wk = op.arrays["weight"]                                   # (9, 4, cout_p)
                w = torch.from_numpy(wk[:, :a["cin"], :cout]).permute(2, 1, 0).reshape(cout, a["cin"], 3, 3)
                y = F.conv2d(x, w.contiguous(), torch.from_numpy(op.arrays["bias"][:cout]), a["stride"], 1)
            else:
                wk = op.arrays["weight"]                                   # (k*k, c_p)
                w = torch.from_numpy(wk[:, :cout]).t().reshape(cout, 1, a["kh"], a["kw"])
                y = F.conv2d(x, w.contiguous(), torch.from_numpy(op.arrays["bias"][:cout]), a["stride"], a["pad"],
                             groups=cout)
            if op.residual:
                r = env[op.residual]
                if op.res_mode == 2:
                    r = r.repeat_interleave(2, dim=2).repeat_interleave(2, dim=3)
                y = y + r
            if op.act == ACT_SIGMOID and a.get("sig_hi", 0):
                y = torch.cat([torch.sigmoid(y[:, :a["sig_hi"]]), y[:, a["sig_hi"]:]], dim=1)
            else:
                y = _act(y, op.act, slope[:cout] if slope is not None else None)
            if a.get("pool"):                                              # fused 3x3 / s2 / p1 max-pool (b2f.h, `pool`)
                y = F.max_pool2d(qz(y), 3, 2, 1)
        elif op.kind == "pool":
            if a["mode"] == 0:
                y = F.max_pool2d(x, a["k"], a["stride"], a["pad"])
            else:
                y = F.avg_pool2d(x, a["k"], a["stride"], a["pad"], ceil_mode=True, count_include_pad=False)
            assert y.shape[2:] == (a["ho"], a["wo"])
        elif op.kind == "eltwise":
            c = a["c"]
            y = x
            if "scale" in op.arrays:
                y = y * torch.from_numpy(op.arrays["scale"][:c]).view(1, -1, 1, 1) + \
                    torch.from_numpy(op.arrays["shift"][:c]).view(1, -1, 1, 1)
            if op.residual:
                y = y + env[op.residual]
            y = _act(y, op.act, slope[:c] if slope is not None else None)
        else:
            raise AssertionError(op.kind)
        spec = plan.tensors[op.dst]
        if y.dim() == 2:
            y = y.view(y.shape[0], -1, 1, 1)
        env[op.dst] = y if spec.f32 else qz(y)
    return {name: env[t][:, off:off + c].permute(0, 2, 3, 1).contiguous().numpy() for name, t, c, off in plan.outputs}
