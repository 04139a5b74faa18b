"""TEST INFRASTRUCTURE -- not part of the product path.

fp32 torch-CPU executor for ONNX graphs: the stand-in for `onnxruntime.InferenceSession`
(absent offline) behind the reference's own model wrappers (reference models/scrfd.py:59-62,83;
models/arcface.py:18-21,51).  Node-by-node, NCHW, float32, no fusion: deliberately the dumbest
possible reading of the graph so it can serve as the numerical oracle for the CUDA engine.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.
"""
from __future__ import annotations

from typing import Dict, List, Sequence

import numpy as np
import torch
import torch.nn.functional as F

from scrfd_arcface_facerecognition_b200.onnx_wire import Graph, Node


def _pads(p: Sequence[int]):
    # ONNX pads = [top, left, bottom, right]  ->  F.pad order (left, right, top, bottom)
    return (p[1], p[3], p[0], p[2])


class TorchGraph:
    def __init__(self, graph: Graph):
        self.g = graph
        self.consts: Dict[str, torch.Tensor] = {
            k: torch.from_numpy(np.array(v)) for k, v in graph.initializers.items()}
        self.input_name = graph.real_inputs()[0].name
        self.output_names = [o.name for o in graph.outputs]

    @torch.no_grad()
    def run(self, x: np.ndarray, want: Sequence[str] = ()) -> Dict[str, np.ndarray]:
        env: Dict[str, torch.Tensor] = dict(self.consts)
        env[self.input_name] = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32))
        for n in self.g.nodes:
            outs = self._exec(n, env)
            for name, val in zip(n.outputs, outs):
                env[name] = val
        names = list(want) if want else self.output_names
        return {k: env[k].numpy() for k in names}

    def _exec(self, n: Node, env) -> List[torch.Tensor]:
        a = n.attrs
        i = [env[k] if k else None for k in n.inputs]
        t = n.op_type
        if t == "Conv":
            p = a.get("pads", [0, 0, 0, 0])
            x = i[0]
            if p[0] != p[2] or p[1] != p[3]:
                x = F.pad(x, _pads(p))
                pad = (0, 0)
            else:
                pad = (p[0], p[1])
            return [F.conv2d(x, i[1], i[2] if len(i) > 2 else None, stride=tuple(a.get("strides", [1, 1])),
                             padding=pad, dilation=tuple(a.get("dilations", [1, 1])), groups=a.get("group", 1))]
        if t == "BatchNormalization":
            x, g, b, m, v = i
            eps = a.get("epsilon", 1e-5)
            shape = [1, -1] + [1] * (x.dim() - 2)
            return [(x - m.view(shape)) / torch.sqrt(v.view(shape) + eps) * g.view(shape) + b.view(shape)]
        if t == "Relu":
            return [torch.relu(i[0])]
        if t == "PRelu":
            s = i[1]
            if s.dim() == 1 and i[0].dim() == 4:
                s = s.view(1, -1, 1, 1)
            return [torch.where(i[0] >= 0, i[0], i[0] * s)]
        if t == "Add":
            return [i[0] + i[1]]
        if t == "Sub":
            return [i[0] - i[1]]
        if t == "Mul":
            return [i[0] * i[1]]
        if t == "Div":
            return [i[0] / i[1]]
        if t == "Sigmoid":
            return [torch.sigmoid(i[0])]
        if t == "MaxPool":
            p = a.get("pads", [0, 0, 0, 0])
            return [F.max_pool2d(i[0], tuple(a["kernel_shape"]), tuple(a.get("strides", [1, 1])),
                                 (p[0], p[1]), ceil_mode=bool(a.get("ceil_mode", 0)))]
        if t == "AveragePool":
            p = a.get("pads", [0, 0, 0, 0])
            return [F.avg_pool2d(i[0], tuple(a["kernel_shape"]), tuple(a.get("strides", [1, 1])),
                                 (p[0], p[1]), ceil_mode=bool(a.get("ceil_mode", 0)),
                                 count_include_pad=bool(a.get("count_include_pad", 0)))]
        if t == "GlobalAveragePool":
            return [i[0].mean(dim=(2, 3), keepdim=True)]
        if t in ("Resize", "Upsample"):
            assert a.get("mode", "nearest") == "nearest", "only nearest resize is supported"
            x = i[0]
            if t == "Upsample":
                scales = i[1]
                sizes = None
            else:
                scales = i[2] if len(i) > 2 and i[2] is not None and i[2].numel() else None
                sizes = i[3] if len(i) > 3 and i[3] is not None and i[3].numel() else None
            if sizes is not None:
                oh, ow = int(sizes[2]), int(sizes[3])
            else:
                oh, ow = int(x.shape[2] * float(scales[2])), int(x.shape[3] * float(scales[3]))
            # asymmetric + floor == integer-ratio nearest replicate
            iy = (torch.arange(oh) * x.shape[2]) // oh
            ix = (torch.arange(ow) * x.shape[3]) // ow
            return [x[:, :, iy][:, :, :, ix]]
        if t == "Transpose":
            return [i[0].permute(*a["perm"]).contiguous()]
        if t == "Reshape":
            shape = [int(v) for v in i[1].tolist()]
            shape = [i[0].shape[k] if v == 0 else v for k, v in enumerate(shape)]
            return [i[0].reshape(shape)]
        if t == "Flatten":
            ax = a.get("axis", 1)
            return [i[0].reshape(int(np.prod(i[0].shape[:ax])), -1)]
        if t == "Gemm":
            A = i[0].t() if a.get("transA", 0) else i[0]
            B = i[1].t() if a.get("transB", 0) else i[1]
            y = a.get("alpha", 1.0) * (A @ B)
            if len(i) > 2 and i[2] is not None:
                y = y + a.get("beta", 1.0) * i[2]
            return [y]
        if t == "MatMul":
            return [i[0] @ i[1]]
        if t == "Concat":
            return [torch.cat([v for v in i], dim=a["axis"])]
        if t == "Shape":
            return [torch.tensor(list(i[0].shape), dtype=torch.int64)]
        if t == "Gather":
            return [torch.index_select(i[0], a.get("axis", 0), i[1].reshape(-1).long()).reshape(
                list(i[0].shape[:a.get("axis", 0)]) + list(i[1].shape) + list(i[0].shape[a.get("axis", 0) + 1:]))]
        if t == "Unsqueeze":
            axes = a.get("axes") or i[1].tolist()
            x = i[0]
            for ax in sorted(axes):
                x = x.unsqueeze(ax)
            return [x]
        if t == "Squeeze":
            axes = a.get("axes") or (i[1].tolist() if len(i) > 1 else None)
            x = i[0]
            if axes is None:
                return [x.squeeze()]
            for ax in sorted(axes, reverse=True):
                x = x.squeeze(ax)
            return [x]
        if t == "Cast":
            to = {1: torch.float32, 6: torch.int32, 7: torch.int64, 11: torch.float64}[a["to"]]
            return [i[0].to(to)]
        if t == "Constant":
            return [torch.from_numpy(np.array(a["value"]))]
        if t == "Slice":
            starts, ends = i[1].tolist(), i[2].tolist()
            axes = i[3].tolist() if len(i) > 3 and i[3] is not None else list(range(len(starts)))
            steps = i[4].tolist() if len(i) > 4 and i[4] is not None else [1] * len(starts)
            x = i[0]
            for s, e, ax, st in zip(starts, ends, axes, steps):
                idx = [slice(None)] * x.dim()
                idx[ax] = slice(s, min(e, x.shape[ax]), st)
                x = x[tuple(idx)]
            return [x]
        if t == "Floor":
            return [torch.floor(i[0])]
        if t in ("Identity", "Dropout"):
            return [i[0]]
        raise NotImplementedError(f"ONNX op {t} is not supported by the oracle executor")
