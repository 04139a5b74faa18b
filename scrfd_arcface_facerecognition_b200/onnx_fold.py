"""Constant folding of the shape-computation glue that ONNX exporters leave around `Resize` and `Reshape`.

Upstream SCRFD / ArcFace exports (reference models/scrfd.py:52-68, models/arcface.py:18-37 load them through
onnxruntime, which folds this at session creation) carry sub-graphs like

    Shape(x) -> Gather(2) -> Mul(2) -> Unsqueeze -> Concat(Slice(Shape(x)), ...) -> Resize(x, , , sizes)
    Constant -> Reshape(t, shape)

whose values depend only on the (static) input size.  `fold_shape_glue` walks the graph once with the input size
known, evaluates every node whose inputs are constants (or `Shape` of a tensor whose shape is known) with numpy, drops
those nodes, and hands the results to the remaining nodes as initializers -- so `graph.compile_graph` only ever sees
Resize with constant scales / sizes and Reshape with a constant shape.  `Constant` nodes become initializers too.
Data-dependent use of these ops (a Gather over activations, say) is left in place and rejected by the compiler.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import numpy as np

from .onnx_wire import Graph, Node

_DTYPES = {1: np.float32, 2: np.uint8, 3: np.int8, 5: np.int16, 6: np.int32, 7: np.int64, 9: np.bool_, 10: np.float16,
           11: np.float64}
_FOLDABLE = {"Gather", "Unsqueeze", "Squeeze", "Concat", "Cast", "Slice", "Mul", "Div", "Add", "Sub", "Floor", "Ceil", "Neg",
             "Identity", "Reshape", "ConstantOfShape", "Range", "Equal", "Where", "Min", "Max"}


def _axes(node: Node, ins: List[Optional[np.ndarray]], pos: int = 1) -> Optional[List[int]]:
    if "axes" in node.attrs:
        a = node.attrs["axes"]
        return list(a) if isinstance(a, (list, tuple)) else [int(a)]
    if len(ins) > pos and ins[pos] is not None:
        return [int(v) for v in np.asarray(ins[pos]).reshape(-1)]
    return None


def _constant_value(node: Node) -> np.ndarray:
    a = node.attrs
    if "value" in a:
        return np.asarray(a["value"])
    if "value_float" in a:
        return np.asarray(a["value_float"], np.float32)
    if "value_floats" in a:
        return np.asarray(a["value_floats"], np.float32)
    if "value_int" in a:
        return np.asarray(a["value_int"], np.int64)
    if "value_ints" in a:
        return np.asarray(a["value_ints"], np.int64)
    raise NotImplementedError(f"Constant node {node.name or node.outputs[0]} without a tensor / scalar value")


def _eval(node: Node, ins: List[Optional[np.ndarray]]) -> np.ndarray:
    t, a = node.op_type, node.attrs
    x = ins[0]
    if t == "Gather":
        return np.take(x, np.asarray(ins[1]).astype(np.int64), axis=int(a.get("axis", 0)))
    if t == "Unsqueeze":
        y = np.asarray(x)
        for ax in sorted(ax if ax >= 0 else ax + y.ndim + 1 for ax in _axes(node, ins)):
            y = np.expand_dims(y, ax)
        return y
    if t == "Squeeze":
        ax = _axes(node, ins)
        return np.squeeze(x) if ax is None else np.squeeze(x, axis=tuple(ax))
    if t == "Concat":
        return np.concatenate([np.atleast_1d(v) for v in ins], axis=int(a.get("axis", 0)))
    if t == "Cast":
        return np.asarray(x).astype(_DTYPES[int(a["to"])])
    if t == "Slice":
        if "starts" in a:                                  # opset < 10: attributes
            starts, ends = list(a["starts"]), list(a["ends"])
            axes, steps = list(a.get("axes", range(len(starts)))), [1] * len(starts)
        else:
            starts, ends = [int(v) for v in ins[1].reshape(-1)], [int(v) for v in ins[2].reshape(-1)]
            axes = [int(v) for v in ins[3].reshape(-1)] if len(ins) > 3 and ins[3] is not None else list(range(len(starts)))
            steps = [int(v) for v in ins[4].reshape(-1)] if len(ins) > 4 and ins[4] is not None else [1] * len(starts)
        y = np.asarray(x)
        for s, e, ax, st in zip(starts, ends, axes, steps):
            idx = [slice(None)] * y.ndim
            idx[ax] = slice(s, e, st)                      # numpy clamps out-of-range ends like ONNX does
            y = y[tuple(idx)]
        return y
    if t in ("Mul", "Add", "Sub", "Min", "Max"):
        f = {"Mul": np.multiply, "Add": np.add, "Sub": np.subtract, "Min": np.minimum, "Max": np.maximum}[t]
        y = f(ins[0], ins[1])
        for extra in ins[2:]:
            y = f(y, extra)
        return y
    if t == "Div":
        p, q = np.asarray(ins[0]), np.asarray(ins[1])
        if p.dtype.kind in "iu" and q.dtype.kind in "iu":
            return (np.trunc(p / q)).astype(p.dtype)       # ONNX integer division truncates toward zero
        return p / q
    if t == "Floor":
        return np.floor(x)
    if t == "Ceil":
        return np.ceil(x)
    if t == "Neg":
        return -np.asarray(x)
    if t == "Identity":
        return np.asarray(x)
    if t == "Reshape":
        shape = [int(v) for v in ins[1].reshape(-1)]
        src = np.asarray(x)
        shape = [src.shape[i] if v == 0 else v for i, v in enumerate(shape)]
        return src.reshape(shape)
    if t == "ConstantOfShape":
        v = np.asarray(a["value"]).reshape(-1) if "value" in a else np.zeros(1, np.float32)
        return np.full([int(d) for d in np.asarray(x).reshape(-1)], v[0], dtype=v.dtype)
    if t == "Range":
        return np.arange(ins[0].item(), ins[1].item(), ins[2].item()).astype(np.asarray(ins[0]).dtype)
    if t == "Equal":
        return np.equal(ins[0], ins[1])
    if t == "Where":
        return np.where(ins[0], ins[1], ins[2])
    raise NotImplementedError(t)


def fold_shape_glue(g: Graph, in_hw: Tuple[int, int], batch: int = 1) -> Graph:
    """Graph without Constant nodes and without shape arithmetic: see the module docstring.  `in_hw` is the static
    spatial input size the plan is compiled for; `batch` is what `Shape` reports for dimension 0."""
    from .graph import _infer_one                      # per-node (C, H, W) inference shared with the compiler
    consts: Dict[str, np.ndarray] = dict(g.initializers)
    new_init: Dict[str, np.ndarray] = dict(g.initializers)
    shapes: Dict[str, Tuple[int, ...]] = {g.real_inputs()[0].name: (3, in_hw[0], in_hw[1])}
    nodes: List[Node] = []
    for n in g.nodes:
        t = n.op_type
        if t == "Constant":
            consts[n.outputs[0]] = _constant_value(n)
            continue
        if t == "Shape" and n.inputs[0] in shapes:
            full = np.asarray((batch,) + tuple(shapes[n.inputs[0]]), np.int64)
            start, end = int(n.attrs.get("start", 0)), n.attrs.get("end")
            consts[n.outputs[0]] = full[start:(None if end is None else int(end))]
            continue
        present = [i for i in n.inputs if i]
        if t in _FOLDABLE and present and all(i in consts for i in present):
            consts[n.outputs[0]] = np.asarray(_eval(n, [consts[i] if i else None for i in n.inputs]))
            continue
        for i in present:                                  # folded values that feed a real node become initializers
            if i in consts and i not in new_init:
                new_init[i] = consts[i]
        nodes.append(n)
        _infer_one(new_init, n, shapes)
    folded = Graph(nodes, new_init, g.inputs, g.outputs, g.name)
    return folded
