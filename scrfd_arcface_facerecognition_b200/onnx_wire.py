"""Minimal ONNX protobuf wire-format reader / writer (no `onnx` package offline).

The reference loads `weights/{det_500m,det_2.5g,det_10g,w600k_mbf,w600k_r50}.onnx` through
onnxruntime (reference models/scrfd.py:59-62, models/arcface.py:18-21; files listed in
download.sh:12-16).  This module reads exactly the subset of `ModelProto` the engine needs
-- graph nodes, attributes, initializers, graph inputs / outputs -- and can also write it,
so synthetic weights travel in the reference's own file format.

Field numbers follow the public onnx.proto3 schema:
  ModelProto   : ir_version=1 producer_name=2 graph=7 opset_import=8
  GraphProto   : node=1 name=2 initializer=5 input=11 output=12
  NodeProto    : input=1 output=2 name=3 op_type=4 attribute=5
  AttributeProto: name=1 f=2 i=3 s=4 t=5 floats=7 ints=8 type=20
  TensorProto  : dims=1 data_type=2 float_data=4 int32_data=5 int64_data=7 name=8 raw_data=9
  ValueInfoProto: name=1 type=2 ; TypeProto.tensor_type=1 {elem_type=1 shape=2{dim=1{dim_value=1 dim_param=2}}}
"""
from __future__ import annotations

import struct
from dataclasses import dataclass, field
from typing import Any, Dict, List, Optional, Tuple

import numpy as np

# ---------------------------------------------------------------------------------------------
# wire primitives
# ---------------------------------------------------------------------------------------------

_VARINT, _I64, _LEN, _I32 = 0, 1, 2, 5


def _read_varint(buf: memoryview, pos: int) -> Tuple[int, int]:
    result = 0
    shift = 0
    while True:
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not (b & 0x80):
            return result, pos
        shift += 7
        if shift > 70:
            raise ValueError("malformed varint")


def _iter_fields(buf: memoryview):
    """Yield (field_number, wire_type, value) where value is int or memoryview."""
    pos = 0
    n = len(buf)
    while pos < n:
        key, pos = _read_varint(buf, pos)
        fno, wt = key >> 3, key & 7
        if wt == _VARINT:
            v, pos = _read_varint(buf, pos)
        elif wt == _I64:
            v = buf[pos:pos + 8]
            pos += 8
        elif wt == _LEN:
            ln, pos = _read_varint(buf, pos)
            v = buf[pos:pos + ln]
            pos += ln
        elif wt == _I32:
            v = buf[pos:pos + 4]
            pos += 4
        else:
            raise ValueError(f"unsupported wire type {wt}")
        yield fno, wt, v


def _signed64(v: int) -> int:
    return v - (1 << 64) if v >= (1 << 63) else v


def _packed_varints(v, wt) -> List[int]:
    if wt == _VARINT:
        return [_signed64(v)]
    out = []
    pos = 0
    while pos < len(v):
        x, pos = _read_varint(v, pos)
        out.append(_signed64(x))
    return out


def _w_varint(x: int) -> bytes:
    if x < 0:
        x += 1 << 64
    out = bytearray()
    while True:
        b = x & 0x7F
        x >>= 7
        if x:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _w_key(fno: int, wt: int) -> bytes:
    return _w_varint((fno << 3) | wt)


def _w_len(fno: int, payload: bytes) -> bytes:
    return _w_key(fno, _LEN) + _w_varint(len(payload)) + payload


def _w_int(fno: int, x: int) -> bytes:
    return _w_key(fno, _VARINT) + _w_varint(x)


# ---------------------------------------------------------------------------------------------
# data model
# ---------------------------------------------------------------------------------------------

_DTYPES = {1: np.float32, 2: np.uint8, 3: np.int8, 6: np.int32, 7: np.int64, 9: np.bool_,
           10: np.float16, 11: np.float64}
_DTYPE_IDS = {np.dtype(v): k for k, v in _DTYPES.items()}


@dataclass
class Node:
    op_type: str
    inputs: List[str]
    outputs: List[str]
    attrs: Dict[str, Any] = field(default_factory=dict)
    name: str = ""


@dataclass
class ValueInfo:
    name: str
    shape: List[Any]          # ints or str (dim_param) or None
    elem_type: int = 1


@dataclass
class Graph:
    nodes: List[Node]
    initializers: Dict[str, np.ndarray]
    inputs: List[ValueInfo]
    outputs: List[ValueInfo]
    name: str = "graph"

    def real_inputs(self) -> List[ValueInfo]:
        """Graph inputs that are not initializers (old exporters list both)."""
        return [vi for vi in self.inputs if vi.name not in self.initializers]


# ---------------------------------------------------------------------------------------------
# reader
# ---------------------------------------------------------------------------------------------

def _parse_tensor(buf: memoryview) -> Tuple[str, np.ndarray]:
    dims: List[int] = []
    dtype_id = 1
    name = ""
    raw: Optional[memoryview] = None
    floats: List[float] = []
    ints: List[int] = []
    for fno, wt, v in _iter_fields(buf):
        if fno == 1:
            dims.extend(_packed_varints(v, wt))
        elif fno == 2:
            dtype_id = v
        elif fno == 4:
            if wt == _LEN:
                floats.extend(np.frombuffer(v, dtype="<f4").tolist())
            else:
                floats.append(struct.unpack("<f", v)[0])
        elif fno in (5, 7):
            ints.extend(_packed_varints(v, wt))
        elif fno == 8:
            name = bytes(v).decode()
        elif fno == 9:
            raw = v
    if dtype_id not in _DTYPES:
        raise ValueError(f"tensor {name!r}: unsupported ONNX data_type {dtype_id}")
    dt = np.dtype(_DTYPES[dtype_id])
    if raw is not None:
        arr = np.frombuffer(raw, dtype=dt.newbyteorder("<")).astype(dt, copy=True)
    elif floats:
        arr = np.asarray(floats, dtype=dt)
    else:
        arr = np.asarray(ints, dtype=dt)
    return name, arr.reshape(dims)


def _parse_attr(buf: memoryview) -> Tuple[str, Any]:
    name = ""
    f = i = s = t = None
    floats: List[float] = []
    ints: List[int] = []
    for fno, wt, v in _iter_fields(buf):
        if fno == 1:
            name = bytes(v).decode()
        elif fno == 2:
            f = struct.unpack("<f", v)[0]
        elif fno == 3:
            i = _signed64(v)
        elif fno == 4:
            s = bytes(v)
        elif fno == 5:
            t = _parse_tensor(v)[1]
        elif fno == 7:
            if wt == _LEN:
                floats.extend(np.frombuffer(v, dtype="<f4").tolist())
            else:
                floats.append(struct.unpack("<f", v)[0])
        elif fno == 8:
            ints.extend(_packed_varints(v, wt))
    if t is not None:
        return name, t
    if ints:
        return name, ints
    if floats:
        return name, floats
    if s is not None:
        return name, s.decode(errors="replace")
    if f is not None:
        return name, f
    if i is not None:
        return name, i
    return name, []


def _parse_node(buf: memoryview) -> Node:
    node = Node("", [], [])
    for fno, wt, v in _iter_fields(buf):
        if fno == 1:
            node.inputs.append(bytes(v).decode())
        elif fno == 2:
            node.outputs.append(bytes(v).decode())
        elif fno == 3:
            node.name = bytes(v).decode()
        elif fno == 4:
            node.op_type = bytes(v).decode()
        elif fno == 5:
            k, val = _parse_attr(v)
            node.attrs[k] = val
    return node


def _parse_value_info(buf: memoryview) -> ValueInfo:
    name = ""
    shape: List[Any] = []
    elem = 1
    for fno, wt, v in _iter_fields(buf):
        if fno == 1:
            name = bytes(v).decode()
        elif fno == 2:                                   # TypeProto
            for f2, _, v2 in _iter_fields(v):
                if f2 != 1:                              # tensor_type
                    continue
                for f3, _, v3 in _iter_fields(v2):
                    if f3 == 1:
                        elem = v3
                    elif f3 == 2:                        # TensorShapeProto
                        for f4, _, v4 in _iter_fields(v3):
                            if f4 != 1:
                                continue
                            dim: Any = None
                            for f5, _, v5 in _iter_fields(v4):
                                if f5 == 1:
                                    dim = _signed64(v5)
                                elif f5 == 2:
                                    dim = bytes(v5).decode()
                            shape.append(dim)
    return ValueInfo(name, shape, elem)


def parse_model(data: bytes) -> Graph:
    """Parse serialized ModelProto bytes into a Graph."""
    buf = memoryview(data)
    gbuf = None
    for fno, wt, v in _iter_fields(buf):
        if fno == 7 and wt == _LEN:
            gbuf = v
    if gbuf is None:
        raise ValueError("not an ONNX ModelProto: no graph field")
    g = Graph([], {}, [], [])
    for fno, wt, v in _iter_fields(gbuf):
        if fno == 1:
            g.nodes.append(_parse_node(v))
        elif fno == 2:
            g.name = bytes(v).decode()
        elif fno == 5:
            name, arr = _parse_tensor(v)
            g.initializers[name] = arr
        elif fno == 11:
            g.inputs.append(_parse_value_info(v))
        elif fno == 12:
            g.outputs.append(_parse_value_info(v))
    return g


def load_model(path: str) -> Graph:
    with open(path, "rb") as f:
        return parse_model(f.read())


# ---------------------------------------------------------------------------------------------
# writer
# ---------------------------------------------------------------------------------------------

def _ser_tensor(name: str, arr: np.ndarray) -> bytes:
    arr = np.asarray(arr)                       # (np.ascontiguousarray would turn a 0-d scalar into shape (1,))
    out = bytearray()
    for d in arr.shape:
        out += _w_int(1, int(d))
    out += _w_int(2, _DTYPE_IDS[arr.dtype])
    out += _w_len(8, name.encode())
    out += _w_len(9, arr.astype(arr.dtype.newbyteorder("<"), copy=False).tobytes())
    return bytes(out)


def _ser_attr(name: str, val: Any) -> bytes:
    out = bytearray(_w_len(1, name.encode()))
    if isinstance(val, np.ndarray):
        out += _w_len(5, _ser_tensor("", val)) + _w_int(20, 4)
    elif isinstance(val, float):
        out += _w_key(2, _I32) + struct.pack("<f", val) + _w_int(20, 1)
    elif isinstance(val, (int, np.integer)):
        out += _w_int(3, int(val)) + _w_int(20, 2)
    elif isinstance(val, str):
        out += _w_len(4, val.encode()) + _w_int(20, 3)
    elif isinstance(val, (list, tuple)) and val and isinstance(val[0], float):
        out += _w_len(7, np.asarray(val, "<f4").tobytes()) + _w_int(20, 6)
    elif isinstance(val, (list, tuple)):
        out += _w_len(8, b"".join(_w_varint(int(x)) for x in val)) + _w_int(20, 7)
    else:
        raise TypeError(f"attribute {name}: {type(val)}")
    return bytes(out)


def _ser_value_info(vi: ValueInfo) -> bytes:
    dims = bytearray()
    for d in vi.shape:
        if isinstance(d, str):
            dims += _w_len(1, _w_len(2, d.encode()))
        elif d is None:
            dims += _w_len(1, b"")
        else:
            dims += _w_len(1, _w_int(1, int(d)))
    tensor_type = _w_int(1, vi.elem_type) + _w_len(2, bytes(dims))
    return _w_len(1, vi.name.encode()) + _w_len(2, _w_len(1, tensor_type))


def serialize_model(g: Graph, producer: str = "b2f-synthetic") -> bytes:
    gb = bytearray()
    for n in g.nodes:
        nb = bytearray()
        for x in n.inputs:
            nb += _w_len(1, x.encode())
        for x in n.outputs:
            nb += _w_len(2, x.encode())
        if n.name:
            nb += _w_len(3, n.name.encode())
        nb += _w_len(4, n.op_type.encode())
        for k, v in n.attrs.items():
            nb += _w_len(5, _ser_attr(k, v))
        gb += _w_len(1, bytes(nb))
    gb += _w_len(2, g.name.encode())
    for name, arr in g.initializers.items():
        gb += _w_len(5, _ser_tensor(name, arr))
    for vi in g.inputs:
        gb += _w_len(11, _ser_value_info(vi))
    for vi in g.outputs:
        gb += _w_len(12, _ser_value_info(vi))
    model = _w_int(1, 7) + _w_len(2, producer.encode()) + _w_len(7, bytes(gb))
    model += _w_len(8, _w_len(1, b"") + _w_int(2, 11))          # opset_import {domain:"", version:11}
    return model


def save_model(g: Graph, path: str) -> None:
    with open(path, "wb") as f:
        f.write(serialize_model(g))
