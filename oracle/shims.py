"""TEST INFRASTRUCTURE -- not part of the product path.

Stand-ins for the two third-party wheels the reference imports but this image lacks
(SURVEY.md section 8c):

  * `onnxruntime.InferenceSession`  (reference models/scrfd.py:4,59-62,83; models/arcface.py:3,18-21,51)
      unpinned in requirements.txt:2-3.  Replaced by oracle.torch_exec.TorchGraph (torch-CPU fp32).
  * `skimage.transform.SimilarityTransform`  (reference utils/helpers.py:3,36,44-45)
      unpinned in requirements.txt:6.  Restated below from the published Umeyama (1991) algorithm
      as implemented by scikit-image `_umeyama` (float64, SVD, reflection fix via det sign).

`install()` injects them into sys.modules so the reference's own files import unmodified.
PARITY NOTE: the reference has no tests or golden vectors -- these shims are
pinned only by (a) real cv2 / numpy running underneath, and (b) the Umeyama closed-form cross-check
in tests/test_oracle.py.
"""
from __future__ import annotations

import os
import sys
import types
from typing import Dict, List

import numpy as np


# ---------------------------------------------------------------------------------------------
# skimage.transform.SimilarityTransform
# ---------------------------------------------------------------------------------------------

def umeyama(src: np.ndarray, dst: np.ndarray, estimate_scale: bool = True) -> np.ndarray:
    """Least-squares similarity transform (Umeyama 1991), as in scikit-image `_umeyama`."""
    src = np.asarray(src, dtype=np.float64)
    dst = np.asarray(dst, dtype=np.float64)
    num, dim = src.shape
    src_mean = src.mean(axis=0)
    dst_mean = dst.mean(axis=0)
    src_demean = src - src_mean
    dst_demean = dst - dst_mean
    A = dst_demean.T @ src_demean / num
    d = np.ones((dim,), dtype=np.float64)
    if np.linalg.det(A) < 0:
        d[dim - 1] = -1
    T = np.eye(dim + 1, dtype=np.float64)
    U, S, V = np.linalg.svd(A)
    rank = np.linalg.matrix_rank(A)
    if rank == 0:
        return np.nan * T
    elif rank == dim - 1:
        if np.linalg.det(U) * np.linalg.det(V) > 0:
            T[:dim, :dim] = U @ V
        else:
            s = d[dim - 1]
            d[dim - 1] = -1
            T[:dim, :dim] = U @ np.diag(d) @ V
            d[dim - 1] = s
    else:
        T[:dim, :dim] = U @ np.diag(d) @ V
    if estimate_scale:
        scale = 1.0 / src_demean.var(axis=0).sum() * (S @ d)
    else:
        scale = 1.0
    T[:dim, dim] = dst_mean - scale * (T[:dim, :dim] @ src_mean.T)
    T[:dim, :dim] *= scale
    return T


class SimilarityTransform:
    def __init__(self, matrix=None):
        self.params = np.eye(3) if matrix is None else np.asarray(matrix, dtype=np.float64)

    def estimate(self, src, dst) -> bool:
        self.params = umeyama(src, dst, True)
        return not np.any(np.isnan(self.params))


# ---------------------------------------------------------------------------------------------
# onnxruntime.InferenceSession
# ---------------------------------------------------------------------------------------------

class _NodeArg:
    def __init__(self, name, shape, type_="tensor(float)"):
        self.name = name
        self.shape = shape
        self.type = type_


_GRAPH_OVERRIDES: Dict[str, object] = {}     # model_path -> Graph, for synthetic weights


def register_graph(path: str, graph) -> None:
    """Make `InferenceSession(path)` resolve to an in-memory Graph (synthetic weights)."""
    _GRAPH_OVERRIDES[os.path.abspath(path)] = graph


class InferenceSession:
    """torch-CPU fp32 session with the two-method surface the reference uses."""

    def __init__(self, model_path, providers=None, sess_options=None, **kw):
        from oracle.torch_exec import TorchGraph
        from scrfd_arcface_facerecognition_b200 import archs, onnx_wire
        key = os.path.abspath(str(model_path))
        if key in _GRAPH_OVERRIDES:
            graph = _GRAPH_OVERRIDES[key]
        elif os.path.exists(key):
            graph = onnx_wire.load_model(key)
        else:
            arch = archs.arch_for_path(key)
            if arch is None:
                raise FileNotFoundError(f"[ONNXRuntimeError] : 3 : NO_SUCHFILE : Load model from {model_path} failed")
            graph = archs.build_arch(arch)
        self._graph = graph
        self._exec = TorchGraph(graph)

    def get_inputs(self) -> List[_NodeArg]:
        return [_NodeArg(vi.name, list(vi.shape)) for vi in self._graph.real_inputs()]

    def get_outputs(self) -> List[_NodeArg]:
        return [_NodeArg(vi.name, list(vi.shape)) for vi in self._graph.outputs]

    def get_providers(self):
        return ["CPUExecutionProvider"]

    def run(self, output_names, input_feed, run_options=None):
        (name, value), = input_feed.items()
        if output_names is None:
            output_names = self._exec.output_names
        res = self._exec.run(np.asarray(value), want=list(output_names))
        return [res[k] for k in output_names]


def install() -> None:
    """Inject `onnxruntime` and `skimage.transform` shims (idempotent; never shadows real wheels)."""
    try:
        import onnxruntime  # noqa: F401
    except Exception:
        m = types.ModuleType("onnxruntime")
        m.InferenceSession = InferenceSession
        m.get_available_providers = lambda: ["CPUExecutionProvider"]
        m.__b2f_shim__ = True
        sys.modules["onnxruntime"] = m
    try:
        from skimage.transform import SimilarityTransform as _S  # noqa: F401
    except Exception:
        sk = types.ModuleType("skimage")
        tr = types.ModuleType("skimage.transform")
        tr.SimilarityTransform = SimilarityTransform
        sk.transform = tr
        sk.__b2f_shim__ = True
        sys.modules["skimage"] = sk
        sys.modules["skimage.transform"] = tr
