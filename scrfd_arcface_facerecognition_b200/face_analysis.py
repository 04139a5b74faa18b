"""`FaceAnalysis`-compatible facade over the engine.

The reference's web application never touches `models.SCRFD` / `models.ArcFace`; it goes through insightface's
`FaceAnalysis(name).prepare(ctx_id, det_size).get(img) -> [Face]` (reference duplicate.py:353-359, 1473-1496;
compare_face_from_api.py:65-74, 150, 217-218) and reads `.bbox .kps .det_score .embedding .normed_embedding`.
This class offers that surface on top of SCRFD + ArcFace of this package, so those call sites run on the B200
engine unchanged.  Model packs follow insightface's naming: buffalo_l = SCRFD-10G + ArcFace-R50 (w600k_r50),
buffalo_m = SCRFD-2.5G + R50, buffalo_s / buffalo_sc = SCRFD-500M + MobileFaceNet (w600k_mbf).
All faces of an image are embedded in one batched pass (fused norm_crop + net), not one call per face.
"""
from __future__ import annotations

import os
from typing import List, Optional

import numpy as np
import torch

from .arcface import ArcFace
from .scrfd import SCRFD

__all__ = ["FaceAnalysis", "Face"]

MODEL_PACKS = {
    "buffalo_l": ("det_10g.onnx", "w600k_r50.onnx"),
    "buffalo_m": ("det_2.5g.onnx", "w600k_r50.onnx"),
    "buffalo_s": ("det_500m.onnx", "w600k_mbf.onnx"),
    "buffalo_sc": ("det_500m.onnx", "w600k_mbf.onnx"),
    "antelopev2": ("det_10g.onnx", "w600k_r50.onnx"),
}


class Face(dict):
    """Attribute-style record like insightface.app.common.Face: missing attributes read as None."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return self.get(name)

    def __setattr__(self, name, value):
        self[name] = value

    @property
    def embedding_norm(self):
        e = self.get("embedding")
        return None if e is None else float(np.linalg.norm(e))

    @property
    def normed_embedding(self):
        e = self.get("embedding")
        return None if e is None else e / np.linalg.norm(e)


class FaceAnalysis:
    def __init__(self, name: str = "buffalo_l", root: str = "weights", allowed_modules: Optional[List[str]] = None,
                 providers=None, **kwargs):
        if name not in MODEL_PACKS:
            raise ValueError(f"unknown model pack {name!r}; known: {sorted(MODEL_PACKS)}")
        det_file, rec_file = MODEL_PACKS[name]
        self.name = name
        self.allowed_modules = allowed_modules
        self._det_path, self._rec_path = os.path.join(root, det_file), os.path.join(root, rec_file)
        self.det_model: Optional[SCRFD] = None
        self.rec_model: Optional[ArcFace] = None
        self.models = {}
        self.det_thresh, self.det_size = 0.5, (640, 640)

    def prepare(self, ctx_id: int = 0, det_thresh: float = 0.5, det_size=(640, 640)):
        if ctx_id is not None and ctx_id >= 0:
            torch.cuda.set_device(ctx_id)
        self.det_thresh, self.det_size = det_thresh, tuple(det_size)
        self.det_model = SCRFD(self._det_path, input_size=self.det_size, conf_thres=det_thresh)
        self.models = {"detection": self.det_model}
        if self.allowed_modules is None or "recognition" in self.allowed_modules:
            self.rec_model = ArcFace(self._rec_path)
            self.models["recognition"] = self.rec_model

    def get(self, img: np.ndarray, max_num: int = 0) -> List[Face]:
        if self.det_model is None:
            raise RuntimeError("FaceAnalysis.prepare() must be called before get()")
        bboxes, kpss = self.det_model.detect(img, max_num=max_num, metric="default")
        if bboxes.shape[0] == 0:
            return []
        embs = None
        if self.rec_model is not None:
            frame = torch.from_numpy(np.ascontiguousarray(img)).cuda(non_blocking=True)[None]
            lm = torch.from_numpy(np.ascontiguousarray(kpss, dtype=np.float32).reshape(-1, 10)).cuda(non_blocking=True)
            idx = torch.zeros(lm.shape[0], dtype=torch.int32, device=frame.device)
            with self.rec_model._lock:       # D2H inside the model lock: another thread's run cannot overwrite the buffer
                embs = self.rec_model.embed_batch(frame, idx, lm, copy=False).cpu().numpy()
        faces = []
        for i in range(bboxes.shape[0]):
            face = Face(bbox=bboxes[i, 0:4], kps=kpss[i], det_score=float(bboxes[i, 4]))
            if embs is not None:
                face["embedding"] = embs[i].flatten()
            faces.append(face)
        return faces
