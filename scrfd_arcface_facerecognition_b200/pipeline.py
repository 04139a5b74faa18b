"""Batched detect -> align -> embed -> match: the reference's per-frame loop at batch scale.

Reference main.py:108-150 (`frame_processor`) runs detect, then one ArcFace call per face, then a
Python scan over the targets.  Here a batch of same-sized frames goes through four device stages
with static shapes (B frames x max_num face slots), so the whole step is one CUDA graph:
  SCRFD.detect_batch -> ArcFace.embed_batch (fused norm_crop) -> Gallery.match (top-1, strict >).
Slots beyond a frame's detection count are computed but masked out by `valid`.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

from . import _lib
from .arcface import ArcFace
from .gallery import Gallery
from .scrfd import SCRFD


class FacePipeline:
    def __init__(self, detector: SCRFD, recognizer: ArcFace, gallery: Optional[Gallery], max_num: int = 16,
                 similarity_thresh: float = 0.4, metric: str = "max"):
        if max_num <= 0:
            raise ValueError("FacePipeline needs max_num > 0 (static face slots per frame)")
        self.det, self.rec, self.gallery = detector, recognizer, gallery
        self.max_num, self.thresh, self.metric = max_num, float(similarity_thresh), metric
        self._graphs: Dict[tuple, tuple] = {}
        self._idx: Dict[int, torch.Tensor] = {}

    def _frame_idx(self, b: int, device) -> torch.Tensor:
        if b not in self._idx:
            self._idx[b] = torch.arange(b, dtype=torch.int32, device=device).repeat_interleave(self.max_num)
        return self._idx[b]

    def process(self, frames: torch.Tensor) -> Dict[str, torch.Tensor]:
        """frames [B,H,W,3] uint8 on the device.  All results stay on the device."""
        b = frames.shape[0]
        # the pipeline owns both models for the step: results stay in the models' persistent buffers (no copies)
        det, kps, counts = self.det.detect_batch(frames, self.max_num, self.metric, max_det=self.max_num, copy=False)
        emb = self.rec.embed_batch(frames, self._frame_idx(b, frames.device), kps.reshape(b * self.max_num, 10), copy=False)
        out = {"det": det, "kps": kps, "counts": counts, "emb": emb}
        out["valid"] = (torch.arange(self.max_num, device=frames.device)[None, :] < counts[:, 0:1])
        if self.gallery is not None:
            # reference main.py:136-142: best = 0, accept only sim > best and sim > thresh (strict)
            s, i = self.gallery.match(emb, 1, max(self.thresh, 0.0), strict=True)
            out["match_score"], out["match_idx"] = s.reshape(b, self.max_num), i.reshape(b, self.max_num)
        return out

    # ---- CUDA-graph replay of the whole step -----------------------------------------------------
    def capture(self, b: int, h: int, w: int, slot: int = 0):
        """Warm up and capture `process` for a [b,h,w,3] batch.  Returns (static_frames, outputs, graph, kernels).
        `slot` distinguishes several graphs of one shape, each bound to its own input buffer: a feeder that uploads
        batch i+1 straight into the other slot's input while batch i runs needs no device-to-device staging copy.
        All slots write the same output tensors (the models' persistent buffers)."""
        version = self.gallery.version if self.gallery is not None else -1
        key = (b, h, w, slot)
        if key in self._graphs:
            if self._graphs[key][4] == version:
                return self._graphs[key][:4]
            del self._graphs[key]       # the gallery changed since capture: its row count / pointers in the graph are stale
        dev = self.det._engine_for(self.det.input_size[1], self.det.input_size[0]).device
        static = torch.zeros((b, h, w, 3), dtype=torch.uint8, device=dev)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(2):
                self.process(static)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        before = _lib.launch_count()
        with torch.cuda.graph(graph):
            outs = self.process(static)
        kernels = _lib.launch_count() - before
        # the graph holds raw pointers: keep the detector's scratch tensors of this capture alive with it
        keep = [dict(v) for v in self.det._scratch.values()]
        self._graphs[key] = (static, outs, graph, kernels, version, keep)
        return self._graphs[key][:4]
