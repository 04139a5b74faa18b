"""SCRFD face detector on the B200 engine -- host-side mirror of reference models/scrfd.py.

Same constructor, attributes and method signatures as the reference class (models/scrfd.py:18-207);
everything between the uint8 frame and the (det, kpss) arrays runs in CUDA kernels reached through
the C-ABI (include/b2f.h):  letterbox+normalise -> conv net -> decode/threshold/sort/NMS/max_num.
Additive API: `detect_batch` keeps results on the device for the batched pipeline.
"""
from __future__ import annotations

import ctypes as C
import os
import threading
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

from . import _lib, archs, onnx_wire
from .engine import NetEngine, stream_ptr
from .graph import compile_graph

__all__ = ["SCRFD"]


def synthetic_weights_allowed() -> bool:
    return os.environ.get("B2F_SYNTHETIC_WEIGHTS", "0") == "1"


def load_graph(model_path: str):
    """ONNX initialisers from `model_path`.  A missing file raises FileNotFoundError, as the reference's
    `onnxruntime.InferenceSession` does (models/scrfd.py:59-68 prints and re-raises).  Only with the explicit opt-in
    B2F_SYNTHETIC_WEIGHTS=1 (the offline tests and bench.py set it: the five files of download.sh:12-16 cannot be
    fetched there) is a missing file of one of those five names replaced by the same architecture with seeded
    random weights -- loudly, because such a model detects and embeds noise."""
    if model_path is not None and os.path.isfile(model_path):
        return onnx_wire.load_model(model_path)
    arch = archs.arch_for_path(str(model_path))
    if arch is None or not synthetic_weights_allowed():
        hint = "" if arch is None else " (set B2F_SYNTHETIC_WEIGHTS=1 to run the architecture with random weights)"
        raise FileNotFoundError(f"Load model from {model_path} failed: file not found{hint}")
    import warnings
    warnings.warn(f"{model_path} is missing: running {arch} with SEEDED RANDOM weights (B2F_SYNTHETIC_WEIGHTS=1); "
                  "detections and embeddings are meaningless outside tests and benchmarks", RuntimeWarning, stacklevel=3)
    return archs.build_arch(arch)


def letterbox_geometry(img_h: int, img_w: int, in_w: int, in_h: int) -> Tuple[int, int, float]:
    """Aspect-preserving size and det_scale, python-float arithmetic as reference models/scrfd.py:123-134."""
    im_ratio = float(img_h) / img_w
    model_ratio = in_h / in_w
    if im_ratio > model_ratio:
        new_height = in_h
        new_width = int(new_height / im_ratio)
    else:
        new_width = in_w
        new_height = int(new_width * im_ratio)
    return new_width, new_height, float(new_height) / img_h


class SCRFD:
    """Drop-in for reference `models.SCRFD` (models/scrfd.py:12-207)."""

    def __init__(self, model_path: str, input_size: Tuple[int] = (640, 640), conf_thres: float = 0.5,
                 iou_thres: float = 0.4) -> None:
        self.input_size = input_size
        self.conf_thres = conf_thres
        self.iou_thres = iou_thres

        # SCRFD model params (reference models/scrfd.py:38-47)
        self.fmc = 3
        self._feat_stride_fpn = [8, 16, 32]
        self._num_anchors = 2
        self.use_kps = True
        self.mean = 127.5
        self.std = 128.0
        self.center_cache = {}
        self.fuse_stem = True               # letterbox + blob + first-layer patches in one kernel when the plan allows
        self.fuse_conv1 = os.environ.get("B2F_FUSE_CONV1", "1") != "0"   # ... and the first convolution itself with them

        self._lock = threading.RLock()      # duplicate.py calls the shared model from a 4-thread pool
        self._engines: Dict[Tuple[int, int], NetEngine] = {}
        self._scratch: Dict[Tuple, Dict[str, torch.Tensor]] = {}
        self._initialize_model(model_path=model_path)

    # ------------------------------------------------------------------------------------------
    def _initialize_model(self, model_path: str):
        try:
            self._graph = load_graph(model_path)
            self.output_names = [o.name for o in self._graph.outputs]
            self.input_names = [i.name for i in self._graph.real_inputs()]
            if len(self.output_names) != 9:
                raise ValueError(f"SCRFD graph must have 9 outputs (got {len(self.output_names)})")
            self._lib = _lib.lib()
            w, h = self.input_size
            self._engine_for(h, w)
        except Exception as e:
            print(f"Failed to load the model: {e}")
            raise

    def _engine_for(self, h: int, w: int) -> NetEngine:
        key = (h, w)
        if key not in self._engines:
            if h % 32 or w % 32:
                raise ValueError("SCRFD input size must be a multiple of 32")
            self._engines[key] = NetEngine(compile_graph(self._graph, (h, w)))
        return self._engines[key]

    def _bufs(self, batch: int, h: int, w: int, max_cand: int, max_det: int) -> Dict[str, torch.Tensor]:
        key = (batch, h, w, max_cand, max_det)
        if key in self._scratch:
            self._scratch[key] = self._scratch.pop(key)         # most recently used last
        else:
            while len(self._scratch) >= 8:                      # a service sees few distinct (batch, shape) keys; cap the rest
                self._scratch.pop(next(iter(self._scratch)))
            dev = self._engine_for(h, w).device
            ws = int(self._lib.b2f_decode_nms_workspace(batch, max_cand))
            self._scratch[key] = dict(
                det=torch.zeros((batch, max_det, 5), dtype=torch.float32, device=dev),
                kps=torch.zeros((batch, max_det, 10), dtype=torch.float32, device=dev),
                keep=torch.zeros((batch, max_det), dtype=torch.int32, device=dev),
                counts=torch.zeros((batch, 4), dtype=torch.int32, device=dev),
                ws=torch.empty(ws, dtype=torch.uint8, device=dev),
                scale=torch.empty(batch, dtype=torch.float32, device=dev),
                hw=torch.empty((batch, 2), dtype=torch.int32, device=dev),
            )
        return self._scratch[key]

    def _levels(self, outs: Dict[str, torch.Tensor]) -> _lib.DetLevels:
        lv = _lib.DetLevels()
        for i in range(3):
            s, b, k = (outs[self.output_names[i + j * self.fmc]] for j in range(3))
            lv.score[i], lv.bbox[i], lv.kps[i] = s.data_ptr(), b.data_ptr(), k.data_ptr()
            lv.score_ps[i], lv.bbox_ps[i], lv.kps_ps[i] = s.stride(-2), b.stride(-2), k.stride(-2)
        return lv

    def _total_anchors(self, h: int, w: int) -> int:
        return sum((h // s) * (w // s) * self._num_anchors for s in self._feat_stride_fpn)

    # ------------------------------------------------------------------------------------------
    def _run_net(self, frames: torch.Tensor, new_w: int, new_h: int, in_w: int, in_h: int):
        """frames: [B,H,W,3] uint8 cuda.  Letterbox + normalise into the engine input, run the net."""
        eng = self._engine_for(in_h, in_w)
        b, h, w, _ = frames.shape
        fused = eng.stem_fused(b) if (self.fuse_stem and self.fuse_conv1) else None
        if fused is not None:             # letterbox + blob + the first convolution in one kernel: no patch tensor at all
            wt, bias, act, out, cout_p = fused
            # algorithmic bytes: the source pixels the resize needs + the first layer's output
            with _lib.span("letterbox_conv1_kernel", b * (new_w * new_h * 3 + out[0].numel() * 2)):
                _lib.check(self._lib.b2f_preprocess_conv1(
                    frames.data_ptr(), b, h, w, new_w, new_h, in_w, in_h, float(self.mean),
                    float(np.float32(1.0 / self.std)), wt.data_ptr(), bias.data_ptr(), cout_p, act, out.data_ptr(),
                    eng.dtype, stream_ptr()), "b2f_preprocess_conv1")
            return eng.run(b, start=2)
        patches = eng.patch_buffer(b) if self.fuse_stem else None
        if patches is not None:           # letterbox + blob + first-layer patch extraction in one pass
            # algorithmic bytes: the source pixels the resize needs + the patch tensor written
            with _lib.span("letterbox_patches_kernel", b * (new_w * new_h * 3 + patches[0][0].numel() * 2)):
                _lib.check(self._lib.b2f_preprocess_patches(
                    frames.data_ptr(), b, h, w, new_w, new_h, in_w, in_h, patches[1], float(self.mean),
                    float(np.float32(1.0 / self.std)), patches[0].data_ptr(), eng.dtype, stream_ptr()),
                    "b2f_preprocess_patches")
            return eng.run(b, start=1)
        x = eng.input_buffer(b)
        _lib.check(self._lib.b2f_preprocess(frames.data_ptr(), b, h, w, new_w, new_h, in_w, in_h,
                                            float(self.mean), float(np.float32(1.0 / self.std)), x.data_ptr(), 4,
                                            eng.dtype, stream_ptr()), "b2f_preprocess")
        return eng.run(b)

    def _decode(self, outs, batch, in_h, in_w, det_scale, image_hw, conf, iou, max_num, metric, max_cand, max_det):
        bufs = self._bufs(batch, in_h, in_w, max_cand, max_det)
        geom = (tuple(float(v) for v in det_scale), tuple(int(v) for row in image_hw for v in row))
        if bufs.get("_geom") != geom:       # unchanged geometry costs no copy (and keeps CUDA-graph capture legal)
            bufs["scale"].copy_(torch.as_tensor(det_scale, dtype=torch.float32))
            bufs["hw"].copy_(torch.as_tensor(image_hw, dtype=torch.int32).reshape(batch, 2))
            bufs["_geom"] = geom
        lv = self._levels(outs)
        # algorithmic bytes: every anchor's score is read (bbox / kps only for candidates), the result rows are written
        with _lib.span("decode_nms_kernel", batch * (self._total_anchors(in_h, in_w) * 4 + max_det * 60)):
            _lib.check(self._lib.b2f_decode_nms(
                C.byref(lv), batch, in_h, in_w, bufs["scale"].data_ptr(), bufs["hw"].data_ptr(), float(conf), float(iou),
                int(max_num), 0 if metric == "max" else 1, max_cand, max_det, bufs["det"].data_ptr(),
                bufs["kps"].data_ptr(), bufs["keep"].data_ptr(), bufs["counts"].data_ptr(), bufs["ws"].data_ptr(),
                bufs["ws"].numel(), stream_ptr()), "b2f_decode_nms")
        return bufs

    # ------------------------------------------------------------------------------------------
    def forward(self, image, threshold):
        """Per-stride (scores, bboxes, kpss) above `threshold`, anchor order, input-pixel units
        (reference models/scrfd.py:70-120).  `image` is the already letterboxed uint8 canvas."""
        with self._lock:
            in_h, in_w = image.shape[0], image.shape[1]
            frames = torch.from_numpy(np.ascontiguousarray(image)).cuda(non_blocking=True)[None]
            outs = self._run_net(frames, in_w, in_h, in_w, in_h)
            total = self._total_anchors(in_h, in_w)
            bufs = self._decode(outs, 1, in_h, in_w, [1.0], [[in_h, in_w]], threshold, -1.0, 0, "max", total, total)
            n = int(bufs["counts"][0, 0].item())
            det = bufs["det"][0, :n].cpu().numpy()
            kps = bufs["kps"][0, :n].cpu().numpy()
            anchor = bufs["keep"][0, :n].cpu().numpy()
        scores_list, bboxes_list, kpss_list = [], [], []
        base = 0
        for s in self._feat_stride_fpn:
            cnt = (in_h // s) * (in_w // s) * self._num_anchors
            m = (anchor >= base) & (anchor < base + cnt)
            scores_list.append(det[m, 4:5].copy())
            bboxes_list.append(det[m, 0:4].copy())
            kpss_list.append(kps[m].reshape(-1, 5, 2).copy())
            base += cnt
        return scores_list, bboxes_list, kpss_list

    def detect(self, image, max_num=0, metric="max"):
        """(det [N,5] f32, kpss [N,5,2] f32) exactly as reference models/scrfd.py:122-178."""
        with self._lock:
            width, height = self.input_size
            new_w, new_h, det_scale = letterbox_geometry(image.shape[0], image.shape[1], width, height)
            frames = torch.from_numpy(np.ascontiguousarray(image)).cuda(non_blocking=True)[None]
            outs = self._run_net(frames, new_w, new_h, width, height)
            total = self._total_anchors(height, width)
            bufs = self._decode(outs, 1, height, width, [np.float32(det_scale)], [[image.shape[0], image.shape[1]]],
                                self.conf_thres, np.float32(self.iou_thres), max_num, metric, total, total)
            n = int(bufs["counts"][0, 0].item())
            det = bufs["det"][0, :n].cpu().numpy()
            kpss = bufs["kps"][0, :n].cpu().numpy().reshape(-1, 5, 2)
        return det, kpss

    def detect_batch(self, frames, max_num=0, metric="max", max_cand: int = 4096, max_det: Optional[int] = None,
                     copy: bool = True):
        """Batched detect over same-sized frames.  frames: [B,H,W,3] uint8 (numpy or cuda tensor).
        Returns device tensors (det [B,max_det,5], kps [B,max_det,5,2], counts [B,4]); row b holds
        counts[b,0] valid detections.  counts[b,3] != 0 flags a candidate / detection overflow.
        The kernels write into scratch tensors kept per (batch, shape): by default the three results are copied out
        under the model lock and belong to the caller (the reference's callers share one model between threads,
        duplicate.py:1954); `copy=False` returns the scratch views, valid until the next call on this model."""
        with self._lock:
            if isinstance(frames, np.ndarray):
                frames = torch.from_numpy(np.ascontiguousarray(frames)).cuda(non_blocking=True)
            b, h, w, _ = frames.shape
            width, height = self.input_size
            new_w, new_h, det_scale = letterbox_geometry(h, w, width, height)
            outs = self._run_net(frames, new_w, new_h, width, height)
            total = self._total_anchors(height, width)
            max_cand = min(max_cand, total)
            if max_det is None:
                max_det = max_num if max_num > 0 else max_cand
            bufs = self._decode(outs, b, height, width, [np.float32(det_scale)] * b, [[h, w]] * b, self.conf_thres,
                                np.float32(self.iou_thres), max_num, metric, max_cand, max_det)
            if copy:
                return bufs["det"].clone(), bufs["kps"].clone().view(b, max_det, 5, 2), bufs["counts"].clone()
            return bufs["det"], bufs["kps"].view(b, max_det, 5, 2), bufs["counts"]

    def nms(self, dets, iou_thres):
        """Greedy NMS keep list (reference models/scrfd.py:180-207) computed on the GPU."""
        dets = np.ascontiguousarray(dets, dtype=np.float32)
        n = dets.shape[0]
        if n == 0:
            return []
        with self._lock:
            d = torch.from_numpy(dets).cuda()
            keep = torch.empty(n, dtype=torch.int32, device=d.device)
            cnt = torch.zeros(1, dtype=torch.int32, device=d.device)
            p2 = 32
            while p2 < n:
                p2 *= 2
            ws = torch.empty(p2 * 8, dtype=torch.uint8, device=d.device)
            _lib.check(self._lib.b2f_nms(d.data_ptr(), n, float(np.float32(iou_thres)), keep.data_ptr(),
                                         cnt.data_ptr(), ws.data_ptr(), ws.numel(), stream_ptr()), "b2f_nms")
            k = int(cnt.item())
            return [int(v) for v in keep[:k].cpu().numpy()]
