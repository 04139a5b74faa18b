"""ArcFace embedder on the B200 engine -- host-side mirror of reference models/arcface.py.

Same constructor, attributes and methods as the reference class (models/arcface.py:10-57):
`ArcFace(model_path)(image, kps) -> (512,) float32`, `get_feat(images) -> (B,512)`.
`get_embedding` is the alias the north-star API names; `embed_batch` is the additive batched entry
that keeps crops and embeddings on the device.
"""
from __future__ import annotations

import os

import threading
from typing import List, Optional, Sequence

import numpy as np
import torch

from . import _lib
from .engine import NetEngine, stream_ptr
from .graph import compile_graph
from .scrfd import load_graph

__all__ = ["ArcFace"]


class ArcFace:
    """Drop-in for reference `models.ArcFace` (models/arcface.py:10-57)."""

    def __init__(self, model_path: str = None, session=None) -> None:
        self.session = session
        self.input_mean = 127.5
        self.input_std = 127.5
        self.taskname = "recognition"
        self._lock = threading.RLock()

        if session is not None and hasattr(session, "_graph"):
            graph = session._graph              # another engine-backed session can be shared
        else:
            graph = load_graph(model_path)
        self._graph = graph
        input_cfg = graph.real_inputs()[0]
        input_shape = list(input_cfg.shape)
        self.input_size = tuple(input_shape[2:4][::-1])
        self.input_shape = input_shape
        self.input_name = input_cfg.name
        self.output_names = [o.name for o in graph.outputs]
        assert len(self.output_names) == 1
        self.output_shape = list(graph.outputs[0].shape)

        self._lib = _lib.lib()
        w, h = self.input_size
        self._engine = NetEngine(compile_graph(graph, (h, w)))
        self._scale = float(np.float32(1.0 / self.input_std))   # blobFromImages multiplies by float32(1/std)
        self.fuse_stem = True
        self.stem8 = os.environ.get("B2F_STEM8", "1") != "0"      # first convolution in its 8-channel stem form
        if session is None:
            self.session = self                                  # keeps `recognizer.session` truthy for callers

    # ------------------------------------------------------------------------------------------
    def _embed_loaded(self, n: int) -> torch.Tensor:
        return self._engine.run(n)[self.output_names[0]].reshape(n, -1)

    def get_feat(self, images) -> np.ndarray:
        """(B,512) float32 embeddings of already-aligned uint8 BGR crops (reference models/arcface.py:39-52)."""
        if not isinstance(images, list):
            images = [images]
        with self._lock:
            w, h = self.input_size
            n = len(images)
            x = self._engine.input_buffer(n)
            shapes = {im.shape for im in images}
            if len(shapes) == 1:
                batch = torch.from_numpy(np.ascontiguousarray(np.stack(images))).cuda(non_blocking=True)
                ih, iw = images[0].shape[:2]
                _lib.check(self._lib.b2f_preprocess(batch.data_ptr(), n, ih, iw, w, h, w, h, float(self.input_mean),
                                                    self._scale, x.data_ptr(), 4, self._engine.dtype, stream_ptr()),
                           "b2f_preprocess")
            else:
                for i, im in enumerate(images):
                    t = torch.from_numpy(np.ascontiguousarray(im)).cuda(non_blocking=True)
                    _lib.check(self._lib.b2f_preprocess(t.data_ptr(), 1, im.shape[0], im.shape[1], w, h, w, h,
                                                        float(self.input_mean), self._scale, x[i].data_ptr(), 4,
                                                        self._engine.dtype, stream_ptr()), "b2f_preprocess")
            return self._embed_loaded(n).cpu().numpy()

    def embed_crops(self, crops_u8: torch.Tensor, copy: bool = True) -> torch.Tensor:
        """[F,112,112,3] uint8 BGR aligned crops ON THE DEVICE -> [F,512] f32 (device): `get_feat` (reference
        models/arcface.py:39-52) without the host round trip, for crops that are already aligned (BASELINE config 3).
        blobFromImages' (x - 127.5) * float32(1/127.5), BGR->RGB and the layout change are one kernel writing the
        16-byte-pixel image the first convolution reads; results alias the engine's buffers unless `copy`."""
        with self._lock:
            f = int(crops_u8.shape[0])
            w, h = self.input_size
            assert tuple(crops_u8.shape[1:]) == (h, w, 3) and crops_u8.dtype == torch.uint8 and crops_u8.is_cuda
            stem8 = self._engine.stem8(f) if (self.fuse_stem and self.stem8) else None
            if stem8 is not None:
                with _lib.span("letterbox_kernel<crop8>", f * (w * h * 3 + w * h * 16)):
                    _lib.check(self._lib.b2f_preprocess(crops_u8.data_ptr(), f, h, w, w, h, w, h, float(self.input_mean),
                                                        self._scale, stem8[0].data_ptr(), 8, self._engine.dtype,
                                                        stream_ptr()), "b2f_preprocess")
                stem8[1]()
                out = self._engine.run(f, start=2)[self.output_names[0]].reshape(f, -1)
            else:
                x = self._engine.input_buffer(f)
                _lib.check(self._lib.b2f_preprocess(crops_u8.data_ptr(), f, h, w, w, h, w, h, float(self.input_mean),
                                                    self._scale, x.data_ptr(), 4, self._engine.dtype, stream_ptr()),
                           "b2f_preprocess")
                out = self._embed_loaded(f)
            return out.clone() if copy else out

    def __call__(self, image, kps):
        """Align by the five landmarks and embed: (512,) float32, not L2-normalised
        (reference models/arcface.py:54-57)."""
        kps = np.asarray(kps, dtype=np.float32)
        assert kps.shape == (5, 2)
        with self._lock:
            frame = torch.from_numpy(np.ascontiguousarray(image)).cuda(non_blocking=True)[None]
            lm = torch.from_numpy(kps.reshape(1, 10)).cuda(non_blocking=True)
            idx = torch.zeros(1, dtype=torch.int32, device=frame.device)
            emb = self.embed_batch(frame, idx, lm, copy=False)      # the D2H copy below happens under the lock
            return emb[0].cpu().numpy().flatten()

    get_embedding = __call__

    def embed_batch(self, frames: torch.Tensor, frame_idx: torch.Tensor, kps: torch.Tensor,
                    crops_u8: Optional[torch.Tensor] = None, copy: bool = True) -> torch.Tensor:
        """frames [B,H,W,3] u8 cuda; frame_idx [F] int32; kps [F,10] (or [F,5,2]) f32 -> [F,512] f32 (device).
        norm_crop (similarity estimate + warpAffine + normalise + layout) is one kernel; the net follows.
        The engine writes into persistent buffers shared by every call of the same capacity, so by default the result
        is copied out (on the stream, still under the model lock) and belongs to the caller -- the reference's
        callers run the shared model from a thread pool (duplicate.py:1954).  `copy=False` returns the aliased view;
        it stays valid only until the next call on this model (the CUDA-graph pipeline, which owns the model, uses it)."""
        with self._lock:
            out = self._embed_batch_view(frames, frame_idx, kps, crops_u8)
            return out.clone() if copy else out

    def _embed_batch_view(self, frames, frame_idx, kps, crops_u8):
        f = int(frame_idx.shape[0])
        w, h = self.input_size
        stem8 = self._engine.stem8(f) if (self.fuse_stem and self.stem8 and crops_u8 is None) else None
        if stem8 is not None:
            # aligned crop kept as an 8-channel image (16 B per pixel); the first convolution reads it tap by tap
            if w == h == 112:                       # one CTA per face, crop staged in shared memory
                # algorithmic bytes: a source region the size of the crop (lower bound) + the 16-byte-pixel crop written
                with _lib.span("warp_patches_kernel<image8>", f * (w * h * 3 + w * h * 16)):
                    _lib.check(self._lib.b2f_norm_crop_image8(
                        frames.data_ptr(), frames.shape[1], frames.shape[2], frame_idx.data_ptr(),
                        kps.reshape(f, 10).contiguous().data_ptr(), f, w, float(self.input_mean), self._scale,
                        stem8[0].data_ptr(), self._engine.dtype, stream_ptr()), "b2f_norm_crop_image8")
            else:
                _lib.check(self._lib.b2f_norm_crop(
                    frames.data_ptr(), frames.shape[1], frames.shape[2], frame_idx.data_ptr(),
                    kps.reshape(f, 10).contiguous().data_ptr(), f, w, float(self.input_mean), self._scale,
                    stem8[0].data_ptr(), 8, self._engine.dtype, None, None, stream_ptr()), "b2f_norm_crop")
            stem8[1]()
            return self._engine.run(f, start=2)[self.output_names[0]].reshape(f, -1)
        patches = self._engine.patch_buffer(f) if (self.fuse_stem and crops_u8 is None and w == h == 112) else None
        if patches is not None and patches[1] == 1:
            # norm_crop + blob + first-layer patch extraction in one kernel (one CTA per face)
            _lib.check(self._lib.b2f_norm_crop_patches(
                frames.data_ptr(), frames.shape[1], frames.shape[2], frame_idx.data_ptr(),
                kps.reshape(f, 10).contiguous().data_ptr(), f, w, float(self.input_mean), self._scale,
                patches[0].data_ptr(), self._engine.dtype, stream_ptr()), "b2f_norm_crop_patches")
            return self._engine.run(f, start=1)[self.output_names[0]].reshape(f, -1)
        x = self._engine.input_buffer(f)
        _lib.check(self._lib.b2f_norm_crop(
            frames.data_ptr(), frames.shape[1], frames.shape[2], frame_idx.data_ptr(),
            kps.reshape(f, 10).contiguous().data_ptr(), f, w, float(self.input_mean), self._scale, x.data_ptr(),
            4, self._engine.dtype, None if crops_u8 is None else crops_u8.data_ptr(), None, stream_ptr()),
            "b2f_norm_crop")
        return self._embed_loaded(f)
