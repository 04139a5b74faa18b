"""Batched video loop: the step either side of the hot path in reference main.py.

The reference reads one frame at a time with cv2.VideoCapture, runs `frame_processor` (detect -> one ArcFace call per
face -> Python scan over the targets -> draw) and writes the frame (main.py:108-188).  Once the engine does tens of
thousands of faces per second that loop is the bottleneck, so this module keeps its semantics and changes its shape:

  * `FrameFeeder`   reads frames from any source with `.read() -> (ok, frame)` (cv2.VideoCapture) or any iterable of
                    HxWx3 uint8 BGR arrays, packs them into pinned host batches and uploads batch i+1 on a copy stream
                    while batch i is being processed
  * `VideoRunner`   detect -> align -> embed -> match for a whole batch (`FacePipeline`), then per frame the list the
                    reference would have drawn: (bbox int32[4], name or "Unknown", similarity), with the reference's
                    strict `>` matching rule (main.py:136-142) and its drawing calls (draw_bbox_info / draw_bbox),
                    on the host frame with cv2 or on the device batch with one kernel (`overlay.FrameOverlay`)

Enrolment (`build_targets`, main.py:78-105) is `VideoRunner.enroll(images_with_names)`: largest face of each image.
"""
from __future__ import annotations

from typing import Callable, Dict, Iterable, Iterator, List, Optional, Sequence, Tuple

import numpy as np
import torch

from .arcface import ArcFace
from .gallery import Gallery
from .pipeline import FacePipeline
from .scrfd import SCRFD

__all__ = ["FrameFeeder", "VideoRunner"]


def _frames_of(source) -> Iterator[np.ndarray]:
    if hasattr(source, "read"):                          # cv2.VideoCapture-like
        while True:
            ok, frame = source.read()
            if not ok:
                return
            yield frame
    else:
        yield from source


class FrameFeeder:
    """Pinned, double-buffered host -> device staging of frame batches (all frames of a stream share one size)."""

    def __init__(self, source, batch: int, device: Optional[torch.device] = None):
        if not torch.cuda.is_available():
            raise RuntimeError("FrameFeeder needs a CUDA device: there is no CPU fallback")
        self.frames = _frames_of(source)
        self.batch = int(batch)
        self.device = device or torch.device("cuda", torch.cuda.current_device())
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self._host: List[torch.Tensor] = []
        self._dev: List[torch.Tensor] = []
        self._ready = [torch.cuda.Event(), torch.cuda.Event()]
        self._free = [torch.cuda.Event(), torch.cuda.Event()]

    def _fill(self, slot: int) -> Tuple[int, List[np.ndarray]]:
        kept: List[np.ndarray] = []
        for frame in self.frames:
            if not self._host:
                h, w = frame.shape[:2]
                self._host = [torch.empty((self.batch, h, w, 3), dtype=torch.uint8).pin_memory() for _ in range(2)]
                self._dev = [torch.empty((self.batch, h, w, 3), dtype=torch.uint8, device=self.device) for _ in range(2)]
                for e in self._free:
                    e.record(torch.cuda.current_stream(self.device))
            self._host[slot][len(kept)].copy_(torch.from_numpy(np.ascontiguousarray(frame)))
            kept.append(frame)
            if len(kept) == self.batch:
                break
        return len(kept), kept

    def _upload(self, slot: int, n: int) -> None:
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(self._free[slot])          # the consumer is done with this device buffer
            self._dev[slot][:n].copy_(self._host[slot][:n], non_blocking=True)
            self._ready[slot].record(self.copy_stream)

    def __iter__(self) -> Iterator[Tuple[torch.Tensor, int, List[np.ndarray]]]:
        """yields (device batch [batch,H,W,3] u8 -- rows >= n are stale, n valid frames, the host frames)"""
        slot = 0
        n, kept = self._fill(slot)
        if n == 0:
            return
        self._upload(slot, n)
        while n:
            nxt = slot ^ 1
            n_next, kept_next = self._fill(nxt)                     # host work for batch i+1 overlaps device work of i
            if n_next:
                self._upload(nxt, n_next)
            cur = torch.cuda.current_stream(self.device)
            cur.wait_event(self._ready[slot])
            yield self._dev[slot], n, kept
            self._free[slot].record(cur)
            slot, n, kept = nxt, n_next, kept_next


class VideoRunner:
    def __init__(self, detector: SCRFD, recognizer: ArcFace, max_num: int = 16, similarity_thresh: float = 0.4,
                 batch: int = 16):
        self.det, self.rec = detector, recognizer
        self.max_num, self.thresh, self.batch = int(max_num), float(similarity_thresh), int(batch)
        self.gallery = Gallery()
        self.names: List[str] = []
        self.pipe = FacePipeline(detector, recognizer, self.gallery, max_num=self.max_num,
                                 similarity_thresh=self.thresh)

    # ---- reference build_targets (main.py:78-105) ------------------------------------------------------
    def enroll(self, images: Iterable[Tuple[np.ndarray, str]]) -> List[str]:
        """(image, name) pairs -> enrolled names; images without a detectable face are skipped, as in the reference."""
        for image, name in images:
            _, kpss = self.det.detect(image, max_num=1)
            if len(kpss) == 0:
                continue
            self.gallery.add(self.rec(image, kpss[0])[None])
            self.names.append(name)
        return list(self.names)

    # ---- reference frame_processor over a whole stream (main.py:108-188) ------------------------------------
    def run(self, source, on_frame: Optional[Callable[[np.ndarray, list], None]] = None, draw=False) -> List[list]:
        """For every frame: [(bbox int32[4], name | "Unknown", similarity float)], in detection order.
        `on_frame(frame, faces)` is called per frame (e.g. a cv2.VideoWriter.write).  `draw=True` overlays the
        reference's boxes and labels on the host frame first (cv2, as main.py:144-148 does); `draw="gpu"` paints the same
        pixels on the batch while it is still in HBM (`overlay.FrameOverlay`, one kernel per batch) and copies the drawn
        frames back over the host frames."""
        from . import helpers
        results: List[list] = []
        colors: Dict[str, tuple] = {}
        overlay = None
        if draw == "gpu":
            from .overlay import FrameOverlay
            overlay = FrameOverlay()
        for dev_batch, n, frames in FrameFeeder(source, self.batch):
            out = self.pipe.process(dev_batch)
            det = out["det"][:n].cpu().numpy()
            counts = out["counts"][:n, 0].cpu().numpy()
            if len(self.names):
                score = out["match_score"][:n].cpu().numpy()
                idx = out["match_idx"][:n].cpu().numpy()
            batch_faces = []
            for f in range(n):
                faces = []
                for s in range(int(counts[f])):
                    bbox = det[f, s, :4].astype(np.int32)                 # main.py:133 truncating cast
                    name, sim = "Unknown", 0.0
                    if len(self.names) and idx[f, s] >= 0:
                        name, sim = self.names[int(idx[f, s])], float(score[f, s])
                    faces.append((bbox, name, sim))
                    if name != "Unknown" and draw:
                        colors.setdefault(name, tuple(int(v) for v in np.random.default_rng(len(colors)).integers(0, 256, 3)))
                    if draw is True:
                        if name != "Unknown":
                            helpers.draw_bbox_info(frames[f], bbox, similarity=sim, name=name, color=colors[name])
                        else:
                            helpers.draw_bbox(frames[f], bbox, (255, 0, 0))
                batch_faces.append(faces)
            if overlay is not None:
                drawn = overlay.draw(dev_batch[:n], batch_faces, colors).cpu().numpy()
                for f in range(n):
                    np.copyto(frames[f], drawn[f])
            for f in range(n):
                results.append(batch_faces[f])
                if on_frame is not None:
                    on_frame(frames[f], batch_faces[f])
        return results
