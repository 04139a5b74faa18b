"""Overlay drawing on device-resident frames.

The reference draws on the host frame with cv2, one face at a time (main.py:144-148 -> utils/helpers.py:126-179:
`draw_bbox`, `draw_bbox_info`).  Here the frames stay in HBM: every cv2 call those two functions make is lowered to draw
commands (inclusive rectangles and 1-bit mask blits) and one kernel, `b2f_draw_overlay`, paints a whole batch.  The
result is the reference's image byte for byte (tests/test_gpu_overlay.py checks it against the reference's own
functions, including boxes that leave the frame, inverted boxes and overlapping faces).

What cv2 (4.x) writes, established by probing it (tests/test_overlay_host.py re-asserts each fact on the CPU box):
  * `cv2.rectangle(img, p1, p2, color, 1)`      -> rows y1, y2 over [min x, max x], columns x1, x2 over [min y, max y]
  * `cv2.rectangle(img, p1, p2, color, FILLED)` -> [min x, max x] x [min y, max y], both ends inclusive
  * `cv2.line(img, p0, p1, color, 3)` for an axis-aligned segment -> the band two pixels either side of the segment
    (a convex-polygon fill of half-width (3 + 1) / 2) plus, at both end points, the filled circle of radius 2 that
    cv2 rasterises as rows of 1, 3, 5, 3, 1 pixels; a zero-length line is the two discs alone
  * `cv2.putText` clips each stroke against the image before rasterising it, so a label that leaves the frame is NOT a
    crop of the unclipped label: the mask is rendered by cv2 itself on a canvas that is the label's bounding window
    intersected with the frame (identical to the full-frame rendering; cached per label and per clipping geometry)
  * colours go through saturate_cast<uchar>: rounded to nearest, clamped to 0..255 (random.randint(0, 256) may be 256)
All of it is clipped to the frame.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib

__all__ = ["DrawList", "FrameOverlay", "lower_draw_bbox", "lower_draw_bbox_info"]

RECT, MASK = 0, 1
_FONT_ARGS = dict(fontScale=1, thickness=1)            # utils/helpers.py:161-169


def _saturate(color) -> int:
    """cv2 colour scalar (b, g, r) -> packed bytes, as saturate_cast<uchar>(double) does."""
    b, g, r = (int(min(max(np.rint(float(c)), 0), 255)) for c in tuple(color)[:3])
    return b | (g << 8) | (r << 16)


class DrawList:
    """Draw commands of one batch: per frame a list of groups (one per face), per group its commands."""

    def __init__(self, batch: int, hw: Tuple[int, int]):
        self.batch, self.hw = int(batch), (int(hw[0]), int(hw[1]))
        self.groups: List[List[List[tuple]]] = [[] for _ in range(self.batch)]      # [frame][group][cmd]
        self._masks: List[np.ndarray] = []
        self._mask_bytes = 0
        self._text_cache: Dict[tuple, Tuple[np.ndarray, int, int]] = {}

    # ---- primitives (one colour per group) ----------------------------------------------------------
    def begin_face(self, frame: int) -> List[tuple]:
        g: List[tuple] = []
        self.groups[frame].append(g)
        return g

    @staticmethod
    def rect(g: List[tuple], xa: int, ya: int, xb: int, yb: int, bgr: int) -> None:
        g.append((RECT, min(xa, xb), min(ya, yb), max(xa, xb), max(ya, yb), bgr, 0))

    def outline(self, g: List[tuple], x1: int, y1: int, x2: int, y2: int, bgr: int) -> None:
        """cv2.rectangle(img, (x1, y1), (x2, y2), color, 1)"""
        self.rect(g, x1, y1, x2, y1, bgr)
        self.rect(g, x1, y2, x2, y2, bgr)
        self.rect(g, x1, y1, x1, y2, bgr)
        self.rect(g, x2, y1, x2, y2, bgr)

    def thick_line(self, g: List[tuple], xa: int, ya: int, xb: int, yb: int, bgr: int, thickness: int = 3) -> None:
        """cv2.line(img, (xa, ya), (xb, yb), color, 3) for an axis-aligned segment."""
        if thickness != 3:
            raise NotImplementedError("GPU overlay lowers cv2.line for the reference's thickness (3) only")
        if xa != xb and ya != yb:
            raise NotImplementedError("GPU overlay lowers axis-aligned lines only (all the reference draws)")
        if (xa, ya) != (xb, yb):
            if ya == yb:
                self.rect(g, xa, ya - 2, xb, ya + 2, bgr)
            else:
                self.rect(g, xa - 2, ya, xa + 2, yb, bgr)
        for cx, cy in ((xa, ya), (xb, yb)):                       # the radius-2 discs at both ends
            self.rect(g, cx - 2, cy, cx + 2, cy, bgr)
            self.rect(g, cx - 1, cy - 1, cx + 1, cy + 1, bgr)
            self.rect(g, cx, cy - 2, cx, cy + 2, bgr)

    def text(self, g: List[tuple], label: str, org: Tuple[int, int], bgr: int) -> None:
        """cv2.putText(img, label, org, FONT_HERSHEY_COMPLEX_SMALL, 1, color, 1)"""
        import cv2
        h, w = self.hw
        (tw, th), base = cv2.getTextSize(label, cv2.FONT_HERSHEY_COMPLEX_SMALL, _FONT_ARGS["fontScale"], _FONT_ARGS["thickness"])
        m = 4
        bx0, by0, bx1, by1 = org[0] - m, org[1] - th - base - m, org[0] + tw + m, org[1] + th + base + m
        wx0, wy0, wx1, wy1 = max(bx0, 0), max(by0, 0), min(bx1, w), min(by1, h)
        if wx1 <= wx0 or wy1 <= wy0:
            return
        # the rendering depends on the label and on where the frame cuts its window, not on where the window is
        key = (label, wx0 - bx0, wy0 - by0, bx1 - wx1, by1 - wy1)
        if key not in self._text_cache:
            canvas = np.zeros((wy1 - wy0, wx1 - wx0), np.uint8)
            cv2.putText(canvas, label, org=(org[0] - wx0, org[1] - wy0), fontFace=cv2.FONT_HERSHEY_COMPLEX_SMALL,
                        color=255, **_FONT_ARGS)
            mask = np.ascontiguousarray((canvas > 0).astype(np.uint8))
            self._text_cache[key] = (mask, self._mask_bytes, 0)
            self._masks.append(mask.reshape(-1))
            self._mask_bytes += mask.size
        mask, off, _ = self._text_cache[key]
        g.append((MASK, wx0, wy0, mask.shape[1], mask.shape[0], bgr, off))

    # ---- packing ------------------------------------------------------------------------------------
    def pack(self):
        """(cmds uint8 [n * 32], frame_groups int32 [batch + 1], group_cmds int32 [groups + 1], masks uint8)"""
        flat: List[tuple] = []
        frame_groups = [0]
        group_cmds = [0]
        for groups in self.groups:
            for g in groups:
                flat.extend(g)
                group_cmds.append(len(flat))
            frame_groups.append(len(group_cmds) - 1)
        cmds = np.zeros((len(flat), 8), np.int64)
        if flat:
            cmds[:, :7] = np.asarray(flat, dtype=np.int64)
        # coordinates of garbage boxes may leave int32 after the arithmetic above: clamp far outside any frame
        cmds[:, 1:5] = np.clip(cmds[:, 1:5], -(1 << 30), 1 << 30)
        packed = cmds.astype(np.int32)
        masks = np.concatenate(self._masks) if self._masks else np.zeros(1, np.uint8)
        return packed, np.asarray(frame_groups, np.int32), np.asarray(group_cmds, np.int32), masks


def lower_draw_bbox(dl: DrawList, g: List[tuple], bbox, color=(0, 255, 0), thickness: int = 3, proportion: float = 0.2) -> None:
    """reference utils/helpers.py:126-152 (`draw_bbox`) as draw commands."""
    x1, y1, x2, y2 = map(int, bbox)
    corner = int(proportion * min(x2 - x1, y2 - y1))
    bgr = _saturate(color)
    dl.outline(g, x1, y1, x2, y2, bgr)
    for cx, cy, sx, sy in ((x1, y1, 1, 1), (x2, y1, -1, 1), (x1, y2, 1, -1), (x2, y2, -1, -1)):
        dl.thick_line(g, cx, cy, cx + sx * corner, cy, bgr, thickness)
        dl.thick_line(g, cx, cy, cx, cy + sy * corner, bgr, thickness)


def lower_draw_bbox_info(dl: DrawList, g: List[tuple], bbox, similarity, name: str, color) -> None:
    """reference utils/helpers.py:155-179 (`draw_bbox_info`) as draw commands."""
    x1, y1, x2, y2 = map(int, bbox)
    bgr = _saturate(color)
    dl.text(g, f"{name}: {similarity:.2f}", (x1, y1 - 10), bgr)
    lower_draw_bbox(dl, g, bbox, color)
    rect_height = int(similarity * (y2 - y1))
    dl.rect(g, x2 + 10, y2 - rect_height, x2 + 20, y2, bgr)


class FrameOverlay:
    """Paints the reference's per-face overlay on a device batch [B, H, W, 3] uint8 BGR, in place."""

    def __init__(self, device: Optional[torch.device] = None):
        if not torch.cuda.is_available():
            raise _lib.B2FError("FrameOverlay needs a CUDA device: there is no CPU fallback")
        self.device = device or torch.device("cuda", torch.cuda.current_device())
        self.lib = _lib.lib()

    def draw(self, frames: torch.Tensor, faces: Sequence[Sequence[tuple]], colors: Optional[Dict[str, tuple]] = None,
             unknown_color=(255, 0, 0)) -> torch.Tensor:
        """faces[f] = [(bbox int[4], name | "Unknown", similarity)] in the order the reference loop visits them
        (main.py:130-148): a known face gets `draw_bbox_info` in colors[name], an unknown one `draw_bbox` in blue."""
        assert frames.is_cuda and frames.dtype == torch.uint8 and frames.dim() == 4 and frames.shape[3] == 3
        assert frames.is_contiguous()
        b, h, w, _ = frames.shape
        dl = DrawList(b, (h, w))
        for f, per_frame in enumerate(faces):
            for bbox, name, sim in per_frame:
                g = dl.begin_face(f)
                if name != "Unknown":
                    lower_draw_bbox_info(dl, g, bbox, sim, name, (colors or {}).get(name, (0, 255, 0)))
                else:
                    lower_draw_bbox(dl, g, bbox, unknown_color)
        self.paint(frames, dl)
        return frames

    def paint(self, frames: torch.Tensor, dl: DrawList) -> None:
        cmds, frame_groups, group_cmds, masks = dl.pack()
        if len(cmds) == 0:
            return
        b, h, w, _ = frames.shape
        # one upload: the four arrays share a pinned staging buffer (a few KB per batch)
        blob = np.concatenate([cmds.view(np.uint8).reshape(-1), frame_groups.view(np.uint8), group_cmds.view(np.uint8), masks])
        dev = torch.from_numpy(blob).pin_memory().to(self.device, non_blocking=True)
        o1 = cmds.nbytes
        o2 = o1 + frame_groups.nbytes
        o3 = o2 + group_cmds.nbytes
        base = dev.data_ptr()
        _lib.check(self.lib.b2f_draw_overlay(frames.data_ptr(), b, h, w, base, base + o1, base + o2, base + o3,
                                             torch.cuda.current_stream().cuda_stream), "b2f_draw_overlay")
