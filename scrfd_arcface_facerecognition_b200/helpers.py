"""Geometry / similarity helpers -- host-side mirror of reference utils/helpers.py.

Same names, argument meaning and return types as the reference functions
(utils/helpers.py:18-123); the arithmetic runs in the CUDA kernels behind include/b2f.h.
`draw_bbox` / `draw_bbox_info` (utils/helpers.py:126-179) are overlay-only and stay thin cv2 calls.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib
from .engine import stream_ptr

# reference utils/helpers.py:6-15
reference_alignment = np.array(
    [[[38.2946, 51.6963], [73.5318, 51.5014], [56.0252, 71.7366], [41.5493, 92.3655], [70.7299, 92.2041]]],
    dtype=np.float32)


def _cuda(a: np.ndarray, dtype) -> torch.Tensor:
    if not torch.cuda.is_available():
        raise _lib.B2FError("utils.helpers needs a CUDA device: there is no CPU fallback")
    return torch.from_numpy(np.ascontiguousarray(a, dtype=dtype)).cuda()


def estimate_norm(landmark, image_size=112):
    """(2x3 float64 similarity onto the ArcFace template, template index 0)  -- utils/helpers.py:18-53."""
    assert landmark.shape == (5, 2)
    lm = _cuda(np.asarray(landmark).reshape(1, 10), np.float32)
    m = torch.empty((1, 6), dtype=torch.float64, device=lm.device)
    _lib.check(_lib.lib().b2f_estimate_norm(lm.data_ptr(), 1, int(image_size), m.data_ptr(), stream_ptr()),
               "b2f_estimate_norm")
    return m.cpu().numpy().reshape(2, 3), 0


def norm_crop_image(image, landmark, image_size=112, mode='arcface'):
    """Aligned uint8 BGR crop, bit-exact cv2.warpAffine semantics  -- utils/helpers.py:56-59."""
    assert landmark.shape == (5, 2)
    frame = _cuda(image, np.uint8)[None]
    lm = _cuda(np.asarray(landmark).reshape(1, 10), np.float32)
    idx = torch.zeros(1, dtype=torch.int32, device=frame.device)
    out = torch.empty((1, image_size, image_size, 3), dtype=torch.uint8, device=frame.device)
    _lib.check(_lib.lib().b2f_norm_crop(frame.data_ptr(), frame.shape[1], frame.shape[2], idx.data_ptr(),
                                        lm.data_ptr(), 1, int(image_size), 127.5, float(np.float32(1 / 127.5)), None,
                                        4, 0, out.data_ptr(), None, stream_ptr()), "b2f_norm_crop")
    return out[0].cpu().numpy()


def distance2bbox(points, distance, max_shape=None):
    """[cx-l, cy-t, cx+r, cy+b] per row  -- utils/helpers.py:62-83 (max_shape is never passed by the reference)."""
    if max_shape is not None:
        raise NotImplementedError("max_shape clamping is a torch-only branch in the reference and is never used")
    n = len(points)
    p, d = _cuda(points, np.float32), _cuda(np.asarray(distance)[:, :4], np.float32)
    out = torch.empty((n, 4), dtype=torch.float32, device=p.device)
    _lib.check(_lib.lib().b2f_distance2bbox(p.data_ptr(), d.data_ptr(), n, out.data_ptr(), stream_ptr()),
               "b2f_distance2bbox")
    return out.cpu().numpy()


def distance2kps(points, distance, max_shape=None):
    """points + offsets for each (x, y) landmark pair  -- utils/helpers.py:86-107."""
    if max_shape is not None:
        raise NotImplementedError("max_shape clamping is a torch-only branch in the reference and is never used")
    n, k2 = np.asarray(distance).shape
    p, d = _cuda(points, np.float32), _cuda(distance, np.float32)
    out = torch.empty((n, k2), dtype=torch.float32, device=p.device)
    _lib.check(_lib.lib().b2f_distance2kps(p.data_ptr(), d.data_ptr(), n, k2, out.data_ptr(), stream_ptr()),
               "b2f_distance2kps")
    return out.cpu().numpy()


def compute_similarity(feat1: np.ndarray, feat2: np.ndarray) -> np.float32:
    """Cosine similarity of two feature vectors as np.float32  -- utils/helpers.py:110-123."""
    a = _cuda(np.asarray(feat1).ravel()[None], np.float32)
    b = _cuda(np.asarray(feat2).ravel()[None], np.float32)
    out = torch.empty(1, dtype=torch.float32, device=a.device)
    _lib.check(_lib.lib().b2f_cosine_pairs(a.data_ptr(), b.data_ptr(), 1, a.shape[1], out.data_ptr(), stream_ptr()),
               "b2f_cosine_pairs")
    return np.float32(out.item())


# ---- overlay drawing (out of the hot path; reference utils/helpers.py:126-179) ----------------------

def draw_bbox(image, bbox, color=(0, 255, 0), thickness=3, proportion=0.2):
    import cv2
    x1, y1, x2, y2 = map(int, bbox)
    corner = int(proportion * min(x2 - x1, y2 - y1))
    cv2.rectangle(image, (x1, y1), (x2, y2), color, 1)
    for (cx, cy, dx, dy) in ((x1, y1, 1, 1), (x2, y1, -1, 1), (x1, y2, 1, -1), (x2, y2, -1, -1)):
        cv2.line(image, (cx, cy), (cx + dx * corner, cy), color, thickness)
        cv2.line(image, (cx, cy), (cx, cy + dy * corner), color, thickness)
    return image


def draw_bbox_info(frame, bbox, similarity, name, color):
    """Label, cornered box and similarity bar of one recognised face, on a host frame  -- utils/helpers.py:155-179.
    Same cv2 calls as the reference, so the same pixels (tests/test_overlay_host.py); `overlay.FrameOverlay` paints the
    identical overlay on frames that stay on the GPU."""
    import cv2
    x1, y1, x2, y2 = map(int, bbox)
    cv2.putText(frame, f"{name}: {similarity:.2f}", org=(x1, y1 - 10), fontFace=cv2.FONT_HERSHEY_COMPLEX_SMALL,
                fontScale=1, color=color, thickness=1)
    draw_bbox(frame, bbox, color)
    bar_x, bar_top = x2 + 10, y2 - int(similarity * (y2 - y1))       # a bar right of the box, filled from the bottom
    cv2.rectangle(frame, (bar_x, bar_top), (bar_x + 10, y2), color, cv2.FILLED)
