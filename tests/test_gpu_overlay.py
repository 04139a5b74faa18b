"""GPU: `b2f_draw_overlay` (through overlay.FrameOverlay) paints exactly the pixels the reference's draw_bbox /
draw_bbox_info write with cv2 (reference utils/helpers.py:126-179, main.py:144-148) -- bit-exact, including boxes that
cross or leave the frame, tiny and inverted boxes, labels cut by the border and faces painted over each other."""
import numpy as np
import pytest
import torch

from tests import overlay_util as ou

pytestmark = pytest.mark.gpu


def _frames(n, h, w, seed):
    return np.random.default_rng(seed).integers(0, 256, (n, h, w, 3), dtype=np.uint8)


@pytest.mark.parametrize("hw,n,max_faces", [((360, 480), 32, 8), ((1080, 1920), 8, 50), ((97, 131), 16, 5)])
def test_overlay_matches_cv2_drawing(ref, hw, n, max_faces):
    from scrfd_arcface_facerecognition_b200.overlay import FrameOverlay
    from utils import helpers as ours
    h, w = hw
    faces = ou.random_faces(np.random.default_rng(h), h, w, n, max_faces=max_faces)
    base = _frames(n, h, w, 2)
    want = ou.draw_host(ref.helpers if ref is not None else ours, [f.copy() for f in base], faces)
    dev = torch.from_numpy(base).cuda()
    got = FrameOverlay().draw(dev, faces, ou.COLORS).cpu().numpy()
    for f in range(n):
        assert (got[f] == want[f]).all(), f"frame {f}: {int((got[f] != want[f]).any(axis=2).sum())} pixels differ"
    assert (got != base).any()


def test_overlay_empty_and_errors():
    from scrfd_arcface_facerecognition_b200 import _lib
    from scrfd_arcface_facerecognition_b200.overlay import FrameOverlay
    base = _frames(3, 64, 64, 3)
    dev = torch.from_numpy(base).cuda()
    out = FrameOverlay().draw(dev, [[], [], []])                       # nothing to draw: frames untouched
    np.testing.assert_array_equal(out.cpu().numpy(), base)
    rc = _lib.lib().b2f_draw_overlay(dev.data_ptr(), 3, 0, 64, None, None, None, None, None)
    assert rc != 0 and b"b2f_draw_overlay" in _lib.lib().b2f_last_error()
