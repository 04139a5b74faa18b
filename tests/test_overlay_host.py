"""CPU: the overlay's host logic.  (a) this repo's `utils.helpers.draw_bbox / draw_bbox_info` write the pixels the
reference's functions write (reference utils/helpers.py:126-179, run verbatim when its tree is present); (b) the cv2
facts the GPU lowering rests on; (c) the lowering itself (scrfd_arcface_facerecognition_b200/overlay.py), executed in
numpy, equals the cv2 drawing for boxes inside, across and outside the frame, tiny / inverted boxes and overlaps."""
import cv2
import numpy as np
import pytest

from tests import overlay_util as ou


def _blank(n, h, w, seed=0):
    rng = np.random.default_rng(seed)
    return [rng.integers(0, 256, (h, w, 3), dtype=np.uint8) for _ in range(n)]


def test_host_helpers_match_reference(ref):
    if ref is None:
        pytest.skip("reference tree not present")
    from utils import helpers as ours
    rng = np.random.default_rng(11)
    faces = ou.random_faces(rng, 360, 480, 24)
    a = ou.draw_host(ref.helpers, _blank(24, 360, 480), faces)
    b = ou.draw_host(ours, _blank(24, 360, 480), faces)
    for x, y in zip(a, b):
        np.testing.assert_array_equal(x, y)


def test_cv2_facts():
    img = np.zeros((40, 40, 3), np.uint8)
    cv2.rectangle(img, (5, 22), (15, 18), (300, -4, 254.6), cv2.FILLED)        # any corner order, inclusive, saturating
    ys, xs = np.nonzero(img[:, :, 0])
    assert (ys.min(), ys.max(), xs.min(), xs.max()) == (18, 22, 5, 15) and tuple(img[18, 5]) == (255, 0, 255)
    img = np.zeros((40, 40), np.uint8)
    cv2.line(img, (10, 10), (16, 10), 255, 3)                                   # band of +-2 and radius-2 discs
    rows = {y: (int(np.nonzero(img[y])[0].min()), int(np.nonzero(img[y])[0].max())) for y in range(8, 13)}
    assert rows == {8: (10, 16), 9: (9, 17), 10: (8, 18), 11: (9, 17), 12: (10, 16)}
    img = np.zeros((40, 40), np.uint8)
    cv2.line(img, (10, 10), (10, 10), 255, 3)
    assert int((img > 0).sum()) == 13 and img[8, 10] and img[10, 8] and img[9, 9] and not img[8, 9]


@pytest.mark.parametrize("hw", [(360, 480), (120, 90)])
def test_lowering_equals_cv2(hw):
    from scrfd_arcface_facerecognition_b200 import overlay
    from utils import helpers as ours
    h, w = hw
    rng = np.random.default_rng(5)
    n = 40
    faces = ou.random_faces(rng, h, w, n, max_faces=8)
    want = ou.draw_host(ours, _blank(n, h, w, 1), faces)
    dl = overlay.DrawList(n, (h, w))
    for f, per_frame in enumerate(faces):
        for bbox, name, sim in per_frame:
            g = dl.begin_face(f)
            if name != "Unknown":
                overlay.lower_draw_bbox_info(dl, g, bbox, sim, name, ou.COLORS[name])
            else:
                overlay.lower_draw_bbox(dl, g, bbox, (255, 0, 0))
    got = ou.execute_draw_list(dl, _blank(n, h, w, 1))
    for f, (x, y) in enumerate(zip(want, got)):
        assert (x == y).all(), f"frame {f}: {int((x != y).any(axis=2).sum())} pixels differ, faces {faces[f]}"
