"""Generate tests/golden/*.npz by running the REFERENCE's own Python verbatim.

Run in the build container only (needs /root/reference or $B2F_REFERENCE):
    python tests/golden/make_golden.py
The reference's models/scrfd.py, models/arcface.py and utils/helpers.py are imported unmodified
(oracle/ref_loader.py) over the two shim modules for its missing wheels (oracle/shims.py); real
cv2 / numpy do the rest.  Inputs are seeded (tests/golden/inputs.py); only outputs are stored.
"""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import ref_loader, shims  # noqa: E402
from tests.golden import inputs  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


class _FakeSession:
    """Feeds fixed head tensors to the reference's SCRFD.forward (reference models/scrfd.py:83)."""

    def __init__(self, outputs):
        self.outputs = outputs

    def run(self, names, feed):
        return [o.copy() for o in self.outputs]


def _ref_detector(ref, in_size=(640, 640)):
    det = ref.SCRFD.__new__(ref.SCRFD)
    det.input_size = in_size
    det.conf_thres, det.iou_thres = 0.5, 0.4
    det.fmc, det._feat_stride_fpn, det._num_anchors, det.use_kps = 3, [8, 16, 32], 2, True
    det.mean, det.std, det.center_cache = 127.5, 128.0, {}
    det.output_names = [f"o{i}" for i in range(9)]
    det.input_names = ["input.1"]
    return det


def postprocess_cases():
    """(name, image_hw, in_size, seed, tie_fraction, max_num, metric)"""
    return [
        ("pp_1080p_s0", (1080, 1920), (640, 640), 0, 0.0, 0, "max"),
        ("pp_1080p_s1_max16", (1080, 1920), (640, 640), 1, 0.0, 16, "max"),
        ("pp_square_s2", (640, 640), (640, 640), 2, 0.0, 0, "max"),
        ("pp_720p_s3_center5", (720, 1280), (640, 640), 3, 0.0, 5, "default"),
        ("pp_portrait_s4", (800, 600), (640, 640), 4, 0.0, 0, "max"),
        ("pp_small_s5", (240, 320), (320, 320), 5, 0.0, 3, "max"),
    ]


def main():
    ref = ref_loader.load()
    if ref is None:
        raise SystemExit("reference tree not found (set B2F_REFERENCE)")
    gold = {}

    # ---- a5..a11: decode / sort / NMS / max_num through the reference's detect() ----------------
    for name, (ih, iw), in_size, seed, ties, max_num, metric in postprocess_cases():
        heads = inputs.head_tensors(seed, in_size[1], in_size[0], ties)
        det = _ref_detector(ref, in_size)
        det.session = _FakeSession(heads)
        img = np.zeros((ih, iw, 3), np.uint8)            # pixel values are irrelevant with a fake session
        d, k = det.detect(img, max_num=max_num, metric=metric)
        gold[name + "_det"], gold[name + "_kps"] = d, k
        # NMS keep list over the score-sorted candidates (reference models/scrfd.py:142-152)
        sl, bl, kl = det.forward(np.zeros((in_size[1], in_size[0], 3), np.uint8), det.conf_thres)
        scores = np.vstack(sl)
        order = scores.ravel().argsort()[::-1]
        new_h = in_size[1] if ih / iw > in_size[1] / in_size[0] else int(in_size[0] * (ih / iw))
        pre = np.hstack((np.vstack(bl) / (float(new_h) / ih), scores)).astype(np.float32)[order]
        gold[name + "_keep"] = np.asarray(det.nms(pre, det.iou_thres), np.int64)
        gold[name + "_ncand"] = np.asarray(len(order))

    # ---- a12 / a13: estimate_norm + norm_crop_image ------------------------------------------------
    for tag, (h, w) in (("1080p", (1080, 1920)), ("vga", (480, 640))):
        img = inputs.smooth_frame(7, h, w)
        lms = inputs.landmarks(8, h, w, 6)
        Ms, crops = [], []
        for lm in lms:
            M, idx = ref.helpers.estimate_norm(lm)
            Ms.append(M)
            crops.append(ref.helpers.norm_crop_image(img, lm))
        gold[f"align_{tag}_M"] = np.stack(Ms)
        gold[f"align_{tag}_crop"] = np.stack(crops)

    # ---- a17 / a18: compute_similarity + best-match scan ---------------------------------------------
    gal = inputs.embeddings(11, 64)
    qs, ids = inputs.planted_queries(gal, 12, 16)
    sims = np.array([[ref.helpers.compute_similarity(t, q) for t in gal] for q in qs], np.float32)
    gold["sim_matrix"] = sims
    best = []
    for q in qs:                                         # reference main.py:136-142 verbatim semantics
        max_similarity, best_idx = 0, -1
        for ti, target in enumerate(gal):
            similarity = ref.helpers.compute_similarity(target, q)
            if similarity > max_similarity and similarity > 0.4:
                max_similarity, best_idx = similarity, ti
        best.append(best_idx)
    gold["best_match"] = np.asarray(best, np.int64)

    # ---- a2..a11 + a4: full detect() through the torch-CPU session on synthetic weights ---------------
    for arch_file, tag in (("det_500m.onnx", "500m"),):
        det = ref.SCRFD(os.path.join("weights", arch_file))
        for fi, (h, w) in enumerate(((640, 640), (480, 640))):
            img = inputs.frame(20 + fi, h, w)
            d, k = det.detect(img, max_num=0)
            gold[f"detect_{tag}_{fi}_det"], gold[f"detect_{tag}_{fi}_kps"] = d, k

    # ---- a14..a16: ArcFace.__call__ through the torch-CPU session on synthetic weights ------------------
    rec = ref.ArcFace(os.path.join("weights", "w600k_mbf.onnx"))
    img = inputs.smooth_frame(30, 480, 640)
    lms = inputs.landmarks(31, 480, 640, 4)
    gold["arcface_mbf_emb"] = np.stack([rec(img, lm) for lm in lms])

    np.savez_compressed(os.path.join(OUT, "reference_outputs.npz"), **gold)
    total = sum(v.nbytes for v in gold.values())
    print(f"wrote {len(gold)} arrays, {total / 1e3:.1f} kB raw ->", os.path.join(OUT, "reference_outputs.npz"))
    for k, v in gold.items():
        print(f"  {k:28s} {str(v.shape):16s} {v.dtype}")


if __name__ == "__main__":
    main()
