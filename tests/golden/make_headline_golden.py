"""Generate tests/golden/headline_outputs.npz: the REFERENCE's API run verbatim on the headline models.

Run in the build container only (needs /root/reference or $B2F_REFERENCE):
    python tests/golden/make_headline_golden.py
`SCRFD("weights/det_10g.onnx").detect`, `SCRFD("weights/det_2.5g.onnx").detect` and
`ArcFace("weights/w600k_r50.onnx")(img, kps)` are the reference's own classes (models/scrfd.py:122-178,
models/arcface.py:54-57) imported unmodified (oracle/ref_loader.py) over the torch-CPU fp32 session shim and the
Umeyama shim (oracle/shims.py) with the seeded synthetic weights of scrfd_arcface_facerecognition_b200/archs.py
(the .onnx files are absent offline).  Inputs are seeded (tests/golden/inputs.py); only outputs are stored.
"""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import ref_loader  # noqa: E402
from tests.golden import inputs  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def detect_cases():
    """(tag, weight file, frame seed, (h, w), max_num, metric)"""
    return [
        ("10g_1080p_all", "det_10g.onnx", 70, (1080, 1920), 0, "max"),
        ("10g_1080p_max16", "det_10g.onnx", 70, (1080, 1920), 16, "max"),
        ("10g_1080p_b_max16", "det_10g.onnx", 71, (1080, 1920), 16, "max"),
        ("10g_640_all", "det_10g.onnx", 72, (640, 640), 0, "max"),
        ("10g_640_max16", "det_10g.onnx", 72, (640, 640), 16, "default"),
        ("2.5g_1080p_max50", "det_2.5g.onnx", 73, (1080, 1920), 50, "max"),
        ("2.5g_720p_all", "det_2.5g.onnx", 74, (720, 1280), 0, "max"),
    ]


class _Tap:
    """Records what the reference's own `session.run` returned during detect() (models/scrfd.py:83): the raw head tensors
    are the ground truth for the network's numerical error, independent of which near-tied anchor wins the NMS."""

    def __init__(self, session):
        self.session, self.last = session, None

    def run(self, names, feed):
        self.last = self.session.run(names, feed)
        return self.last

    def __getattr__(self, name):
        return getattr(self.session, name)


CAND_FLOOR = 0.3        # anchors whose reference score reaches this are stored (a few hundred per frame)


def candidates(heads):
    """global anchor index (level base + pixel * 2 + anchor), score, bbox [4], kps [10] (stride units) of every anchor with
    score >= CAND_FLOOR, from the nine reference outputs (three strides x score / bbox / kps)."""
    idx, sc, bb, kp = [], [], [], []
    base = 0
    for lvl in range(3):
        s, b, k = heads[lvl].reshape(-1), heads[lvl + 3].reshape(-1, 4), heads[lvl + 6].reshape(-1, 10)
        sel = np.nonzero(s >= CAND_FLOOR)[0]
        idx.append(sel + base), sc.append(s[sel]), bb.append(b[sel]), kp.append(k[sel])
        base += len(s)
    return (np.concatenate(idx).astype(np.int64), np.concatenate(sc).astype(np.float32),
            np.concatenate(bb).astype(np.float32), np.concatenate(kp).astype(np.float32))


def main():
    ref = ref_loader.load()
    if ref is None:
        raise SystemExit("reference tree not found (set B2F_REFERENCE)")
    gold = {}
    dets = {}
    for tag, weight, seed, (h, w), max_num, metric in detect_cases():
        if weight not in dets:
            dets[weight] = ref.SCRFD(os.path.join("weights", weight))
            dets[weight].session = _Tap(dets[weight].session)
        img = inputs.frame(seed, h, w)
        d, k = dets[weight].detect(img, max_num=max_num, metric=metric)
        gold[f"detect_{tag}_det"], gold[f"detect_{tag}_kps"] = d, k
        ci, cs, cb, ck = candidates(dets[weight].session.last)
        gold[f"detect_{tag}_cand_anchor"], gold[f"detect_{tag}_cand_score"] = ci, cs
        gold[f"detect_{tag}_cand_bbox"], gold[f"detect_{tag}_cand_kps"] = cb, ck
        print(f"detect_{tag}: {len(d)} detections, score range {d[:, 4].min() if len(d) else 0:.3f}..{d[:, 4].max() if len(d) else 0:.3f}, "
              f"{len(ci)} anchors >= {CAND_FLOOR}")

    rec = ref.ArcFace(os.path.join("weights", "w600k_r50.onnx"))
    # (a) the embeddings of the reference's own detections on a 1080p frame (main.py:130-134: detect, then one call per face)
    img = inputs.frame(70, 1080, 1920)
    kps = gold["detect_10g_1080p_max16_kps"]
    gold["arcface_r50_emb_detected"] = np.stack([rec(img, k) for k in kps])
    # (b) plausible faces on a smooth frame (interpolation-sensitive)
    img = inputs.smooth_frame(75, 1080, 1920)
    lms = inputs.landmarks(76, 1080, 1920, 6)
    gold["arcface_r50_emb_smooth"] = np.stack([rec(img, lm) for lm in lms])
    # (c) get_feat on aligned crops (models/arcface.py:39-52)
    crops = [ref.helpers.norm_crop_image(img, lm) for lm in lms[:3]]
    gold["arcface_r50_get_feat"] = rec.get_feat(crops)

    np.savez_compressed(os.path.join(OUT, "headline_outputs.npz"), **gold)
    for k, v in gold.items():
        print(f"  {k:36s} {str(v.shape):14s} {v.dtype}")
    print("wrote", os.path.join(OUT, "headline_outputs.npz"))


if __name__ == "__main__":
    main()
