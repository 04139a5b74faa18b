"""GPU parity at the HEADLINE configuration (SCRFD-10G / SCRFD-2.5G + ArcFace-R50) against outputs of the reference's
own classes run verbatim (tests/golden/make_headline_golden.py -> tests/golden/headline_outputs.npz).

Tier T3 of SURVEY.md section 8c, measured and bounded: detections are paired by IoU, the pixel error of boxes and
landmarks and the score error are reported (written to gpurun_out/headline_parity.json when that directory exists) and
asserted; embeddings are compared by cosine (>= 0.999, the north-star bar); identities by top-1 agreement.
The conv nets run with fp16 operands / fp32 accumulation against the reference's fp32: the bounds below are what that
costs in image pixels, they are not the bit-exact tier (decode / NMS given identical head tensors: test_gpu_kernels.py)."""
import json
import os

import numpy as np
import pytest
import torch

from tests.golden import inputs
from tests.golden.make_headline_golden import detect_cases

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REPORT = {}


@pytest.fixture(scope="module")
def gold():
    return dict(np.load(os.path.join(ROOT, "tests", "golden", "headline_outputs.npz")))


@pytest.fixture(scope="module")
def models():
    from models import SCRFD, ArcFace
    cache = {}

    def get(name):
        if name not in cache:
            cache[name] = (ArcFace if name.startswith("w600k") else SCRFD)(os.path.join("weights", name))
        return cache[name]
    return get


def _iou(box, d):
    x1, y1 = np.maximum(box[0], d[:, 0]), np.maximum(box[1], d[:, 1])
    x2, y2 = np.minimum(box[2], d[:, 2]), np.minimum(box[3], d[:, 3])
    inter = np.clip(x2 - x1, 0, None) * np.clip(y2 - y1, 0, None)
    return inter / ((box[2] - box[0]) * (box[3] - box[1]) + (d[:, 2] - d[:, 0]) * (d[:, 3] - d[:, 1]) - inter)


def pair_detections(gd, d):
    """reference row -> engine row (or -1): best IoU, at least 0.7, each engine row used once."""
    used, out = set(), []
    for box in gd:
        if len(d) == 0:
            out.append(-1)
            continue
        iou = _iou(box, d)
        j = int(np.argmax(iou))
        ok = iou[j] >= 0.7 and j not in used
        out.append(j if ok else -1)
        if ok:
            used.add(j)
    return np.asarray(out, np.int64)


def _dump():
    out_dir = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out_dir):
        with open(os.path.join(out_dir, "headline_parity.json"), "w") as f:
            json.dump(REPORT, f, indent=1, sort_keys=True)


@pytest.mark.parametrize("case", detect_cases(), ids=lambda c: c[0])
def test_detect_matches_reference_run(gold, models, case):
    tag, weight, seed, (h, w), max_num, metric = case
    det = models(weight)
    img = inputs.frame(seed, h, w)
    d, k = det.detect(img, max_num=max_num, metric=metric)
    gd, gk = gold[f"detect_{tag}_det"], gold[f"detect_{tag}_kps"]
    assert d.dtype == np.float32 and k.dtype == np.float32 and d.shape[1] == 5 and k.shape[1:] == (5, 2)
    pair = pair_detections(gd, d)
    m = pair >= 0
    box_err = np.abs(d[pair[m], :4] - gd[m, :4])
    kps_err = np.abs(k[pair[m]] - gk[m])
    score_err = np.abs(d[pair[m], 4] - gd[m, 4])
    same_rank = float((pair[m] == np.nonzero(m)[0]).mean()) if m.any() else 1.0
    rep = dict(reference=len(gd), engine=len(d), paired=int(m.sum()), same_rank_fraction=same_rank,
               box_px_max=float(box_err.max(initial=0)), box_px_mean=float(box_err.mean()) if m.any() else 0.0,
               kps_px_max=float(kps_err.max(initial=0)), kps_px_mean=float(kps_err.mean()) if m.any() else 0.0,
               score_abs_max=float(score_err.max(initial=0)), frame=[h, w], max_num=max_num,
               unpaired_reference_scores=[float(s) for s in gd[~m, 4]])
    REPORT[f"detect_{tag}"] = rep
    _dump()
    print(tag, rep)
    # every reference detection that is not a coin flip at the 0.5 threshold (score within 0.02 of it) must be found
    # when the list is not truncated; with max_num the area ranking of near-equal boxes may swap a few
    solid = gd[:, 4] >= 0.52
    if max_num == 0:
        assert m[solid].all(), f"missing reference detections: {gd[~m & solid]}"
        assert abs(len(d) - len(gd)) <= max(2, int(0.1 * len(gd)))
    else:
        assert len(d) == len(gd) and m.mean() >= 0.85
    scale = max(h / 640.0, w / 640.0, 1.0)           # letterbox factor: head error in input pixels is multiplied by it
    assert rep["box_px_max"] <= 0.75 * scale and rep["kps_px_max"] <= 0.75 * scale     # measured: see profiles/r02_headline_parity.json
    assert rep["box_px_mean"] <= 0.15 * scale and rep["kps_px_mean"] <= 0.15 * scale
    assert rep["score_abs_max"] <= 0.02


def test_r50_embeddings_match_reference_run(gold, models):
    rec = models("w600k_r50.onnx")
    assert rec.input_size == (112, 112) and len(rec.output_names) == 1
    cos_all = {}
    img = inputs.frame(70, 1080, 1920)
    kps = gold["detect_10g_1080p_max16_kps"]                       # the reference's own landmarks: embed in isolation
    e = np.stack([rec(img, kk) for kk in kps])
    assert e.shape == (16, 512) and e.dtype == np.float32
    g = gold["arcface_r50_emb_detected"]
    cos_all["detected"] = (e * g).sum(1) / np.linalg.norm(e, axis=1) / np.linalg.norm(g, axis=1)
    img = inputs.smooth_frame(75, 1080, 1920)
    lms = inputs.landmarks(76, 1080, 1920, 6)
    e = np.stack([rec(img, lm) for lm in lms])
    g = gold["arcface_r50_emb_smooth"]
    cos_all["smooth"] = (e * g).sum(1) / np.linalg.norm(e, axis=1) / np.linalg.norm(g, axis=1)
    from utils.helpers import norm_crop_image
    f = rec.get_feat([norm_crop_image(img, lm) for lm in lms[:3]])
    g = gold["arcface_r50_get_feat"]
    cos_all["get_feat"] = (f * g).sum(1) / np.linalg.norm(f, axis=1) / np.linalg.norm(g, axis=1)
    nrm = np.abs(np.linalg.norm(f, axis=1) / np.linalg.norm(g, axis=1) - 1).max()
    REPORT["arcface_r50"] = {k: dict(cos_min=float(v.min()), cos_mean=float(v.mean())) for k, v in cos_all.items()}
    REPORT["arcface_r50"]["norm_rel_err_max"] = float(nrm)
    _dump()
    print(REPORT["arcface_r50"])
    assert min(v.min() for v in cos_all.values()) >= 0.999           # north-star embedding bar
    assert nrm <= 1e-2                                                # the un-normalised length (callers divide by it)


def test_end_to_end_identities_agree_with_reference_run(gold, models):
    """detect -> align -> embed -> match on the engine against a gallery enrolled from the REFERENCE's embeddings:
    every engine face that pairs with a reference detection must come back as that detection's identity."""
    from scrfd_arcface_facerecognition_b200.gallery import Gallery
    det, rec = models("det_10g.onnx"), models("w600k_r50.onnx")
    img = inputs.frame(70, 1080, 1920)
    gd = gold["detect_10g_1080p_max16_det"]
    G = Gallery()
    G.add(np.concatenate([gold["arcface_r50_emb_detected"], inputs.embeddings(77, 4000) * 20]))
    d, k = det.detect(img, max_num=16)
    pair = pair_detections(gd, d)
    agree, sims = 0, []
    for ref_row, eng_row in enumerate(pair):
        if eng_row < 0:
            continue
        idx, s = G.best_match(rec(img, k[eng_row]), 0.4)
        agree += int(idx == ref_row)
        sims.append(s)
    paired = int((pair >= 0).sum())
    REPORT["end_to_end"] = dict(reference_faces=len(gd), paired=paired, identity_agree=agree,
                                similarity_min=float(min(sims)) if sims else None)
    _dump()
    print(REPORT["end_to_end"])
    assert paired >= 14 and agree == paired
    assert min(sims) >= 0.995          # engine landmarks differ by < 1 px from the reference's: the crop moves, the identity does not
