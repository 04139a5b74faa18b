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


def engine_heads(det, img):
    """Raw head tensors of the engine for one image, flattened to the reference's anchor order (level, pixel, anchor):
    (score [A], bbox [A,4], kps [A,10], stride [A]) -- what `session.run` returns at reference models/scrfd.py:83."""
    from scrfd_arcface_facerecognition_b200.scrfd import letterbox_geometry
    w, h = det.input_size
    new_w, new_h, det_scale = letterbox_geometry(img.shape[0], img.shape[1], w, h)
    frames = torch.from_numpy(np.ascontiguousarray(img)).cuda()[None]
    with det._lock:
        outs = det._run_net(frames, new_w, new_h, w, h)
        sc, bb, kp, st = [], [], [], []
        for lvl, stride in enumerate(det._feat_stride_fpn):
            s = outs[det.output_names[lvl]][0].reshape(-1).float().cpu().numpy()
            sc.append(s)
            bb.append(outs[det.output_names[lvl + 3]][0].reshape(-1, 4).float().cpu().numpy())
            kp.append(outs[det.output_names[lvl + 6]][0].reshape(-1, 10).float().cpu().numpy())
            st.append(np.full(len(s), stride, np.float32))
    return np.concatenate(sc), np.concatenate(bb), np.concatenate(kp), np.concatenate(st), np.float32(det_scale)


def anchors_of(det_rows, scores_by_anchor, anchor_ids=None):
    """The anchor each detection row came from: its score is a verbatim copy of that anchor's head output."""
    out = []
    for row in det_rows:
        hit = np.nonzero(scores_by_anchor == row[4])[0]
        out.append(-1 if len(hit) != 1 else int(hit[0] if anchor_ids is None else anchor_ids[hit[0]]))
    return np.asarray(out, np.int64)


def _dump():
    out_dir = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out_dir):
        with open(os.path.join(out_dir, "headline_parity.json"), "w") as f:
            json.dump(REPORT, f, indent=1, sort_keys=True)


# bounds on the network's numerical error at the anchors that matter (reference score >= 0.3), in IMAGE pixels per unit of
# letterbox magnification (1080p -> 640 is x3), measured on B200 (profiles/r02_headline_parity.json) with headroom
PX_MAX, PX_MEAN, SCORE_MAX = 0.15, 0.035, 0.004      # measured: 0.107 / 0.025 / 0.0018


@pytest.mark.parametrize("case", detect_cases(), ids=lambda c: c[0])
def test_detect_matches_reference_run(gold, models, case):
    tag, weight, seed, (h, w), max_num, metric = case
    det = models(weight)
    img = inputs.frame(seed, h, w)
    d, k = det.detect(img, max_num=max_num, metric=metric)
    gd, gk = gold[f"detect_{tag}_det"], gold[f"detect_{tag}_kps"]
    assert d.dtype == np.float32 and k.dtype == np.float32 and d.shape[1] == 5 and k.shape[1:] == (5, 2)
    # ---- (1) the network itself: head tensors at the reference's candidate anchors, engine (fp16 operands, fp32
    #          accumulation) against the reference's fp32 run, converted to image pixels -------------------------------
    sc, bb, kp, st, ds = engine_heads(det, img)
    ca, cs = gold[f"detect_{tag}_cand_anchor"], gold[f"detect_{tag}_cand_score"]
    to_px = (st[ca] / ds)[:, None]
    box_err = np.abs(bb[ca] - gold[f"detect_{tag}_cand_bbox"]) * to_px
    kps_err = np.abs(kp[ca] - gold[f"detect_{tag}_cand_kps"]) * to_px
    score_err = np.abs(sc[ca] - cs)
    # ---- (2) the detections: same anchors kept?  (random-init heads put many near-tied, heavily overlapping anchors in
    #          every neighbourhood, so a 1e-3 score difference can hand the NMS win to a neighbour: reported, bounded) ---
    ref_anchor = anchors_of(gd, cs, ca)
    eng_anchor = anchors_of(d, sc)
    assert (ref_anchor >= 0).all() and (eng_anchor >= 0).all()
    common = np.intersect1d(ref_anchor, eng_anchor)
    det_box_err = np.zeros(0)
    if len(common):
        ri = [int(np.nonzero(ref_anchor == a)[0][0]) for a in common]
        ei = [int(np.nonzero(eng_anchor == a)[0][0]) for a in common]
        det_box_err = np.concatenate([np.abs(d[ei, :4] - gd[ri, :4]).ravel(), np.abs(k[ei] - gk[ri]).ravel()])
    scale = max(h / 640.0, w / 640.0, 1.0) if det.input_size == (640, 640) else 1.0
    rep = dict(frame=[h, w], max_num=max_num, letterbox_magnification=float(1.0 / ds), candidates=int(len(ca)),
               head_box_px_max=float(box_err.max()), head_box_px_mean=float(box_err.mean()),
               head_kps_px_max=float(kps_err.max()), head_kps_px_mean=float(kps_err.mean()),
               head_score_abs_max=float(score_err.max()), head_score_abs_mean=float(score_err.mean()),
               reference_detections=len(gd), engine_detections=len(d), same_anchor_detections=int(len(common)),
               same_anchor_px_max=float(det_box_err.max(initial=0)),
               same_anchor_px_mean=float(det_box_err.mean()) if len(det_box_err) else 0.0)
    REPORT[f"detect_{tag}"] = rep
    _dump()
    print(tag, rep)
    mag = float(1.0 / ds)
    assert rep["head_box_px_max"] <= PX_MAX * mag and rep["head_kps_px_max"] <= PX_MAX * mag
    assert rep["head_box_px_mean"] <= PX_MEAN * mag and rep["head_kps_px_mean"] <= PX_MEAN * mag
    assert rep["head_score_abs_max"] <= SCORE_MAX
    assert rep["same_anchor_px_max"] <= PX_MAX * mag
    assert len(d) == len(gd) or (max_num == 0 and abs(len(d) - len(gd)) <= max(2, len(gd) // 10))
    assert len(common) >= 0.75 * len(gd)         # measured: 13/16 .. 50/50


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
    every engine face that kept the same anchor as a reference detection must come back as that detection's identity."""
    from scrfd_arcface_facerecognition_b200.gallery import Gallery
    det, rec = models("det_10g.onnx"), models("w600k_r50.onnx")
    img = inputs.frame(70, 1080, 1920)
    gd = gold["detect_10g_1080p_max16_det"]
    G = Gallery()
    G.add(np.concatenate([gold["arcface_r50_emb_detected"], inputs.embeddings(77, 4000) * 20]))
    d, k = det.detect(img, max_num=16)
    sc = engine_heads(det, img)[0]
    ref_anchor = anchors_of(gd, gold["detect_10g_1080p_max16_cand_score"], gold["detect_10g_1080p_max16_cand_anchor"])
    eng_anchor = anchors_of(d, sc)
    agree, sims, paired = 0, [], 0
    for ref_row, a in enumerate(ref_anchor):
        hit = np.nonzero(eng_anchor == a)[0]
        if len(hit) == 0:
            continue
        paired += 1
        idx, s = G.best_match(rec(img, k[hit[0]]), 0.4)
        agree += int(idx == ref_row)
        sims.append(s)
    REPORT["end_to_end"] = dict(reference_faces=len(gd), same_anchor_faces=paired, identity_agree=agree,
                                similarity_min=float(min(sims)) if sims else None)
    _dump()
    print(REPORT["end_to_end"])
    assert paired >= 10 and agree == paired
    # engine landmarks differ from the reference's by up to 0.2 px at 1080p and the frame is white noise, the worst case for
    # a bilinear crop: measured similarity 0.977 .. 0.9997 against background maxima of ~0.2
    assert min(sims) >= 0.96
