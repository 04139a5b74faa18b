"""TEST INFRASTRUCTURE -- not part of the product path.

CPU restatement (numpy, integer / float32 arithmetic spelled out) of every function on the
hot path of Kumar2421/scrfd_arcface_facerecognition.  It exists so the GPU box,
which has no /root/reference, still has a checker: tests/test_oracle.py pins every function below
against the reference's own Python executed verbatim (oracle/ref_loader.py) and against real cv2,
and the committed fixtures under tests/golden/ carry those results to the GPU box.

PARITY STATUS: the reference has no tests, fixtures or known-answer vectors of its own
(SURVEY.md section 4), so this oracle is pinned against *outputs of the reference run here*
(tests/golden/make_golden.py is the generating script), not against reference-owned vectors.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  Nothing here is fast and nothing here is shipped.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np

F32 = np.float32

# reference utils/helpers.py:6-15 (the single ArcFace 112x112 five-point template)
ARCFACE_TEMPLATE = np.array([[38.2946, 51.6963], [73.5318, 51.5014], [56.0252, 71.7366],
                             [41.5493, 92.3655], [70.7299, 92.2041]], dtype=np.float32)


# ---------------------------------------------------------------------------------------------
# a2 / a3: letterbox geometry, cv2.resize, blobFromImage        (reference models/scrfd.py:122-138, 76-82)
# ---------------------------------------------------------------------------------------------

def letterbox_geometry(img_h: int, img_w: int, in_w: int, in_h: int) -> Tuple[int, int, float]:
    """new_w, new_h, det_scale exactly as reference models/scrfd.py:123-134 (python float / int())."""
    im_ratio = float(img_h) / img_w
    model_ratio = in_h / in_w
    if im_ratio > model_ratio:
        new_h = in_h
        new_w = int(new_h / im_ratio)
    else:
        new_w = in_w
        new_h = int(new_w * im_ratio)
    return new_w, new_h, float(new_h) / img_h


def _linear_coeffs(dst: int, src: int) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Source index and 11-bit fixed-point weight pair per destination index (cv2 INTER_LINEAR, 8U)."""
    scale = 1.0 / (dst / src)                              # cv2 computes inv_scale then scale in double
    d = np.arange(dst, dtype=np.float64)
    f = ((d + 0.5) * scale - 0.5).astype(np.float32)
    s = np.floor(f).astype(np.int64)
    f = (f - s.astype(np.float32)).astype(np.float32)
    return s, f, scale


def resize_linear_u8(img: np.ndarray, new_w: int, new_h: int) -> np.ndarray:
    """cv2.resize(img, (new_w, new_h)) for uint8 HWC, default INTER_LINEAR, restated as integers.

    Follows OpenCV's 8-bit path: exact 2x decimation is promoted to the 2x2 area average;
    otherwise separable bilinear with 11-bit coefficients, horizontal pass into int32, vertical
    pass ((b0*(S0>>4))>>16 + (b1*(S1>>4))>>16 + 2)>>2.   (reference call site models/scrfd.py:135)
    """
    h, w = img.shape[:2]
    if (new_w, new_h) == (w, h):
        return img.copy()
    if w == 2 * new_w and h == 2 * new_h:
        a = img.astype(np.int32)
        return ((a[0::2, 0::2] + a[0::2, 1::2] + a[1::2, 0::2] + a[1::2, 1::2] + 2) >> 2).astype(np.uint8)
    sx, fx, _ = _linear_coeffs(new_w, w)
    sy, fy, _ = _linear_coeffs(new_h, h)
    # horizontal border handling: clamp and zero the fraction
    lo = sx < 0
    fx = np.where(lo, F32(0), fx)
    sx = np.where(lo, 0, sx)
    hi = sx >= w - 1
    fx = np.where(hi, F32(0), fx)
    sx = np.where(hi, w - 1, sx)
    a1 = np.rint(fx * F32(2048)).astype(np.int32)
    a0 = np.rint((F32(1) - fx) * F32(2048)).astype(np.int32)
    sx1 = np.minimum(sx + 1, w - 1)
    b1 = np.rint(fy * F32(2048)).astype(np.int32)
    b0 = np.rint((F32(1) - fy) * F32(2048)).astype(np.int32)
    y0 = np.clip(sy, 0, h - 1)
    y1 = np.clip(sy + 1, 0, h - 1)
    src = img.astype(np.int32)
    rows = src[:, sx] * a0[None, :, None] + src[:, sx1] * a1[None, :, None]       # (h, new_w, c)
    r0 = rows[y0]
    r1 = rows[y1]
    out = (((b0[:, None, None] * (r0 >> 4)) >> 16) + ((b1[:, None, None] * (r1 >> 4)) >> 16) + 2) >> 2
    return np.clip(out, 0, 255).astype(np.uint8)


def letterbox_u8(img: np.ndarray, in_w: int, in_h: int) -> Tuple[np.ndarray, float]:
    """Top-left zero letterbox (reference models/scrfd.py:135-138)."""
    new_w, new_h, det_scale = letterbox_geometry(img.shape[0], img.shape[1], in_w, in_h)
    canvas = np.zeros((in_h, in_w, 3), np.uint8)
    canvas[:new_h, :new_w] = resize_linear_u8(img, new_w, new_h)
    return canvas, det_scale


def blob_from_bgr(images_u8: np.ndarray, scale: float, mean: float) -> np.ndarray:
    """cv2.dnn.blobFromImage(s)(img, scale, size, (mean,)*3, swapRB=True) without resize:
    (float32(x) - mean) * float32(scale), BGR->RGB, HWC->NCHW.
    (reference models/scrfd.py:76-82 with scale=1/128; models/arcface.py:44-50 with scale=1/127.5)"""
    x = np.asarray(images_u8)
    if x.ndim == 3:
        x = x[None]
    x = x[..., ::-1].astype(np.float32)
    x = (x - F32(mean)) * F32(scale)
    return np.ascontiguousarray(x.transpose(0, 3, 1, 2))


# ---------------------------------------------------------------------------------------------
# a5-a11: anchor decode, threshold, sort, NMS, max_num        (reference models/scrfd.py:89-207,
#                                                               utils/helpers.py:62-107)
# ---------------------------------------------------------------------------------------------

def decode_level(scores: np.ndarray, bbox: np.ndarray, kps: np.ndarray, stride: int,
                 in_h: int, in_w: int, thr: float, num_anchors: int = 2):
    """One FPN level: returns (pos_scores (P,1), pos_boxes (P,4), pos_kps (P,5,2)) in anchor order."""
    hs, ws = in_h // stride, in_w // stride
    ys, xs = np.divmod(np.arange(hs * ws), ws)
    cx = np.repeat((xs * stride).astype(np.float32), num_anchors)            # scrfd.py:96-107
    cy = np.repeat((ys * stride).astype(np.float32), num_anchors)
    sc = np.asarray(scores, np.float32).reshape(-1)
    d = np.asarray(bbox, np.float32).reshape(-1, 4) * F32(stride)            # scrfd.py:92
    k = np.asarray(kps, np.float32).reshape(-1, 10) * F32(stride)            # scrfd.py:94
    pos = np.nonzero(sc >= thr)[0]                                           # scrfd.py:109
    boxes = np.stack([cx - d[:, 0], cy - d[:, 1], cx + d[:, 2], cy + d[:, 3]], axis=-1)   # helpers.py:62-83
    pts = np.empty((sc.shape[0], 10), np.float32)                            # helpers.py:86-107
    pts[:, 0::2] = cx[:, None] + k[:, 0::2]
    pts[:, 1::2] = cy[:, None] + k[:, 1::2]
    return sc[pos].reshape(-1, 1), boxes[pos], pts[pos].reshape(-1, 5, 2)


def nms_order(scores: np.ndarray) -> np.ndarray:
    """`scores.argsort()[::-1]` with numpy's sort pinned to kind='stable' (the reference uses the
    default, whose tie order is unspecified; SURVEY.md section 8c pins it this way)."""
    return np.argsort(scores, kind="stable")[::-1]


def nms(dets: np.ndarray, iou_thres: float) -> List[int]:
    """Greedy NMS with the '+1' pixel-area convention; keep j iff ovr <= thr, so a NaN overlap
    suppresses (reference models/scrfd.py:180-207).  float32 throughout, one rounding per op."""
    d = np.asarray(dets, np.float32)
    x1, y1, x2, y2 = d[:, 0], d[:, 1], d[:, 2], d[:, 3]
    one = F32(1)
    areas = (x2 - x1 + one) * (y2 - y1 + one)
    order = nms_order(d[:, 4])
    thr = F32(iou_thres)
    alive = np.ones(len(order), bool)
    keep: List[int] = []
    with np.errstate(invalid="ignore", divide="ignore"):
        for pos, i in enumerate(order):
            if not alive[pos]:
                continue
            keep.append(int(i))
            rest = order[pos + 1:]
            w = np.maximum(F32(0), np.minimum(x2[i], x2[rest]) - np.maximum(x1[i], x1[rest]) + one)
            h = np.maximum(F32(0), np.minimum(y2[i], y2[rest]) - np.maximum(y1[i], y1[rest]) + one)
            inter = w * h
            ovr = inter / (areas[i] + areas[rest] - inter)
            alive[pos + 1:] &= (ovr <= thr)
    return keep


def scrfd_postprocess(outputs: Sequence[np.ndarray], in_h: int, in_w: int, det_scale: float,
                      conf_thres: float, iou_thres: float, max_num: int = 0, metric: str = "max",
                      image_hw: Optional[Tuple[int, int]] = None,
                      strides: Sequence[int] = (8, 16, 32)):
    """Nine head tensors -> (det (N,5) f32, kps (N,5,2) f32)   (reference models/scrfd.py:89-119,142-177)."""
    fmc = len(strides)
    sl, bl, kl = [], [], []
    for i, s in enumerate(strides):
        a, b, c = decode_level(outputs[i], outputs[i + fmc], outputs[i + 2 * fmc], s, in_h, in_w, conf_thres)
        sl.append(a), bl.append(b), kl.append(c)
    scores = np.vstack(sl)
    order = nms_order(scores.ravel())
    boxes = (np.vstack(bl) / F32(det_scale)).astype(np.float32)
    kpss = (np.vstack(kl) / F32(det_scale)).astype(np.float32)
    pre = np.hstack((boxes, scores)).astype(np.float32)[order]
    keep = nms(pre, iou_thres)
    det = pre[keep]
    kpss = kpss[order][keep]
    if 0 < max_num < det.shape[0]:
        area = (det[:, 2] - det[:, 0]) * (det[:, 3] - det[:, 1])
        if metric == "max":
            values = area
        else:
            cy, cx = image_hw[0] // 2, image_hw[1] // 2
            ox = (det[:, 0] + det[:, 2]) / F32(2) - F32(cx)
            oy = (det[:, 1] + det[:, 3]) / F32(2) - F32(cy)
            values = area - (ox * ox + oy * oy) * F32(2)
        b = np.argsort(values, kind="stable")[::-1][:max_num]
        det, kpss = det[b], kpss[b]
    return det, kpss


# ---------------------------------------------------------------------------------------------
# a12 / a13: five-point similarity + warpAffine                 (reference utils/helpers.py:18-59)
# ---------------------------------------------------------------------------------------------

def estimate_norm_closed_form(landmark: np.ndarray, image_size: int = 112) -> np.ndarray:
    """2x3 float64 similarity mapping `landmark` onto the ArcFace template.  Closed form of the
    2-D Umeyama solution (rotation+uniform scale+translation; the reflection branch of the SVD
    form collapses to the same expression in 2-D).  (reference utils/helpers.py:18-53)"""
    src = np.asarray(landmark, np.float64)
    dst = ARCFACE_TEMPLATE.astype(np.float64)
    if image_size != 112:
        dst = (float(image_size) / 112 * ARCFACE_TEMPLATE).astype(np.float64)
    ms, md = src.mean(0), dst.mean(0)
    s, d = src - ms, dst - md
    a = (s[:, 0] * d[:, 0] + s[:, 1] * d[:, 1]).sum()
    b = (s[:, 0] * d[:, 1] - s[:, 1] * d[:, 0]).sum()
    v = (s * s).sum()
    p, q = a / v, b / v
    M = np.array([[p, -q, 0.0], [q, p, 0.0]])
    M[:, 2] = md - M[:, :2] @ ms
    return M


def warp_affine_u8(img: np.ndarray, M: np.ndarray, size: int = 112) -> np.ndarray:
    """cv2.warpAffine(img, M, (size,size), borderValue=0.0) for uint8 HWC, restated as integers:
    invert M in float64; 10-bit coordinate grid with 5 fractional interpolation bits; 4 taps with
    15-bit weights; constant zero border.  (reference utils/helpers.py:58)"""
    M = np.asarray(M, np.float64)
    h, w = img.shape[:2]
    D = M[0, 0] * M[1, 1] - M[0, 1] * M[1, 0]
    D = 1.0 / D if D != 0 else 0.0
    i00, i01 = M[1, 1] * D, -M[0, 1] * D
    i10, i11 = -M[1, 0] * D, M[0, 0] * D
    i02 = -i00 * M[0, 2] - i01 * M[1, 2]
    i12 = -i10 * M[0, 2] - i11 * M[1, 2]
    xs = np.arange(size, dtype=np.float64)
    adelta = np.rint(i00 * xs * 1024).astype(np.int64)
    bdelta = np.rint(i10 * xs * 1024).astype(np.int64)
    ys = np.arange(size, dtype=np.float64)
    X0 = np.rint((i01 * ys + i02) * 1024).astype(np.int64) + 16
    Y0 = np.rint((i11 * ys + i12) * 1024).astype(np.int64) + 16
    X = (X0[:, None] + adelta[None, :]) >> 5
    Y = (Y0[:, None] + bdelta[None, :]) >> 5
    sx, fx = X >> 5, X & 31
    sy, fy = Y >> 5, Y & 31
    # cv2 builds its 32x32 bilinear table in float and converts to short with rounding; for the
    # bilinear kernel the products (32-fx)(32-fy)*32 are already integers summing to 32768.
    w00 = (32 - fx) * (32 - fy) * 32
    w01 = fx * (32 - fy) * 32
    w10 = (32 - fx) * fy * 32
    w11 = fx * fy * 32
    src = img.astype(np.int64)

    def tap(yy, xx):
        ok = (yy >= 0) & (yy < h) & (xx >= 0) & (xx < w)
        v = src[np.clip(yy, 0, h - 1), np.clip(xx, 0, w - 1)]
        return v * ok[..., None]

    acc = (tap(sy, sx) * w00[..., None] + tap(sy, sx + 1) * w01[..., None]
           + tap(sy + 1, sx) * w10[..., None] + tap(sy + 1, sx + 1) * w11[..., None])
    return ((acc + 16384) >> 15).astype(np.uint8)


# ---------------------------------------------------------------------------------------------
# a17-a22: cosine similarity, best-match scan, gallery top-k, greedy duplicate merge
# ---------------------------------------------------------------------------------------------

def compute_similarity(f1: np.ndarray, f2: np.ndarray) -> np.float32:
    """dot / (|a| |b|) on ravelled float32 inputs (reference utils/helpers.py:110-123)."""
    a, b = np.asarray(f1).ravel(), np.asarray(f2).ravel()
    return np.dot(a, b) / (np.linalg.norm(a) * np.linalg.norm(b))


def best_match(embedding: np.ndarray, targets: np.ndarray, thresh: float) -> Tuple[int, float]:
    """Strict-greater scan over the target list, initial best 0, default 'Unknown' (= -1)
    (reference main.py:136-142)."""
    best, idx = 0.0, -1
    for t in range(len(targets)):
        s = compute_similarity(targets[t], embedding)
        if s > best and s > thresh:
            best, idx = s, t
    return idx, float(best)


def normalize_rows(x: np.ndarray) -> np.ndarray:
    """e / |e|   (reference duplicate.py:1491-1496; qdrant normalises at upsert for Cosine)."""
    x = np.asarray(x, np.float32)
    return x / np.linalg.norm(x, axis=-1, keepdims=True)


def search_similar(query: np.ndarray, gallery: np.ndarray, k: int, threshold: float):
    """Exact cosine top-k with score >= threshold, sorted descending (ties: lower row first).
    Contract of QdrantManager.search_similar (reference qdrant_manager.py:138-188) over a
    brute-force Cosine collection (config.json:99-100).  Returns (idx (<=k,), score (<=k,))."""
    q = normalize_rows(np.asarray(query, np.float32).reshape(1, -1))[0].astype(np.float64)
    g = normalize_rows(gallery).astype(np.float64)
    s = g @ q
    order = np.lexsort((np.arange(len(s)), -s))[:k]
    order = order[s[order] >= threshold]
    return order, s[order].astype(np.float32)


def merge_duplicates(emb: np.ndarray, thr: float) -> np.ndarray:
    """Greedy one-hop leader merge in ascending id order: leader[i] = lowest surviving j<=i with
    cos(j,i) >= thr at the time j is visited.  Semantics of find_and_merge_duplicates
    (reference duplicate.py:2726-2797).  Returns leader index per row."""
    g = normalize_rows(emb).astype(np.float64)
    n = len(g)
    leader = np.arange(n)
    alive = np.ones(n, bool)
    for i in range(n):
        if not alive[i]:
            continue
        s = g[i + 1:] @ g[i]
        hit = np.nonzero((s >= thr) & alive[i + 1:])[0] + i + 1
        leader[hit] = i
        alive[hit] = False
    return leader


def online_clusters(emb: np.ndarray, grouping_thr: float, search_thr: float = 0.0,
                    duplicate_thr: float = None) -> np.ndarray:
    """Online leader clustering in visit order: each embedding searches the persons created so far (one stored
    embedding per person = its first visit) and joins the BEST match when its similarity >= grouping_thr, else
    becomes a new person.  Semantics of the per-visit decision at reference duplicate.py:1853-1949 (JSON variant
    :2166-2262) over `search_person` (:1619-1643, cosine top-k with score >= search_thr, best first), made
    deterministic by processing visits in index order instead of thread-pool order.
    Ties between equally similar persons go to the earliest person.  Returns the person (leader index) of every row.
    `duplicate_thr` (config `duplicate_similarity_threshold`, 0.95): `is_duplicate_image` (duplicate.py:2618-2652)
    drops a visit whose best cosine to an existing person reaches it -- label -1, no visit stored.  That check only
    runs when the database already owns the `low_similarity_images` table (otherwise its query raises and the
    function returns False), hence None = off, the behaviour of a fresh database.
    Pinned by tests/golden/cluster_outputs.npz (the reference's own loop, tests/golden/make_cluster_golden.py)."""
    g = normalize_rows(emb).astype(np.float64)
    n = len(g)
    label = np.full(n, -1, np.int64)
    leaders: list = []
    for i in range(n):
        if leaders:
            full = g[leaders] @ g[i]
            if duplicate_thr is not None and full.max() >= duplicate_thr:
                continue                                # skipped as a duplicate image
            s = np.where(full >= search_thr, full, -np.inf)
            j = int(np.argmax(s))                       # first maximum = earliest person among equals
            if np.isfinite(s[j]) and s[j] >= grouping_thr:
                label[i] = leaders[j]
                continue
        leaders.append(i)
        label[i] = i
    return label


def online_similarities(emb: np.ndarray, label: np.ndarray, search_thr: float = 0.0) -> np.ndarray:
    """The similarity the reference records with each visit of the online loop (duplicate.py:1854-1855:
    `search_results[0]['similarity'] if search_results else 0.0`): for a visit that joined a person its cosine to that
    person, for a visit that founded one the best cosine to the persons existing before it (0.0 when the search
    returned nothing, i.e. none >= search_thr); the very first person is stored with 1.0 (duplicate.py:1826), a
    skipped visit (label -1) with nothing (0).  `label` as returned by `online_clusters`."""
    g = normalize_rows(emb).astype(np.float32)
    out = np.zeros(len(g), np.float32)
    leaders: list = []
    for i in range(len(g)):
        if label[i] < 0:
            continue
        if label[i] != i:
            out[i] = np.float32(g[i] @ g[label[i]])
            continue
        if leaders:
            s = g[leaders] @ g[i]
            s = s[s >= search_thr]
            out[i] = s.max() if len(s) else 0.0
        else:
            out[i] = 1.0
        leaders.append(i)
    return out
