"""CPU: pin the restated oracle (oracle/restate.py) against (a) the committed outputs of the reference
run verbatim, (b) the reference itself when its tree is present, (c) real cv2."""
import cv2
import numpy as np
import pytest

from oracle import restate, shims
from tests.golden import inputs
from tests.golden.make_golden import postprocess_cases


def _geometry(ih, iw, in_size):
    new_w, new_h, det_scale = restate.letterbox_geometry(ih, iw, in_size[0], in_size[1])
    return det_scale


@pytest.mark.parametrize("case", postprocess_cases(), ids=lambda c: c[0])
def test_postprocess_matches_reference_outputs(golden, case):
    name, (ih, iw), in_size, seed, ties, max_num, metric = case
    heads = inputs.head_tensors(seed, in_size[1], in_size[0], ties)
    det, kps = restate.scrfd_postprocess(heads, in_size[1], in_size[0], _geometry(ih, iw, in_size), 0.5, 0.4,
                                         max_num, metric, (ih, iw))
    assert det.dtype == np.float32 and kps.dtype == np.float32
    np.testing.assert_array_equal(det, golden[name + "_det"])       # bit-exact
    np.testing.assert_array_equal(kps, golden[name + "_kps"])


@pytest.mark.parametrize("case", postprocess_cases(), ids=lambda c: c[0])
def test_nms_keep_matches_reference_outputs(golden, case):
    name, (ih, iw), in_size, seed, ties, max_num, metric = case
    heads = inputs.head_tensors(seed, in_size[1], in_size[0], ties)
    ds = _geometry(ih, iw, in_size)
    sl, bl = [], []
    for i, s in enumerate((8, 16, 32)):
        a, b, _ = restate.decode_level(heads[i], heads[i + 3], heads[i + 6], s, in_size[1], in_size[0], 0.5)
        sl.append(a), bl.append(b)
    scores = np.vstack(sl)
    order = restate.nms_order(scores.ravel())
    pre = np.hstack((np.vstack(bl) / np.float32(ds), scores)).astype(np.float32)[order]
    assert len(order) == int(golden[name + "_ncand"])
    np.testing.assert_array_equal(np.asarray(restate.nms(pre, 0.4)), golden[name + "_keep"])


def test_postprocess_against_live_reference(ref):
    if ref is None:
        pytest.skip("reference tree not present on this box")
    from tests.golden.make_golden import _FakeSession, _ref_detector
    for seed in (40, 41):
        heads = inputs.head_tensors(seed, 640, 640, 0.0, score_mu=-3.5)
        det = _ref_detector(ref)
        det.session = _FakeSession(heads)
        d, k = det.detect(np.zeros((1080, 1920, 3), np.uint8), max_num=0)
        d2, k2 = restate.scrfd_postprocess(heads, 640, 640, 360 / 1080, 0.5, 0.4)
        np.testing.assert_array_equal(d, d2)
        np.testing.assert_array_equal(k, k2)


def test_empty_detection_shapes():
    heads = inputs.head_tensors(0, 320, 320, score_mu=-30.0)
    det, kps = restate.scrfd_postprocess(heads, 320, 320, 1.0, 0.5, 0.4)
    assert det.shape == (0, 5) and kps.shape == (0, 5, 2)


def test_nms_nan_overlap_suppresses():
    # two degenerate boxes whose union area is zero: ovr = 0/0 = NaN must suppress (np.where(ovr <= thr))
    d = np.array([[5, 5, 4, 4, 0.9], [5, 5, 4, 4, 0.8], [50, 50, 60, 60, 0.7]], np.float32)
    assert restate.nms(d, 0.4) == [0, 2]


@pytest.mark.parametrize("size", [(1080, 1920, 640, 360), (720, 1280, 640, 360), (480, 640, 640, 480),
                                  (500, 375, 480, 640), (333, 517, 640, 412), (64, 64, 640, 640),
                                  (777, 1333, 640, 373), (112, 112, 112, 112), (150, 130, 112, 112)])
def test_resize_restatement_matches_cv2(size):
    h, w, nw, nh = size
    img = inputs.frame(3, h, w)
    np.testing.assert_array_equal(cv2.resize(img, (nw, nh)), restate.resize_linear_u8(img, nw, nh))
    img = inputs.smooth_frame(4, max(h, 32), max(w, 32))[:h, :w]
    np.testing.assert_array_equal(cv2.resize(img, (nw, nh)), restate.resize_linear_u8(img, nw, nh))


def test_letterbox_pad_value_and_geometry():
    img = inputs.frame(5, 1080, 1920)
    canvas, ds = restate.letterbox_u8(img, 640, 640)
    assert ds == 360 / 1080
    np.testing.assert_array_equal(canvas[:360], img[1::3, 1::3])
    assert not canvas[360:].any()
    blob = restate.blob_from_bgr(canvas, 1 / 128, 127.5)
    assert blob[0, 0, 400, 0] == np.float32(-0.99609375)                   # pad rows are real input to the net


def test_blob_matches_cv2():
    img = inputs.frame(6, 64, 96)
    np.testing.assert_array_equal(cv2.dnn.blobFromImage(img, 1 / 128, (96, 64), (127.5,) * 3, swapRB=True),
                                  restate.blob_from_bgr(img, 1 / 128, 127.5))
    np.testing.assert_array_equal(cv2.dnn.blobFromImages([img, img], 1 / 127.5, (96, 64), (127.5,) * 3, swapRB=True),
                                  restate.blob_from_bgr(np.stack([img, img]), 1 / 127.5, 127.5))


def test_align_matches_reference_outputs(golden):
    for tag, (h, w) in (("1080p", (1080, 1920)), ("vga", (480, 640))):
        img = inputs.smooth_frame(7, h, w)
        lms = inputs.landmarks(8, h, w, 6)
        for i, lm in enumerate(lms):
            M = restate.estimate_norm_closed_form(lm)
            np.testing.assert_allclose(M, golden[f"align_{tag}_M"][i], rtol=0, atol=1e-9)
            # integer restatement of warpAffine, fed the reference's own M: bit-exact
            np.testing.assert_array_equal(restate.warp_affine_u8(img, golden[f"align_{tag}_M"][i]),
                                          golden[f"align_{tag}_crop"][i])
            # and cv2 itself agrees with the restatement for the closed-form M
            np.testing.assert_array_equal(restate.warp_affine_u8(img, M),
                                          cv2.warpAffine(img, M, (112, 112), borderValue=0.0))


def test_umeyama_closed_form_equals_svd_form():
    rng = np.random.default_rng(0)
    for _ in range(200):
        lm = rng.uniform(0, 500, (5, 2)).astype(np.float32)
        a = restate.estimate_norm_closed_form(lm)
        b = shims.umeyama(lm, restate.ARCFACE_TEMPLATE)[:2]
        assert np.abs(a - b).max() < 1e-8 * max(1.0, np.abs(b).max())


def test_similarity_and_best_match_match_reference_outputs(golden):
    gal = inputs.embeddings(11, 64)
    qs, ids = inputs.planted_queries(gal, 12, 16)
    sims = np.array([[restate.compute_similarity(t, q) for t in gal] for q in qs], np.float32)
    np.testing.assert_array_equal(sims, golden["sim_matrix"])
    best = [restate.best_match(q, gal, 0.4)[0] for q in qs]
    np.testing.assert_array_equal(best, golden["best_match"])
    np.testing.assert_array_equal(best, ids)                 # planted identities are recovered


def test_search_and_merge_semantics():
    gal = inputs.embeddings(13, 200)
    q = gal[17] * 3 + 0.1 * inputs.embeddings(14, 1)[0]
    idx, sc = restate.search_similar(q, gal, 5, 0.35)
    assert idx[0] == 17 and np.all(np.diff(sc) <= 0) and np.all(sc >= 0.35)
    # chain a-b-c with cos(a,b), cos(b,c) >= thr but cos(a,c) < thr: one hop only, c stays its own leader
    a = np.zeros(8, np.float32); a[0] = 1
    b = np.zeros(8, np.float32); b[0], b[1] = 0.8, 0.6
    c = np.zeros(8, np.float32); c[0], c[1] = 0.28, 0.96
    leader = restate.merge_duplicates(np.stack([a, b, c]), 0.8)
    assert list(leader) == [0, 0, 2]


def test_full_detect_oracle_path_matches_reference_outputs(golden):
    """restated pre/post-processing around the same torch-CPU net == the reference's detect()."""
    from oracle.torch_exec import TorchGraph
    from scrfd_arcface_facerecognition_b200 import archs
    tg = TorchGraph(archs.build_arch("scrfd_500m"))
    for fi, (h, w) in enumerate(((640, 640), (480, 640))):
        img = inputs.frame(20 + fi, h, w)
        canvas, ds = restate.letterbox_u8(img, 640, 640)
        out = tg.run(restate.blob_from_bgr(canvas, 1 / 128, 127.5))
        det, kps = restate.scrfd_postprocess([out[n] for n in tg.output_names], 640, 640, ds, 0.5, 0.4)
        np.testing.assert_array_equal(det, golden[f"detect_500m_{fi}_det"])
        np.testing.assert_array_equal(kps, golden[f"detect_500m_{fi}_kps"])


def test_online_clusters_semantics():
    """reference duplicate.py:1853-1949: join the best earlier person at >= grouping threshold, else found a new one"""
    a = np.zeros(8, np.float32); a[0] = 1
    b = np.zeros(8, np.float32); b[0], b[1] = 0.8, 0.6          # cos(a,b) = 0.8
    c = np.zeros(8, np.float32); c[0], c[1] = 0.28, 0.96        # cos(b,c) = 0.8, cos(a,c) = 0.28
    d = np.zeros(8, np.float32); d[0], d[1] = 0.6, 0.8          # cos(a,d) = 0.6, cos(c,d) = 0.936
    lab = restate.online_clusters(np.stack([a, b, c, d]), 0.8)
    assert list(lab) == [0, 0, 2, 2]                              # b joins a; c founds (only b is close, b is no person); d joins c
    emb = inputs.clustered(31, 40, 3)
    lab = restate.online_clusters(emb, 0.8)
    founders = restate.merge_duplicates(emb, 0.8) == np.arange(len(emb))
    assert ((lab == np.arange(len(emb))) == founders).all()      # same founders as the greedy merge
    assert len(np.unique(lab)) == 40


# ---------------------------------------------------------------------------------------------
# a19-a21 pinned to the reference's own clustering code (tests/golden/make_cluster_golden.py ran duplicate.py and
# qdrant_manager.py verbatim over oracle/fakes.py; outputs in tests/golden/cluster_outputs.npz)
# ---------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def cluster_golden():
    import os
    return dict(np.load(os.path.join(os.path.dirname(__file__), "golden", "cluster_outputs.npz")))


@pytest.mark.parametrize("case", inputs.cluster_cases(), ids=lambda c: c[0])
def test_clustering_restatement_matches_reference_run(cluster_golden, case):
    name, rows = case
    g = cluster_golden
    grouping, search, duplicate, merge = g["thresholds"]
    label = restate.online_clusters(rows, grouping, search)
    np.testing.assert_array_equal(label, g[f"online_{name}_label"])                      # a20, fresh database
    np.testing.assert_allclose(restate.online_similarities(rows, label, search), g[f"online_{name}_sim"], rtol=0, atol=1e-6)
    label = restate.online_clusters(rows, grouping, search, duplicate)
    np.testing.assert_array_equal(label, g[f"online_dup_{name}_label"])                  # a20 with the 0.95 duplicate gate
    np.testing.assert_allclose(restate.online_similarities(rows, label, search), g[f"online_dup_{name}_sim"], rtol=0, atol=1e-6)
    np.testing.assert_array_equal(restate.merge_duplicates(rows, merge), g[f"merge_{name}_leader"])   # a21
    qs, _ = inputs.planted_queries(rows, 90, 12, noise=0.8)
    for qi, q in enumerate(qs):                                                          # a19: k = 5, score >= 0.35
        idx, sc = restate.search_similar(q, rows, 5, search)
        gi, gs = g[f"search_{name}_idx"][qi], g[f"search_{name}_score"][qi]
        np.testing.assert_array_equal(gi[:len(idx)], idx)
        assert (gi[len(idx):] < 0).all()
        np.testing.assert_allclose(gs[:len(idx)], sc, rtol=0, atol=1e-6)


def test_cluster_golden_exercises_the_hard_cases(cluster_golden):
    """The fixture is only worth its name if greedy one-hop != transitive closure somewhere and the duplicate gate fires."""
    g = cluster_golden
    n = len(g["merge_chains_tight_leader"])
    survivors = int((g["merge_chains_tight_leader"] == np.arange(n)).sum())
    assert 30 < survivors < n                                                            # 30 chains; closure would leave 30
    assert int((g["online_dup_mixed_label"] < 0).sum()) > 50 and int((g["online_mixed_label"] < 0).sum()) == 0


def test_fake_qdrant_semantics():
    """The restated qdrant local mode: cosine normalises at upsert, `>=` threshold, descending order, upsert replaces."""
    from oracle import fakes
    c = fakes.QdrantClient(":memory:")
    c.create_collection("x", fakes.VectorParams(4, fakes.Distance.COSINE))
    c.upsert("x", [fakes.PointStruct(1, [2, 0, 0, 0], {"name": "a"}), fakes.PointStruct(2, [1, 1, 0, 0], {"name": "b"}),
                   fakes.PointStruct(3, [0, 0, 3, 0], {"name": "c"})])
    r = c.search("x", [5, 0, 0, 0], limit=10, score_threshold=float(np.float32(1 / np.sqrt(2))) - 1e-7)
    assert [p.id for p in r] == [1, 2] and abs(r[0].score - 1.0) < 1e-7
    assert np.allclose(c.retrieve("x", [2], with_vectors=True)[0].vector, [2 ** -0.5, 2 ** -0.5, 0, 0])
    c.upsert("x", [fakes.PointStruct(2, [0, 0, 0, 1], {"name": "b2"})])
    assert c.get_collection("x").points_count == 3 and [p.id for p in c.search("x", [5, 0, 0, 0], limit=1)] == [1]
    c.delete("x", fakes.PointIdsList([1]))
    assert c.get_collection("x").points_count == 2
    c.delete("x", fakes.FilterSelector(fakes.Filter()))
    assert c.get_collection("x").points_count == 0 and c.search("x", [1, 0, 0, 0], limit=3) == []


def test_port_equals_reference_run_on_headline_detector():
    """bench.py's CPU arm on the GPU box is the restated port (no reference tree there): on SCRFD-10G it must return
    exactly what the reference's own detect() returned here (tests/golden/headline_outputs.npz)."""
    import os
    from oracle.torch_exec import TorchGraph
    from scrfd_arcface_facerecognition_b200 import archs
    g = dict(np.load(os.path.join(os.path.dirname(__file__), "golden", "headline_outputs.npz")))
    tg = TorchGraph(archs.build_arch("scrfd_10g"))
    for tag, seed, hw, max_num, metric in (("10g_640_all", 72, (640, 640), 0, "max"), ("10g_1080p_max16", 70, (1080, 1920), 16, "max")):
        img = inputs.frame(seed, *hw)
        canvas, ds = restate.letterbox_u8(img, 640, 640)
        out = tg.run(restate.blob_from_bgr(canvas, 1 / 128, 127.5))
        det, kps = restate.scrfd_postprocess([out[n] for n in tg.output_names], 640, 640, ds, 0.5, 0.4, max_num, metric, hw)
        np.testing.assert_array_equal(det, g[f"detect_{tag}_det"])
        np.testing.assert_array_equal(kps, g[f"detect_{tag}_kps"])
