"""GPU parity tests at the model / API level: the drop-in classes against the oracle and the committed
reference outputs.  Tolerances follow SURVEY.md section 8c: network outputs are tier T2 (fp16 operands vs the
fp32 oracle, measured and asserted), embeddings cosine >= 0.999, post-processing bit-exact given the
same head tensors (covered in test_gpu_kernels.py)."""
import numpy as np
import pytest
import torch

from oracle import restate
from oracle.torch_exec import TorchGraph
from scrfd_arcface_facerecognition_b200 import _lib, archs
from scrfd_arcface_facerecognition_b200.engine import NetEngine
from scrfd_arcface_facerecognition_b200.graph import compile_graph
from tests.golden import inputs

pytestmark = pytest.mark.gpu


def _cos(a, b):
    return (a * b).sum(-1) / np.linalg.norm(a, axis=-1) / np.linalg.norm(b, axis=-1)


def _engine_outputs(key, hw, frames_u8, mean, scale, dtype=None):
    g = archs.build_arch(key)
    eng = NetEngine(compile_graph(g, hw), dtype=dtype)
    n = len(frames_u8)
    x = eng.input_buffer(n)
    blob = restate.blob_from_bgr(frames_u8, scale, mean)                       # exact f32 NCHW RGB
    x.zero_()
    x[..., :3] = torch.from_numpy(blob).permute(0, 2, 3, 1).to(x.dtype).cuda()
    outs = eng.run(n)
    torch.cuda.synchronize()
    ref = TorchGraph(g).run(blob)
    return eng, outs, ref


@pytest.mark.parametrize("key", ["arcface_r50", "arcface_mbf"])
def test_embedder_matches_oracle(key):
    crops = np.stack([inputs.smooth_frame(40 + i, 112, 112) for i in range(3)] + [inputs.frame(44, 112, 112)])
    eng, outs, ref = _engine_outputs(key, (112, 112), crops, 127.5, 1 / 127.5)
    (name, got), = outs.items()
    got = got.reshape(4, -1)[:, :512].cpu().numpy()
    want = ref[name]
    cos = _cos(got, want)
    rel = np.abs(got - want).max() / np.abs(want).max()
    print(f"{key}: cosine {cos}, max rel err {rel:.2e}")
    assert cos.min() >= 0.999                                                    # north-star embedding bar
    assert rel <= 2e-2


@pytest.mark.parametrize("key", ["arcface_r50", "arcface_mbf"])
def test_embedder_bf16_activations_meet_the_cosine_bar(key):
    """B2F_DTYPE=bf16 path (bf16 operands / activations, fp32 accumulation): SURVEY 7.4 measured 0.99994 on the simulator"""
    crops = np.stack([inputs.smooth_frame(45 + i, 112, 112) for i in range(3)] + [inputs.frame(49, 112, 112)])
    eng, outs, ref = _engine_outputs(key, (112, 112), crops, 127.5, 1 / 127.5, dtype=1)
    (name, got), = outs.items()
    cos = _cos(got.reshape(4, -1)[:, :512].cpu().numpy(), ref[name])
    print(f"{key} bf16: cosine {cos}")
    assert cos.min() >= 0.999


@pytest.mark.parametrize("key,hw", [("scrfd_10g", (640, 640)), ("scrfd_2.5g", (320, 352)), ("scrfd_500m", (640, 640))])
def test_detector_heads_match_oracle(key, hw):
    frames = np.stack([inputs.frame(50, hw[0], hw[1]), inputs.smooth_frame(51, hw[0], hw[1])])
    eng, outs, ref = _engine_outputs(key, hw, frames, 127.5, 1 / 128)
    worst = {}
    for name, tname, c, off in eng.plan.outputs:
        got = outs[name].reshape(2, -1, c).cpu().numpy()
        want = ref[name].reshape(2, got.shape[1], c)       # (2*H*W*2, k) -> per frame, per pixel, 2 anchors x k
        worst[name] = float(np.abs(got - want).max())
    print(key, {k: f"{v:.2e}" for k, v in worst.items()})
    names = [o[0] for o in eng.plan.outputs]
    # scores are probabilities; bbox / kps are in stride units (x8..32 px): fp16 operands / activations vs the fp32 oracle.
    # Measured on B200: scores <= 2.2e-3, bbox / kps <= 6.0e-3 stride units (0.05-0.19 px at 640 x 640) -- asserted with 2x
    # headroom.  The error is spread over all ~57 layers: keeping only the last head convolutions (or even every
    # activation) in fp32 removes at most half of it, the weights' own 16-bit rounding is the rest.
    assert max(worst[n] for n in names[:3]) <= 5e-3
    assert max(worst[n] for n in names[3:]) <= 1.2e-2


def test_scrfd_api_matches_reference_detect(golden):
    from models import SCRFD
    det = SCRFD("weights/det_500m.onnx")
    assert det.input_size == (640, 640) and det.conf_thres == 0.5 and det.iou_thres == 0.4
    assert det._feat_stride_fpn == [8, 16, 32] and det._num_anchors == 2 and det.fmc == 3
    assert len(det.output_names) == 9 and len(det.input_names) == 1
    for fi, (h, w) in enumerate(((640, 640), (480, 640))):
        img = inputs.frame(20 + fi, h, w)
        d, k = det.detect(img, max_num=0)
        gd, gk = golden[f"detect_500m_{fi}_det"], golden[f"detect_500m_{fi}_kps"]
        assert d.dtype == np.float32 and k.dtype == np.float32 and d.shape[1] == 5 and k.shape[1:] == (5, 2)
        assert (np.diff(d[:, 4]) <= 0).all()                                     # descending score
        # tier T3: match detections by IoU against the reference's fp32 run
        matched = 0
        for box in gd:
            x1, y1 = np.maximum(box[0], d[:, 0]), np.maximum(box[1], d[:, 1])
            x2, y2 = np.minimum(box[2], d[:, 2]), np.minimum(box[3], d[:, 3])
            inter = np.clip(x2 - x1, 0, None) * np.clip(y2 - y1, 0, None)
            iou = inter / ((box[2] - box[0]) * (box[3] - box[1]) + (d[:, 2] - d[:, 0]) * (d[:, 3] - d[:, 1]) - inter)
            j = int(np.argmax(iou))
            if iou[j] > 0.9 and abs(d[j, 4] - box[4]) < 0.03:
                matched += 1
        print(f"frame {fi}: reference {len(gd)} detections, engine {len(d)}, matched {matched}")
        assert matched >= 0.9 * len(gd) and abs(len(d) - len(gd)) <= 0.1 * len(gd)
        # max_num keeps the largest faces, sorted by area
        d5, k5 = det.detect(img, max_num=5)
        area = (d5[:, 2] - d5[:, 0]) * (d5[:, 3] - d5[:, 1])
        assert len(d5) == 5 and (np.diff(area) <= 0).all()
    # forward(): per-stride lists in anchor order, unscaled
    canvas, _ = restate.letterbox_u8(inputs.frame(20, 640, 640), 640, 640)
    sl, bl, kl = det.forward(canvas, 0.5)
    assert len(sl) == 3 and all(s.shape[1] == 1 for s in sl) and all(k.shape[1:] == (5, 2) for k in kl)
    assert sum(len(s) for s in sl) >= len(golden["detect_500m_0_det"])
    # nms(): same keep list as the reference implementation on the same array
    pre = np.hstack((np.vstack(bl), np.vstack(sl))).astype(np.float32)
    pre = pre[pre[:, 4].argsort(kind="stable")[::-1]]
    assert det.nms(pre, 0.4) == restate.nms(pre, 0.4)
    # zero detections -> empty arrays, like the reference
    det.conf_thres = 1.1
    d, k = det.detect(inputs.frame(1, 240, 320))
    assert d.shape == (0, 5) and k.shape == (0, 5, 2)


def test_arcface_api_matches_reference_call(golden):
    from models import ArcFace
    rec = ArcFace("weights/w600k_mbf.onnx")
    assert rec.input_size == (112, 112) and rec.input_mean == 127.5 and rec.input_std == 127.5
    assert rec.taskname == "recognition" and len(rec.output_names) == 1
    img = inputs.smooth_frame(30, 480, 640)
    lms = inputs.landmarks(31, 480, 640, 4)
    embs = np.stack([rec(img, lm) for lm in lms])
    assert embs.shape == (4, 512) and embs.dtype == np.float32
    cos = _cos(embs, golden["arcface_mbf_emb"])
    print("ArcFace(mbf) cosine vs reference:", cos)
    assert cos.min() >= 0.999
    np.testing.assert_array_equal(rec.get_embedding(img, lms[0]), embs[0])
    # get_feat on aligned crops == __call__
    from utils.helpers import norm_crop_image, compute_similarity, estimate_norm
    crops = [norm_crop_image(img, lm) for lm in lms]
    feats = rec.get_feat(crops)
    assert feats.shape == (4, 512)
    assert _cos(feats, embs).min() >= 0.99999
    M, idx = estimate_norm(lms[0])
    assert M.shape == (2, 3) and M.dtype == np.float64 and idx == 0
    s = compute_similarity(embs[0], golden["arcface_mbf_emb"][0])
    assert isinstance(s, np.float32) and s >= 0.999


def test_helpers_decode_functions():
    from utils.helpers import distance2bbox, distance2kps
    rng = np.random.default_rng(0)
    pts = rng.uniform(0, 640, (100, 2)).astype(np.float32)
    d4 = rng.uniform(0, 50, (100, 4)).astype(np.float32)
    d10 = rng.normal(0, 10, (100, 10)).astype(np.float32)
    want = np.stack([pts[:, 0] - d4[:, 0], pts[:, 1] - d4[:, 1], pts[:, 0] + d4[:, 2], pts[:, 1] + d4[:, 3]], -1)
    np.testing.assert_array_equal(distance2bbox(pts, d4), want)
    want = d10.copy()
    want[:, 0::2] += pts[:, :1]
    want[:, 1::2] += pts[:, 1:]
    np.testing.assert_array_equal(distance2kps(pts, d10), want)


def test_pipeline_end_to_end_and_graph_replay():
    from models import SCRFD, ArcFace
    from scrfd_arcface_facerecognition_b200.gallery import Gallery
    from scrfd_arcface_facerecognition_b200.pipeline import FacePipeline
    det, rec = SCRFD("weights/det_500m.onnx"), ArcFace("weights/w600k_mbf.onnx")
    frames = np.stack([inputs.frame(60 + i, 480, 640) for i in range(3)])
    # enrol: the per-image reference path (build_targets, reference main.py:78-105)
    targets = []
    for f in frames:
        b, k = det.detect(f, max_num=1)
        targets.append(rec(f, k[0]))
    G = Gallery()
    G.add(np.stack(targets))
    pipe = FacePipeline(det, rec, G, max_num=4, similarity_thresh=0.4)
    out = pipe.process(torch.from_numpy(frames).cuda())
    torch.cuda.synchronize()
    counts = out["counts"].cpu().numpy()
    assert (counts[:, 0] == 4).all() and (counts[:, 3] == 0).all()
    # batched results == the single-image API, frame by frame
    for i, f in enumerate(frames):
        d, k = det.detect(f, max_num=4)
        np.testing.assert_array_equal(out["det"][i].cpu().numpy(), d)
        np.testing.assert_array_equal(out["kps"][i].cpu().numpy(), k)
        e = np.stack([rec(f, kk) for kk in k])
        assert _cos(out["emb"].reshape(3, 4, -1)[i].cpu().numpy(), e).min() >= 0.99999
    # the largest face of frame i was enrolled as target i
    assert (out["match_idx"][:, 0].cpu().numpy() == np.arange(3)).all()
    assert (out["match_score"][:, 0].cpu().numpy() > 0.99).all()
    # CUDA-graph replay gives the same answer
    static, gouts, graph, kernels = pipe.capture(3, 480, 640)
    static.copy_(torch.from_numpy(frames).cuda())
    graph.replay()
    torch.cuda.synchronize()
    assert kernels > 50
    np.testing.assert_array_equal(gouts["det"].cpu().numpy(), out["det"].cpu().numpy())
    np.testing.assert_array_equal(gouts["match_idx"].cpu().numpy(), out["match_idx"].cpu().numpy())
    assert _lib.launch_count() > 0
