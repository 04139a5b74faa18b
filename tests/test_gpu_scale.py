"""BASELINE.json configs 3-5 at (or near) their full sizes, checked through size-independent properties, plus the
SURVEY section 8(f) rows (FaceAnalysis facade, QdrantManager-shaped store).  Everything here needs a B200."""
import numpy as np
import pytest
import torch

from tests.golden import inputs

pytestmark = pytest.mark.gpu


def _unit(x):
    return x / x.norm(dim=1, keepdim=True)


# ---------------------------------------------------------------------------------------------------------
# config 3: top-1 of 100k queries against a 1M x 512 gallery, row-sharded 8 ways
# ---------------------------------------------------------------------------------------------------------
def test_config3_full_size_top1_and_eight_way_sharding():
    from scrfd_arcface_facerecognition_b200.gallery import Gallery, merge_shard_topk, shard_range
    g_rows, q_rows, shards = 1_000_000, 100_000, 8
    gen = torch.Generator(device="cuda").manual_seed(3)
    gal = torch.randn((g_rows, 512), generator=gen, device="cuda")
    ids = torch.randperm(g_rows, generator=gen, device="cuda")[:q_rows]
    noise = _unit(torch.randn((q_rows, 512), generator=gen, device="cuda"))
    queries = (_unit(gal[ids]) + 0.8 * noise) * (5 + 20 * torch.rand((q_rows, 1), generator=gen, device="cuda"))
    del noise
    G = Gallery()
    G.set_shard(gal, 0)
    s_full, i_full = [], []
    for lo in range(0, q_rows, 25_000):
        s, i = G.match(queries[lo:lo + 25_000], 1, 0.4, strict=True)
        s_full.append(s.clone()), i_full.append(i.clone())
    s_full, i_full = torch.cat(s_full), torch.cat(i_full)
    # planted identities are the ground truth: cos(query, planted row) ~ 0.78, every other row < 0.3
    assert (i_full[:, 0] == ids).all(), f"{int((i_full[:, 0] != ids).sum())} of {q_rows} top-1 identities wrong"
    exact = (_unit(queries[:4096]) * _unit(gal[ids[:4096]])).sum(1)
    assert (s_full[:4096, 0] - exact).abs().max().item() <= 2e-6          # fp32 re-score of the winner
    del G
    # the same answer from 8 row shards merged by (score desc, index asc) -- what the 8-GPU run exchanges
    part_s, part_i = [], []
    for r in range(shards):
        b, e = shard_range(g_rows, r, shards)
        Gs = Gallery(rank=r, world_size=shards)
        Gs.set_shard(gal[b:e], b)
        ss, ii = [], []
        for lo in range(0, q_rows, 25_000):
            s, i = Gs.match_local(queries[lo:lo + 25_000], 1, 0.4, strict=True)
            ss.append(s.clone()), ii.append(i.clone())
        part_s.append(torch.cat(ss)), part_i.append(torch.cat(ii))
        del Gs
    ms, mi = merge_shard_topk(torch.stack(part_s), torch.stack(part_i), 1)
    assert torch.equal(mi, i_full)
    assert torch.equal(ms, s_full)


def test_config3_embedding_is_batch_invariant():
    """A crop's embedding must not depend on which other crops share its batch (crops are sharded across GPUs)."""
    from models import ArcFace
    rec = ArcFace("weights/w600k_r50.onnx")
    frames = torch.from_numpy(np.stack([inputs.frame(70 + i, 360, 480) for i in range(4)])).cuda()
    n = 1500
    kps = torch.from_numpy(inputs.landmarks(71, 360, 480, n).reshape(n, 10)).cuda()
    fidx = (torch.arange(n, device="cuda") % 4).to(torch.int32)
    big = rec.embed_batch(frames, fidx, kps).clone()
    for lo, hi in ((0, 1), (5, 133), (1000, 1500)):
        part = rec.embed_batch(frames, fidx[lo:hi].contiguous(), kps[lo:hi].contiguous())
        assert torch.equal(part, big[lo:hi])
    assert torch.isfinite(big).all() and big.norm(dim=1).min().item() > 0


def test_stem8_path_equals_patch_path():
    """the first ArcFace convolution in its 8-channel stem form (crop kept as 16-byte pixels, one TMA box per filter tap)
    against the 32-channel patch tensor path: same embeddings up to the order of the fp32 accumulation of the first
    layer (cosine >= 0.99999, max abs difference <= 2e-3 of the embedding scale), also for MobileFaceNet"""
    from models import ArcFace
    frames = torch.from_numpy(np.stack([inputs.frame(80 + i, 360, 480) for i in range(3)])).cuda()
    n = 70
    kps = torch.from_numpy(inputs.landmarks(81, 360, 480, n).reshape(n, 10)).cuda()
    fidx = (torch.arange(n, device="cuda") % 3).to(torch.int32)
    for path in ("weights/w600k_r50.onnx", "weights/w600k_mbf.onnx"):
        rec = ArcFace(path)
        if rec._engine.stem8(n) is None:                  # a plan that does not open with patches + 1x1
            assert "r50" not in path
            continue
        rec.stem8 = True
        a = rec.embed_batch(frames, fidx, kps).clone()
        rec.stem8 = False
        b = rec.embed_batch(frames, fidx, kps).clone()
        cos = torch.nn.functional.cosine_similarity(a, b, dim=1)
        assert cos.min().item() >= 0.99999, cos.min().item()
        assert (a - b).abs().max().item() <= 2e-3 * b.abs().max().item()


# ---------------------------------------------------------------------------------------------------------
# config 4: all-pairs cosine clustering of 200k embeddings, upper triangle block-partitioned over ranks
# ---------------------------------------------------------------------------------------------------------
def test_config4_full_size_clustering_and_block_partition():
    from scrfd_arcface_facerecognition_b200.gallery import Gallery, row_blocks
    centres, members = 50_000, 4
    n = centres * members
    gen = torch.Generator(device="cuda").manual_seed(4)
    c = _unit(torch.randn((centres, 512), generator=gen, device="cuda"))
    x = c.repeat_interleave(members, 0) + 0.25 * _unit(torch.randn((n, 512), generator=gen, device="cuda"))
    perm = torch.randperm(n, generator=gen, device="cuda")
    x = x[perm]
    cluster = (perm // members)
    G = Gallery()
    G.add(x)
    leader = G.merge_duplicates(0.8)
    # within a cluster cos ~ 0.94, across clusters < 0.35: every row's leader is the lowest index of its cluster
    first = torch.full((centres,), n, dtype=torch.int64, device="cuda")
    first.scatter_reduce_(0, cluster, torch.arange(n, device="cuda"), reduce="amin")
    np.testing.assert_array_equal(leader, first[cluster].cpu().numpy())
    # block partition: rank r owns the row blocks dealt to it; union of the pair lists == one pass
    all_pairs = G.duplicate_pairs(0.8)
    assert all_pairs.numel() == centres * members * (members - 1) // 2
    for world in (2, 8):
        per_rank = [[] for _ in range(world)]
        for r, b, e in row_blocks(n, world):
            per_rank[r].append(G.duplicate_pairs(0.8, b, e))
        sizes = [sum(p.numel() for p in ps) for ps in per_rank]
        merged = torch.sort(torch.cat([p for ps in per_rank for p in ps])).values
        assert torch.equal(merged, all_pairs)
        assert max(sizes) <= 1.2 * (sum(sizes) / world) + 64              # cyclic dealing balances the triangle
    np.testing.assert_array_equal(G.resolve_pairs(all_pairs), leader)


# ---------------------------------------------------------------------------------------------------------
# config 5: SCRFD-2.5G + R50 on 1080p frames, max_num = 50, frames sharded round-robin
# ---------------------------------------------------------------------------------------------------------
def test_config5_frame_sharding_is_invisible():
    from models import SCRFD, ArcFace
    from scrfd_arcface_facerecognition_b200.pipeline import FacePipeline
    det, rec = SCRFD("weights/det_2.5g.onnx"), ArcFace("weights/w600k_r50.onnx")
    frames = torch.from_numpy(np.stack([inputs.frame(80 + i, 1080, 1920) for i in range(6)])).cuda()
    pipe = FacePipeline(det, rec, None, max_num=50)
    whole = {k: v.clone() for k, v in pipe.process(frames).items()}
    assert whole["det"].shape == (6, 50, 5) and whole["emb"].shape == (300, 512)
    for world in (2, 3):
        for r in range(world):                                # frame i -> rank i mod world
            mine = pipe.process(frames[r::world].contiguous())
            assert torch.equal(mine["det"], whole["det"][r::world])
            assert torch.equal(mine["kps"], whole["kps"][r::world])
            assert torch.equal(mine["counts"], whole["counts"][r::world])
            assert torch.equal(mine["emb"].reshape(-1, 50, 512), whole["emb"].reshape(6, 50, 512)[r::world])
    # single-image API on one frame == its row of the batch (reference SCRFD.detect, max_num = 50)
    d, k = det.detect(frames[4].cpu().numpy(), max_num=50)
    cnt = int(whole["counts"][4, 0].item())
    np.testing.assert_array_equal(whole["det"][4, :cnt].cpu().numpy(), d)
    np.testing.assert_array_equal(whole["kps"][4, :cnt].cpu().numpy(), k)


# ---------------------------------------------------------------------------------------------------------
# section 8(f): FaceAnalysis facade and QdrantManager surface
# ---------------------------------------------------------------------------------------------------------
def test_face_analysis_facade_matches_detect_plus_arcface():
    from models import SCRFD, ArcFace
    from scrfd_arcface_facerecognition_b200.face_analysis import FaceAnalysis
    app = FaceAnalysis(name="buffalo_s")
    with pytest.raises(RuntimeError):
        app.get(inputs.frame(90, 480, 640))
    app.prepare(ctx_id=0, det_size=(640, 640))
    img = inputs.frame(90, 480, 640)
    faces = app.get(img, max_num=5)
    det, rec = SCRFD("weights/det_500m.onnx"), ArcFace("weights/w600k_mbf.onnx")
    boxes, kpss = det.detect(img, max_num=5, metric="default")
    assert len(faces) == len(boxes) == 5
    for f, b, k in zip(faces, boxes, kpss):
        np.testing.assert_array_equal(f.bbox, b[:4])
        np.testing.assert_array_equal(f.kps, k)
        assert f.det_score == float(b[4]) and f["det_score"] == f.det_score
        np.testing.assert_array_equal(f.embedding, rec(img, k))
        np.testing.assert_allclose(np.linalg.norm(f.normed_embedding), 1.0, atol=1e-6)
        assert f.gender is None                                              # absent attributes read as None
    best = max(faces, key=lambda f: getattr(f, "det_score", 0.0))           # duplicate.py:1479
    assert best.det_score == max(float(b[4]) for b in boxes)
    assert FaceAnalysis(name="buffalo_l")._det_path.endswith("det_10g.onnx")


def test_shared_models_are_safe_across_threads():
    """reference duplicate.py:1954 runs the shared FaceAnalysis object from a 4-thread pool: results handed to one
    thread must not be overwritten by another thread's run on the same engine buffers (they are copied out under the
    model lock).  Four threads, different images with the same face count, many rounds == the sequential answers."""
    from concurrent.futures import ThreadPoolExecutor
    from scrfd_arcface_facerecognition_b200.face_analysis import FaceAnalysis
    app = FaceAnalysis(name="buffalo_s")
    app.prepare(ctx_id=0, det_size=(640, 640))
    imgs = [inputs.frame(120 + i, 480, 640) for i in range(4)]
    want = [[(f.bbox.copy(), f.embedding.copy()) for f in app.get(im, max_num=3)] for im in imgs]
    assert all(len(w) == 3 for w in want)

    def work(i):
        out = []
        for _ in range(12):
            out.append([(f.bbox, f.embedding) for f in app.get(imgs[i], max_num=3)])
        return out
    with ThreadPoolExecutor(max_workers=4) as ex:
        results = list(ex.map(work, range(4)))
    for i, rounds in enumerate(results):
        for faces in rounds:
            assert len(faces) == 3
            for (b, e), (wb, we) in zip(faces, want[i]):
                np.testing.assert_array_equal(b, wb)
                np.testing.assert_array_equal(e, we)
    # the batched entries hand out copies by default: a second call must not change the first call's tensors
    rec, det = app.rec_model, app.det_model
    frames = torch.from_numpy(np.stack(imgs[:2])).cuda()
    d1, k1, c1 = det.detect_batch(frames, max_num=3)
    keep = (d1.clone(), k1.clone())
    det.detect_batch(torch.from_numpy(np.stack(imgs[2:])).cuda(), max_num=3)
    torch.cuda.synchronize()
    assert torch.equal(d1, keep[0]) and torch.equal(k1, keep[1])
    fidx = torch.zeros(3, dtype=torch.int32, device="cuda")
    e1 = rec.embed_batch(frames, fidx, k1[0].reshape(3, 10))
    e1_keep = e1.clone()
    rec.embed_batch(frames, fidx + 1, k1[1].reshape(3, 10))
    torch.cuda.synchronize()
    assert torch.equal(e1, e1_keep)


def test_engine_buffers_are_bucketed_by_capacity():
    """a service that sees every face count 1..K must not keep one activation set per count: buffers are shared by
    all batch sizes of one power-of-two capacity, and results do not depend on the capacity they ran in"""
    from models import ArcFace
    rec = ArcFace("weights/w600k_mbf.onnx")
    frames = torch.from_numpy(np.stack([inputs.frame(130, 360, 480)])).cuda()
    kps = torch.from_numpy(inputs.landmarks(131, 360, 480, 40).reshape(40, 10)).cuda()
    fidx = torch.zeros(40, dtype=torch.int32, device="cuda")
    full = rec.embed_batch(frames, fidx, kps)
    for n in range(1, 41):
        part = rec.embed_batch(frames, fidx[:n].contiguous(), kps[:n].contiguous())
        assert torch.equal(part, full[:n])
    eng = rec._engine
    assert sorted(eng._pools) == [1, 2, 4, 8, 16, 32, 64]
    one_set = sum(b.numel() for b in eng._pools[64])
    assert eng.buffer_bytes() <= 2.2 * one_set                      # geometric: at most ~2x the largest capacity


def test_qdrant_manager_surface():
    from oracle import restate
    from scrfd_arcface_facerecognition_b200.vector_store import QdrantManager
    qm = QdrantManager({"vector_database": {"mode": "memory", "collection_name": "faces", "vector_size": 512}})
    emb = inputs.embeddings(21, 300)
    for i, e in enumerate(emb):
        assert qm.add_embedding(1000 + i, e, {"name": f"p{i}", "quality": 0.5 + i * 1e-3})
    assert qm.get_embedding_count() == 300
    assert not qm.add_embedding(7, np.zeros(100, np.float32), {})               # wrong size -> False, nothing stored
    assert qm.search_similar(np.zeros(100, np.float32)) == []
    qs, ids = inputs.planted_queries(emb, 22, 8)
    for q, pid in zip(qs, ids):
        res = qm.search_similar(q, k=5, threshold=0.3)
        idx, sc = restate.search_similar(q, emb, 5, 0.3)
        assert [r["person_id"] for r in res] == [1000 + int(i) for i in idx]
        np.testing.assert_allclose([r["similarity"] for r in res], sc, atol=2e-6)
        assert res[0]["name"] == f"p{pid}" and res[0]["metadata"]["person_id"] == 1000 + pid
    # get_embedding returns the stored (unit-norm, as Qdrant's Cosine collections keep it) vector
    v = qm.get_embedding(1005)
    np.testing.assert_allclose(v, emb[5] / np.linalg.norm(emb[5]), atol=1e-6)
    assert qm.get_embedding(5) is None
    # upsert replaces in place; delete removes and keeps the others searchable
    assert qm.update_embedding(1005, emb[6], {"name": "moved"}) and qm.get_embedding_count() == 300
    res = qm.search_similar(emb[6], k=2, threshold=0.9)
    assert [r["person_id"] for r in res] == [1005, 1006] and res[0]["name"] == "moved"   # equal scores: insertion order
    assert qm.delete_embedding(1005) and qm.delete_embedding(424242)
    assert qm.get_embedding_count() == 299 and qm.get_embedding(1005) is None
    res = qm.search_similar(emb[6], k=2, threshold=0.9)
    assert [r["person_id"] for r in res] == [1006]
    res = qm.search_similar(emb[299], k=1)
    assert res[0]["person_id"] == 1299
    # k beyond the kernel's running top-8 (duplicate.py searches with k = collection size)
    res = qm.search_similar(qs[0], k=299, threshold=-1.0)
    assert len(res) == 299 and all(a["similarity"] >= b["similarity"] for a, b in zip(res, res[1:]))
    idx, sc = restate.search_similar(qs[0], np.delete(emb, 5, 0), 8, -1.0)
    np.testing.assert_allclose([r["similarity"] for r in res[:8]], sc, atol=2e-6)
    # duplicate leaders over the collection (duplicate.py:2726-2797)
    qm.clear_all()
    assert qm.get_embedding_count() == 0 and qm.search_similar(qs[0]) == []
    cl = inputs.clustered(23, 20, 3)
    for i, e in enumerate(cl):
        qm.add_embedding(f"id{i}", e, {})
    want = restate.merge_duplicates(cl, 0.8)
    got = qm.find_duplicate_leaders(0.8)
    assert got == {f"id{i}": f"id{int(l)}" for i, l in enumerate(want) if int(l) != i}
    assert qm.get_collection_info()["points_count"] == 60


def test_video_runner_matches_the_per_frame_reference_loop():
    """section 8(f) rank 3: batched, double-buffered video loop == reference frame_processor (main.py:108-150) per frame"""
    from models import SCRFD, ArcFace
    from oracle import restate
    from scrfd_arcface_facerecognition_b200.video import FrameFeeder, VideoRunner
    det, rec = SCRFD("weights/det_500m.onnx"), ArcFace("weights/w600k_mbf.onnx")
    frames = [inputs.frame(120 + i, 480, 640) for i in range(10)]

    class FakeCapture:                                   # cv2.VideoCapture surface: read() -> (ok, frame)
        def __init__(self, fr):
            self.fr, self.i = fr, 0

        def read(self):
            self.i += 1
            return (True, self.fr[self.i - 1].copy()) if self.i <= len(self.fr) else (False, None)

    runner = VideoRunner(det, rec, max_num=3, similarity_thresh=0.4, batch=4)
    names = runner.enroll([(frames[i], f"person{i}") for i in range(3)])
    assert names == ["person0", "person1", "person2"]
    targets = []
    for i in range(3):                                    # reference build_targets (main.py:78-105)
        _, k = det.detect(frames[i], max_num=1)
        targets.append(rec(frames[i], k[0]))
    seen = []
    got = runner.run(FakeCapture(frames), on_frame=lambda fr, faces: seen.append(len(faces)))
    assert len(got) == 10 and seen == [len(g) for g in got]
    for f, faces in zip(frames, got):
        boxes, kpss = det.detect(f, max_num=3)
        assert len(faces) == len(boxes)
        for (bbox, name, sim), b, k in zip(faces, boxes, kpss):
            np.testing.assert_array_equal(bbox, b[:4].astype(np.int32))
            idx, best = restate.best_match(rec(f, k), np.stack(targets), 0.4)     # strict '>' scan, main.py:136-142
            assert name == (f"person{idx}" if idx >= 0 else "Unknown")
            assert abs(sim - best) <= 1e-5
    assert got[0][0][1] == "person0" and got[1][0][1] == "person1" and got[2][0][1] == "person2"
    # drawing touches the host frames like the reference's draw_bbox / draw_bbox_info
    drawn = [f.copy() for f in frames[:4]]
    runner.run(drawn, draw=True)
    assert any((d != f).any() for d, f in zip(drawn, frames[:4]))
    # ... and the overlay painted on the device batch (one kernel per batch) is the same image, byte for byte
    drawn_gpu = [f.copy() for f in frames[:4]]
    runner.run(drawn_gpu, draw="gpu")
    for a, b in zip(drawn, drawn_gpu):
        np.testing.assert_array_equal(a, b)
    # the feeder alone: batches of 4, 4, 2 with the frames intact
    sizes = []
    for dev_batch, n, kept in FrameFeeder(frames, 4):
        sizes.append(n)
        torch.cuda.current_stream().synchronize()
        np.testing.assert_array_equal(dev_batch[:n].cpu().numpy(), np.stack(kept))
    assert sizes == [4, 4, 2]
