"""Generate tests/golden/cluster_outputs.npz by running the REFERENCE's clustering code verbatim.

Run in the build container only (needs /root/reference or $B2F_REFERENCE):
    python tests/golden/make_cluster_golden.py
`duplicate.py` and `qdrant_manager.py` are imported unmodified from the reference tree; the wheels they need and
this image lacks (`qdrant_client`, `insightface`) are replaced by oracle/fakes.py.  What runs verbatim:
  * the per-visit online decision of `SmartFaceRecognition.process_visit_data` (duplicate.py:1721-2005; decision
    :1853-1949) with `is_duplicate_image` (:2618-2652), `search_person` (:1619-1643), `add_person` (:1531-1602),
    `store_visit_info` (:1657-1675) on a temp SQLite file, over `QdrantManager` (qdrant_manager.py:91-188);
  * `find_and_merge_duplicates` (duplicate.py:2726-2797) with `merge_duplicate_persons` (:2679-2724);
  * `QdrantManager.search_similar` (qdrant_manager.py:138-188).
The only substitution on the instance is `extract_face_embedding` (image download + insightface model, outside rows
a19-a21): it returns the seeded synthetic embedding of the visit.  `image_processing.max_workers` is set to 1 in the
config so that visits are processed in index order (the reference's thread-pool order is nondeterministic).
Inputs are seeded (tests/golden/inputs.py `cluster_cases`); only outputs are stored.
"""
from __future__ import annotations

import json
import logging
import os
import sqlite3
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import fakes, ref_loader  # noqa: E402
from tests.golden import inputs  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def load_reference_app(work: str):
    """Import the reference's duplicate.py with cwd = a scratch directory that has the folders it touches at import."""
    root = ref_loader.reference_root()
    if root is None:
        raise SystemExit("reference tree not found (set B2F_REFERENCE)")
    fakes.install()
    for d in ("static", "templates", "clustering_results"):
        os.makedirs(os.path.join(work, d), exist_ok=True)
    os.chdir(work)
    sys.path.insert(0, root)
    try:
        import duplicate  # noqa: the reference's module, verbatim
    finally:
        sys.path.remove(root)
    assert os.path.abspath(duplicate.__file__).startswith(os.path.abspath(root))
    logging.disable(logging.CRITICAL)
    return duplicate, root


def make_config(root: str, work: str, tag: str) -> str:
    cfg = json.load(open(os.path.join(root, "config.json")))
    cfg["system"]["database_path"] = os.path.join(work, f"{tag}.db")
    cfg["system"]["image_cache_dir"] = os.path.join(work, "image_cache")
    cfg["image_processing"]["max_workers"] = 1           # visits in index order
    cfg["vector_database"]["mode"] = "memory"
    path = os.path.join(work, f"{tag}.json")
    json.dump(cfg, open(path, "w"))
    return path


def url(i: int) -> str:
    return f"http://synthetic.invalid/visit_{i:06d}.jpg"


def new_system(duplicate, root, work, tag, rows):
    s = duplicate.SmartFaceRecognition(config_file=make_config(root, work, tag))
    unit = rows / np.linalg.norm(rows, axis=1, keepdims=True)          # duplicate.py:1491-1496
    table = {}
    for i in range(len(rows)):
        e = unit[i].astype(np.float32)
        table[url(i)] = {"embedding": e, "quality": {"overall": 0.9, "blur": 0.9, "pose": 0.9, "lighting": 0.9},
                         "bbox": np.array([0, 0, 100, 100], np.float32), "det_score": 0.99, "face_confidence": 0.99,
                         "face_hash": s.compute_face_hash(e), "image_source": url(i)}
    s.extract_face_embedding = lambda image_source, save_image=False, output_dir=None: table.get(image_source)
    return s, table


def run_online(duplicate, root, work, tag, rows, with_duplicate_check: bool):
    s, table = new_system(duplicate, root, work, tag, rows)
    if with_duplicate_check:
        # a database that has seen an unusable image before owns the `low_similarity_images` table, and only then does
        # is_duplicate_image reach its 0.95 embedding check (on a fresh file the missing table raises and it returns False)
        s.store_low_similarity_image("warmup", "c", "", "http://synthetic.invalid/none.jpg", None, 0.0, None, "no face")
    visits = {"visits": [{"id": f"v{i}", "image": url(i), "customerId": f"c{i}", "entryTime": f"t{i}"}
                         for i in range(len(rows))]}
    vpath = os.path.join(work, f"{tag}_visits.json")
    json.dump(visits, open(vpath, "w"))
    results = s.process_visit_data(vpath, output_folder=None, save_images=False)
    conn = sqlite3.connect(s.database_path)
    person_row = {pid: int(path.rsplit("_", 1)[1].split(".")[0])
                  for pid, path in conn.execute("SELECT id, image_path FROM persons")}
    label = np.full(len(rows), -1, np.int64)
    sim = np.zeros(len(rows), np.float32)
    for pid, vid, sm in conn.execute("SELECT person_id, visit_id, similarity FROM person_visits"):
        label[int(vid[1:])] = person_row[pid]
        sim[int(vid[1:])] = sm
    conn.close()
    return label, sim, results, s


def run_merge(duplicate, root, work, tag, rows, thr):
    s, table = new_system(duplicate, root, work, tag, rows)
    ids = []
    for i in range(len(rows)):                               # every row is a person (ids ascend with the row index)
        pid = s.add_person(f"P{i}", url(i), table[url(i)])
        assert pid == i + 1
        s.store_visit_info(pid, f"v{i}", f"c{i}", "", url(i), None, 1.0)
        ids.append(pid)
    s.find_and_merge_duplicates(thr)
    conn = sqlite3.connect(s.database_path)
    leader = np.full(len(rows), -1, np.int64)
    for pid, vid in conn.execute("SELECT person_id, visit_id FROM person_visits"):
        leader[int(vid[1:])] = pid - 1
    survivors = sorted(pid - 1 for (pid,) in conn.execute("SELECT id FROM persons"))
    conn.close()
    assert survivors == sorted(set(leader.tolist()))
    assert s.vector_db.get_embedding_count() == len(survivors)
    return leader


def run_search(duplicate, root, work, tag, rows, queries, k, thr):
    s, table = new_system(duplicate, root, work, tag, rows)
    for i in range(len(rows)):
        s.vector_db.add_embedding(i, table[url(i)]["embedding"], {"name": f"P{i}"})
    idx = np.full((len(queries), k), -1, np.int64)
    score = np.zeros((len(queries), k), np.float32)
    for qi, q in enumerate(queries):
        res = s.vector_db.search_similar(q, k=k, threshold=thr)
        for j, r in enumerate(res):
            idx[qi, j], score[qi, j] = r["person_id"], r["similarity"]
    return idx, score


def margins(rows, thresholds):
    u = (rows / np.linalg.norm(rows, axis=1, keepdims=True)).astype(np.float64)
    s = u @ u.T
    np.fill_diagonal(s, -2.0)
    return {t: float(np.abs(s - t).min()) for t in thresholds}


def main():
    work = tempfile.mkdtemp(prefix="b2f_cluster_golden_")
    here = os.getcwd()
    duplicate, root = load_reference_app(work)
    cfg = json.load(open(os.path.join(root, "config.json")))["face_recognition"]
    g_thr, s_thr, d_thr, m_thr = (cfg["grouping_threshold_file"], cfg["similarity_threshold"],
                                  cfg["duplicate_similarity_threshold"], cfg["merge_duplicate_threshold"])
    gold = {"thresholds": np.asarray([g_thr, s_thr, d_thr, m_thr], np.float64)}
    for name, rows in inputs.cluster_cases():
        mg = margins(rows, (g_thr, s_thr, d_thr, m_thr))
        assert min(mg.values()) > 1e-5, (name, mg)           # no pair sits on a threshold: fp32 / fp64 agree on every decision
        label, sim, res, _ = run_online(duplicate, root, work, f"on_{name}", rows, False)
        gold[f"online_{name}_label"], gold[f"online_{name}_sim"] = label, sim
        label2, sim2, res2, _ = run_online(duplicate, root, work, f"ond_{name}", rows, True)
        gold[f"online_dup_{name}_label"], gold[f"online_dup_{name}_sim"] = label2, sim2
        leader = run_merge(duplicate, root, work, f"mg_{name}", rows, m_thr)
        gold[f"merge_{name}_leader"] = leader
        qs, _ = inputs.planted_queries(rows, 90, 12, noise=0.8)
        idx, score = run_search(duplicate, root, work, f"se_{name}", rows, qs, 5, s_thr)
        gold[f"search_{name}_idx"], gold[f"search_{name}_score"] = idx, score
        print(f"{name}: n={len(rows)} margins={ {k: round(v, 6) for k, v in mg.items()} } "
              f"online persons={int((label == np.arange(len(rows))).sum())} "
              f"(dup-check: {int((label2 == np.arange(len(rows))).sum())} persons, {int((label2 < 0).sum())} skipped; "
              f"results {res2}) merge survivors={int((leader == np.arange(len(rows))).sum())} "
              f"search hits={int((idx >= 0).sum())}")
    os.chdir(here)
    np.savez_compressed(os.path.join(OUT, "cluster_outputs.npz"), **gold)
    print("wrote", os.path.join(OUT, "cluster_outputs.npz"), sum(v.nbytes for v in gold.values()), "bytes raw")


if __name__ == "__main__":
    main()
