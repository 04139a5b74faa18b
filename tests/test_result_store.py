"""CPU: the clustering-result JSON and the SQLite tables against a fixture produced by the reference's own writers
(tests/golden/make_result_store_golden.py: json_storage.py run here, CREATE TABLE statements of duplicate.py run in
SQLite), and the label -> rows / groups glue against the oracle's online clustering."""
import json
import os
import re
import sqlite3
from datetime import datetime, timezone

import numpy as np
import pytest

from oracle import restate
from scrfd_arcface_facerecognition_b200 import result_store as rs

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "result_store_golden.json")))


def test_group_records_match_reference_writer():
    mine = rs.format_groups_for_json(GOLD["groups_in"])
    assert mine == GOLD["groups_json"]
    assert [list(g) for g in mine] == [list(g) for g in GOLD["groups_json"]]           # field order too
    assert list(mine[0]["visits"][0]) == list(GOLD["groups_json"][0]["visits"][0])
    assert rs.format_groups_for_json([]) == []


def test_payload_envelope_matches_reference_writer(tmp_path):
    ref = GOLD["payload"]
    now = datetime(2026, 3, 4, 5, 6, 7, 123456, tzinfo=timezone.utc)
    mine = rs.clustering_payload(GOLD["groups_in"], ref["total_processed"], GOLD["results_in"], job_id=ref["job_id"], now=now)
    assert list(mine) == list(ref)                                   # same keys, same order
    for k in ref:
        if k != "timestamp":
            assert mine[k] == ref[k], k
    assert mine["timestamp"] == "2026-03-04T05:06:07.123456Z"
    assert re.fullmatch(r"\d{4}-\d\d-\d\dT\d\d:\d\d:\d\d\.\d{6}Z", ref["timestamp"])       # the reference's form
    path = rs.save_clustering_results(GOLD["groups_in"], 6, GOLD["results_in"], str(tmp_path), now=now)
    name = os.path.basename(path)
    assert re.sub(r"\d", "D", re.sub(r"_[0-9a-f]{8}\.json$", "_JJJJJJJJ.json", name)) == GOLD["file_name_pattern"]
    text = open(path, encoding="utf-8").read()
    assert text.startswith('{\n  "job_id"') == GOLD["indent_two"]
    assert json.loads(text)["groups"] == GOLD["groups_json"]
    uuid_like = json.loads(text)["job_id"]
    assert re.fullmatch(r"[0-9a-f]{8}-[0-9a-f]{4}-4[0-9a-f]{3}-[89ab][0-9a-f]{3}-[0-9a-f]{12}", uuid_like)


def test_sqlite_schema_matches_reference_statements():
    conn = sqlite3.connect(":memory:")
    rs.PersonDatabase(connection=conn)
    for name in ("persons", "face_quality", "person_visits", "low_similarity_images"):
        assert [list(r) for r in conn.execute(f"PRAGMA table_info({name})")] == GOLD["tables"][name], name
        assert [list(r) for r in conn.execute(f"PRAGMA foreign_key_list({name})")] == GOLD["tables"][name + "::fk"], name
        uniq = sorted([c[2] for c in conn.execute(f"PRAGMA index_info({ix[1]})")][0]
                      for ix in conn.execute(f"PRAGMA index_list({name})") if ix[2])
        assert uniq == GOLD["tables"][name + "::unique"], name


def _visits(n):
    return [{"id": f"v{i}", "customerId": f"c{i}", "image": f"http://example.invalid/{i}.jpg", "entryTime": f"t{i}", "branchId": "b",
             "camera": "cam", "entryEventIds": [{"event": "entry", "fileName": f"f{i}.jpg"}]} for i in range(n)]


def test_labels_to_rows_and_groups(tmp_path):
    rng = np.random.default_rng(11)
    centres = rng.standard_normal((5, 512)).astype(np.float32)
    emb = np.concatenate([centres[i % 5][None] + 0.03 * rng.standard_normal((1, 512)).astype(np.float32) for i in range(17)])
    label = restate.online_clusters(emb, 0.6)
    sim = restate.online_similarities(emb, label)
    assert sorted(set(label.tolist())) == [0, 1, 2, 3, 4]
    conn = sqlite3.connect(":memory:")
    db = rs.PersonDatabase(connection=conn)
    out = rs.write_online_clustering(_visits(17), label, sim, db, str(tmp_path), clock=lambda: 1767261600.9)
    assert out["results"] == {"processed": 17, "recognized": 12, "new_persons": 5, "no_faces": 0, "low_quality": 0,
                              "download_failed": 0, "duplicate_faces": 0, "low_similarity": 0}
    assert out["person_ids"] == {0: 1, 1: 2, 2: 3, 3: 4, 4: 5}                  # sequential ids in visit order
    persons = conn.execute("SELECT id, name, image_path, match_count FROM persons ORDER BY id").fetchall()
    assert persons[0] == (1, "Person_c0_1767261600", "http://example.invalid/0.jpg", 3)
    assert [p[3] for p in persons] == [3, 3, 2, 2, 2]                            # joins of each person
    rows = conn.execute("SELECT person_id, visit_id, customer_id, entry_time, image_url, similarity FROM person_visits ORDER BY id").fetchall()
    assert len(rows) == 17 and rows[0] == (1, "v0", "c0", "t0", "http://example.invalid/0.jpg", 1.0)
    assert [r[0] for r in rows] == [int(label[i]) + 1 for i in range(17)]
    np.testing.assert_allclose([r[5] for r in rows[1:]], sim[1:], rtol=0, atol=1e-7)
    payload = json.load(open(out["json_path"], encoding="utf-8"))
    assert payload["total_groups"] == 17 and payload["total_processed"] == 17 and payload["results"] == out["results"]
    g7 = payload["groups"][7]                                                    # one group per processed visit
    assert g7["person_id"] == int(label[7]) + 1 and g7["group_id"] == "c7" and g7["fileName"] == "f7.jpg" and g7["camera"] == "cam"
    assert g7["visits"] == [{"visit_id": "v7", "customer_id": "c7", "image_url": "http://example.invalid/7.jpg", "entry_time": "t7",
                             "similarity": pytest.approx(float(sim[7]))}]
    web = db.get_person_groups_for_web()
    assert [w["person_id"] for w in web[:2]] == [1, 2] and web[0]["visit_count"] == 4


def test_duplicate_hash_and_bad_labels():
    conn = sqlite3.connect(":memory:")
    db = rs.PersonDatabase(connection=conn)
    label, sim = np.array([0, 1, 1]), np.array([0.0, 0.1, 0.9], np.float32)
    out = rs.write_online_clustering(_visits(3), label, sim, db, None, face_hashes=["h", "h", "x"])
    assert out["results"]["new_persons"] == 1 and out["results"]["duplicate_faces"] == 2 and out["json_path"] is None
    with pytest.raises(ValueError):
        rs.write_online_clustering(_visits(2), np.array([1, 1]), np.zeros(2), db)          # a later row cannot lead
    with pytest.raises(ValueError):
        rs.write_online_clustering(_visits(3), np.array([0, 0, 1]), np.zeros(3), db)       # row 1 is not a founder
    with pytest.raises(ValueError):
        rs.write_online_clustering(_visits(3), np.array([0, 0]), np.zeros(2), db)
    db.store_low_similarity_image("v9", "c9", "t", "u", None, 0.0, None, "No face detected, low confidence, or side face")
    assert conn.execute("SELECT COUNT(*) FROM low_similarity_images").fetchone()[0] == 1
