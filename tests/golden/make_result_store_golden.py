"""Generates tests/golden/result_store_golden.json by running the REFERENCE's own writers in this container:

* `json_storage.JSONStorageManager.format_groups_for_json` / `save_clustering_results` (reference json_storage.py) on
  synthetic person groups (made-up ids and URLs -- nothing from the reference's data files);
* the CREATE TABLE statements of reference duplicate.py (:201-252, :1677-1699), pulled out of the source text and run
  in an in-memory SQLite to record each table's `PRAGMA table_info` (duplicate.py itself cannot be imported here:
  insightface / qdrant_client are absent).

usage: python tests/golden/make_result_store_golden.py [/root/reference]
"""
import glob
import importlib.util
import json
import os
import re
import sqlite3
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))


def synthetic_groups():
    ev = [{"event": "entry", "fileName": "cam3_000017.jpg", "camera": "cam-3", "age": "31", "gender": "F"}]
    v = lambda i, **kw: dict({"visit_id": f"v{i}", "customer_id": f"c{i}", "customerId": f"c{i}", "image_url": f"http://example.invalid/{i}.jpg",
                              "image": f"http://example.invalid/{i}.jpg", "entry_time": f"2026-01-0{i % 9 + 1}T10:00:00Z",
                              "entryTime": f"2026-01-0{i % 9 + 1}T10:00:00Z", "similarity": 0.5 + 0.05 * i, "branchId": "b1",
                              "camera": "", "entryEventIds": []}, **kw)
    return [
        {"person_id": 1, "person_name": "Person_c0_1767261600", "visits": [v(0, similarity=1.0)]},
        {"person_id": 1, "person_name": "Person_c0_1767261600", "visits": [v(1, entryEventIds=ev)]},
        {"person_id": 2, "person_name": "Person_c2_1767261601", "visits": [v(2, camera="lobby", customer={"age": 44, "gender": "male"})]},
        {"person_id": 3, "visits": [v(3, age="27", gender="m"), v(4, age="x", gender="unknown"), v(5, similarity=None)]},
        {"person_id": 4, "person_name": "empty", "visits": []},
        {"person_id": 5, "person_name": "bare", "visits": [{"id": "raw-1", "image": "http://example.invalid/raw.jpg"}]},
    ]


def main():
    ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
    cwd = os.getcwd()
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        os.chdir(tmp)                                    # the module creates ./clustering_results at import
        try:
            spec = importlib.util.spec_from_file_location("ref_json_storage", os.path.join(ref, "json_storage.py"))
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
            mgr = mod.JSONStorageManager(output_dir=os.path.join(tmp, "out"))
            groups = synthetic_groups()
            results = {"processed": 6, "recognized": 3, "new_persons": 3, "no_faces": 0, "low_quality": 0, "download_failed": 0,
                       "duplicate_faces": 0, "low_similarity": 0}
            out["groups_in"] = groups
            out["results_in"] = results
            out["groups_json"] = mgr.format_groups_for_json(groups)
            assert mgr.save_clustering_results(groups, 6, results)
            path, = glob.glob(os.path.join(tmp, "out", "clustering_results_*.json"))
            out["file_name_pattern"] = re.sub(r"\d", "D", re.sub(r"_[0-9a-f]{8}\.json$", "_JJJJJJJJ.json", os.path.basename(path)))
            text = open(path, encoding="utf-8").read()
            payload = json.loads(text)
            out["payload"] = payload
            out["indent_two"] = text.startswith('{\n  "job_id"')
        finally:
            os.chdir(cwd)
    src = open(os.path.join(ref, "duplicate.py"), encoding="utf-8").read()
    stmts = re.findall(r"CREATE TABLE IF NOT EXISTS\s+\w+\s*\(.*?\n\s*\)\s*'''", src, flags=re.S)
    db = sqlite3.connect(":memory:")
    tables = {}
    for st in stmts:
        st = st.rstrip("'").strip()
        name = re.match(r"CREATE TABLE IF NOT EXISTS\s+(\w+)", st).group(1)
        if name in tables:
            continue
        db.execute(st)
        tables[name] = [list(r) for r in db.execute(f"PRAGMA table_info({name})")]
        tables[name + "::fk"] = [list(r) for r in db.execute(f"PRAGMA foreign_key_list({name})")]
        tables[name + "::unique"] = sorted(
            [c[2] for c in db.execute(f"PRAGMA index_info({ix[1]})")][0] for ix in db.execute(f"PRAGMA index_list({name})") if ix[2])
    out["tables"] = tables
    with open(os.path.join(HERE, "result_store_golden.json"), "w") as f:
        json.dump(out, f, indent=1)                       # key order kept: it is part of the format
    print("tables:", [t for t in tables if "::" not in t], "groups:", len(out["groups_json"]))


if __name__ == "__main__":
    main()
