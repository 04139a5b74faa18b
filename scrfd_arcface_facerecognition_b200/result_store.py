"""On-disk results of a clustering job, written from the labels the GPU path produces.

Two formats the reference's web UI consumes, kept field for field:

* the ``clustering_results_<timestamp>_<job>.json`` payload (reference ``json_storage.py:31-141`` group records,
  ``:192-245`` envelope: job_id, status, timestamp, total_processed, total_groups, results, message, groups);
* the SQLite tables ``persons``, ``face_quality``, ``person_visits`` (reference ``duplicate.py:201-252``) and
  ``low_similarity_images`` (``:1677-1699``), with the row writers ``add_person`` (``:1556-1560``), ``store_visit_info``
  (``:1657-1675``), ``update_person_stats`` (``:1645-1655``), ``store_low_similarity_image``.

The reference fills both from a 4-thread pool, so its group order and person ids depend on thread timing
(``duplicate.py:1953-1975``); here the visits are taken in index order, which is the order
``Gallery.online_clusters`` decides them in.  `write_online_clustering` is the glue: visit
records + their embeddings -> GPU labels -> rows and payload.
"""
from __future__ import annotations

import json
import os
import sqlite3
import time
import uuid
from collections import Counter
from datetime import datetime, timezone
from typing import Any, Dict, List, Optional, Sequence

import numpy as np

RESULT_KEYS = ("processed", "recognized", "new_persons", "no_faces", "low_quality", "download_failed", "duplicate_faces",
               "low_similarity")                        # counters of one job, reference duplicate.py:1754-1763

_TABLES = {
    "persons": (
        "id INTEGER PRIMARY KEY AUTOINCREMENT", "name TEXT NOT NULL", "image_path TEXT", "face_quality REAL",
        "face_hash TEXT UNIQUE", "created_at TIMESTAMP DEFAULT CURRENT_TIMESTAMP",
        "last_seen TIMESTAMP DEFAULT CURRENT_TIMESTAMP", "match_count INTEGER DEFAULT 0"),
    "face_quality": (
        "id INTEGER PRIMARY KEY AUTOINCREMENT", "person_id INTEGER", "quality_score REAL", "blur_score REAL",
        "pose_score REAL", "lighting_score REAL", "created_at TIMESTAMP DEFAULT CURRENT_TIMESTAMP",
        "FOREIGN KEY (person_id) REFERENCES persons (id)"),
    "person_visits": (
        "id INTEGER PRIMARY KEY AUTOINCREMENT", "person_id INTEGER", "visit_id TEXT", "customer_id TEXT", "entry_time TEXT",
        "image_url TEXT", "saved_image_path TEXT", "similarity REAL", "processed_at TIMESTAMP DEFAULT CURRENT_TIMESTAMP",
        "FOREIGN KEY (person_id) REFERENCES persons (id)"),
    "low_similarity_images": (
        "id INTEGER PRIMARY KEY AUTOINCREMENT", "visit_id TEXT", "customer_id TEXT", "entry_time TEXT", "image_url TEXT",
        "saved_image_path TEXT", "similarity REAL", "best_match_name TEXT", "reason TEXT",
        "processed_at TIMESTAMP DEFAULT CURRENT_TIMESTAMP"),
}


def schema_sql() -> List[str]:
    return [f"CREATE TABLE IF NOT EXISTS {name} ({', '.join(cols)})" for name, cols in _TABLES.items()]


class PersonDatabase:
    """The reference's person-metadata SQLite file.  One connection per call, like the reference, so instances can be
    shared between threads; `path` may be ':memory:' only through `connection` (tests)."""

    def __init__(self, database_path: str = "face_database.db", connection: Optional[sqlite3.Connection] = None) -> None:
        self.database_path = database_path
        self._shared = connection
        self.setup_database()

    def _open(self) -> sqlite3.Connection:
        return self._shared if self._shared is not None else sqlite3.connect(self.database_path)

    def _close(self, conn: sqlite3.Connection) -> None:
        if conn is not self._shared:
            conn.close()

    def setup_database(self) -> None:
        conn = self._open()
        for stmt in schema_sql():
            conn.execute(stmt)
        conn.commit()
        self._close(conn)

    def add_person(self, name: str, image_path: Optional[str], face_quality: Optional[float], face_hash: Optional[str]) -> int:
        """New `persons` row -> its id, or -1 when `face_hash` is already present (the column is UNIQUE; the reference
        reports the failed insert the same way)."""
        conn = self._open()
        try:
            cur = conn.execute("INSERT INTO persons (name, image_path, face_quality, face_hash) VALUES (?, ?, ?, ?)",
                               (name, image_path, face_quality, face_hash))
            conn.commit()
            return int(cur.lastrowid)
        except sqlite3.IntegrityError:
            conn.rollback()
            return -1
        finally:
            self._close(conn)

    def update_person_stats(self, person_id: int) -> None:
        conn = self._open()
        conn.execute("UPDATE persons SET last_seen = CURRENT_TIMESTAMP, match_count = match_count + 1 WHERE id = ?", (person_id,))
        conn.commit()
        self._close(conn)

    def store_visit_info(self, person_id: int, visit_id: str, customer_id: str, entry_time: str, image_url: str,
                         saved_image_path: Optional[str], similarity: float) -> None:
        conn = self._open()
        conn.execute("INSERT OR REPLACE INTO person_visits (person_id, visit_id, customer_id, entry_time, image_url, "
                     "saved_image_path, similarity) VALUES (?, ?, ?, ?, ?, ?, ?)",
                     (person_id, visit_id, customer_id, entry_time, image_url, saved_image_path, float(similarity)))
        conn.commit()
        self._close(conn)

    def store_low_similarity_image(self, visit_id: str, customer_id: str, entry_time: str, image_url: str,
                                   saved_image_path: Optional[str], similarity: float, best_match_name: Optional[str] = None,
                                   reason: Optional[str] = None) -> None:
        conn = self._open()
        conn.execute("INSERT INTO low_similarity_images (visit_id, customer_id, entry_time, image_url, saved_image_path, "
                     "similarity, best_match_name, reason) VALUES (?, ?, ?, ?, ?, ?, ?, ?)",
                     (visit_id, customer_id, entry_time, image_url, saved_image_path, float(similarity), best_match_name, reason))
        conn.commit()
        self._close(conn)

    def get_person_groups_for_web(self) -> List[Dict[str, Any]]:
        """Persons with their visits, most-matched first (the read side of reference duplicate.py:2349-2492, reduced
        to the fields both of its branches return)."""
        conn = self._open()
        persons = conn.execute("SELECT id, name, image_path, face_quality, match_count, last_seen FROM persons "
                               "ORDER BY match_count DESC, last_seen DESC, id ASC").fetchall()
        out = []
        for pid, name, image_path, quality, match_count, last_seen in persons:
            visits = conn.execute("SELECT visit_id, customer_id, entry_time, image_url, saved_image_path, similarity FROM "
                                  "person_visits WHERE person_id = ? ORDER BY id ASC", (pid,)).fetchall()
            out.append({"person_id": pid, "name": name, "image_path": image_path, "face_quality": quality,
                        "match_count": match_count, "last_seen": last_seen, "visit_count": len(visits), "avg_quality": quality,
                        "images": [{"visit_id": v, "customer_id": c, "entry_time": t, "image_url": u, "image_path": p or u,
                                    "similarity": s} for v, c, t, u, p, s in visits]})
        self._close(conn)
        return out


# ---- the clustering_results JSON -------------------------------------------------------------------------------------
def _average_age(visits: Sequence[Dict[str, Any]]) -> Optional[int]:
    ages = []
    for v in visits:
        for rec in [v] + list(v.get("entryEventIds", [])):
            if "age" in rec:
                try:
                    ages.append(int(rec["age"]))
                except (ValueError, TypeError):
                    pass
    return round(sum(ages) / len(ages)) if ages else None


def _common_gender(visits: Sequence[Dict[str, Any]]) -> Optional[str]:
    seen = []
    for v in visits:
        for rec in [v] + list(v.get("entryEventIds", [])):
            g = rec.get("gender") if "gender" in rec else None
            if g and g.lower() in ("male", "female", "m", "f"):
                seen.append(g.lower())
    return Counter(seen).most_common(1)[0][0] if seen else None


def format_groups_for_json(person_groups: Sequence[Dict[str, Any]]) -> List[Dict[str, Any]]:
    """person groups ({person_id, person_name, visits:[...]}) -> the group records of the JSON file.  A group's header
    fields come from its first visit (camera falls back to the first entry event), `group_score` is the mean visit
    similarity rounded to 3 places, groups without visits are dropped (reference json_storage.py:31-141)."""
    out = []
    for g in person_groups:
        visits = g.get("visits", [])
        if not visits:
            continue
        pid = g.get("person_id")
        sims = [v.get("similarity", 0.0) for v in visits if v.get("similarity") is not None]
        first = visits[0]
        events = first.get("entryEventIds", [])
        ev0 = events[0] if events else {}
        customer = first.get("customer", {})
        age, gender = customer.get("age"), customer.get("gender")
        out.append({
            "group_id": first.get("customerId", first.get("customer_id", "")),
            "person_id": pid,
            "person_name": g.get("person_name", f"Person_{pid}"),
            "timestamp": first.get("entryTime", first.get("entry_time", "")),
            "group_score": round(sum(sims) / len(sims) if sims else 0.0, 3),
            "camera": first.get("camera", "") or ev0.get("camera", ""),
            "event": ev0.get("event", ""),
            "branchId": first.get("branchId", ""),
            "fileName": ev0.get("fileName", ""),
            "age": age if age is not None else _average_age(visits),
            "gender": gender if gender is not None else _common_gender(visits),
            "visit_count": len(visits),
            "visits": [{"visit_id": v.get("visit_id", v.get("id")),
                        "customer_id": v.get("customerId", v.get("customer_id")),
                        "image_url": v.get("image_url", v.get("image")),
                        "entry_time": v.get("entryTime", v.get("entry_time")),
                        "similarity": v.get("similarity", 0.0)} for v in visits],
        })
    return out


def clustering_payload(groups: Sequence[Dict[str, Any]], total_processed: int, results: Dict[str, Any],
                       job_id: Optional[str] = None, now: Optional[datetime] = None) -> Dict[str, Any]:
    json_groups = format_groups_for_json(groups)
    now = now or datetime.now(timezone.utc)
    return {"job_id": job_id or str(uuid.uuid4()), "status": "finished",
            "timestamp": now.astimezone(timezone.utc).isoformat().replace("+00:00", "Z"),
            "total_processed": total_processed, "total_groups": len(json_groups), "results": results,
            "message": f"Processing completed. Created {len(json_groups)} groups from {total_processed} images",
            "groups": json_groups}


def save_clustering_results(groups: Sequence[Dict[str, Any]], total_processed: int, results: Dict[str, Any],
                            output_dir: str = "clustering_results", job_id: Optional[str] = None,
                            now: Optional[datetime] = None) -> str:
    """Writes clustering_results_<YYYYmmdd_HHMMSS>_<job[:8]>.json (indent 2, UTF-8 kept) and returns its path."""
    os.makedirs(output_dir, exist_ok=True)
    payload = clustering_payload(groups, total_processed, results, job_id, now)
    stamp = (now or datetime.now()).strftime("%Y%m%d_%H%M%S")
    path = os.path.join(output_dir, f"clustering_results_{stamp}_{payload['job_id'][:8]}.json")
    with open(path, "w", encoding="utf-8") as f:
        json.dump(payload, f, indent=2, ensure_ascii=False)
    return path


# ---- from GPU labels to both formats ---------------------------------------------------------------------------------
def visit_group(person_id: int, person_name: str, visit: Dict[str, Any], index: int, similarity: float) -> Dict[str, Any]:
    """The one-visit person group the reference emits per processed visit (duplicate.py:1828-1846, :1876-1894)."""
    visit_id = visit.get("id", f"visit_{index}")
    customer_id = visit.get("customerId", f"customer_{index}")
    image_url, entry_time = visit.get("image"), visit.get("entryTime", "")
    return {"person_id": person_id, "person_name": person_name,
            "visits": [{"visit_id": visit_id, "customer_id": customer_id, "customerId": customer_id, "image_url": image_url,
                        "image": image_url, "entry_time": entry_time, "entryTime": entry_time, "similarity": float(similarity),
                        "branchId": visit.get("branchId", ""), "camera": visit.get("camera", ""),
                        "entryEventIds": visit.get("entryEventIds", [])}]}


def write_online_clustering(visits: Sequence[Dict[str, Any]], labels: np.ndarray, similarity: np.ndarray,
                            db: PersonDatabase, output_dir: Optional[str] = None, face_hashes: Optional[Sequence[str]] = None,
                            qualities: Optional[Sequence[float]] = None, clock=time.time, job_id: Optional[str] = None,
                            now: Optional[datetime] = None) -> Dict[str, Any]:
    """Persist one clustering job.  `labels[i]` is the visit index of the person visit i belongs to
    (`GalleryManager.online_person_labels` / `Gallery.online_clusters`: labels[i] == i founds a person) and
    `similarity[i]` the cosine the decision was taken on (`online_similarities`).  Visits are written in index order:
    founders become `persons` rows named Person_<customer>_<unix time> (duplicate.py:1821, :1916), every visit becomes a
    `person_visits` row, joins bump the person's match_count.  Returns {results, groups, person_ids, json_path}."""
    n = len(visits)
    labels = np.asarray(labels).reshape(-1)
    similarity = np.asarray(similarity, np.float32).reshape(-1)
    if len(labels) != n or len(similarity) != n:
        raise ValueError(f"{n} visits but {len(labels)} labels / {len(similarity)} similarities")
    results = {k: 0 for k in RESULT_KEYS}
    groups: List[Dict[str, Any]] = []
    person_of: Dict[int, tuple] = {}
    for i, visit in enumerate(visits):
        lead = int(labels[i])
        if lead > i or lead < 0 or int(labels[lead]) != lead:
            raise ValueError(f"visit {i} is labelled with person {lead}, which is not an earlier founder")
        visit_id = visit.get("id", f"visit_{i}")
        customer_id = visit.get("customerId", f"customer_{i}")
        entry_time, image_url = visit.get("entryTime", ""), visit.get("image")
        results["processed"] += 1
        sim = 1.0 if (lead == i and not person_of) else float(similarity[i])       # the very first person is stored at 1.0
        if lead == i:
            name = f"Person_{customer_id}_{int(clock())}"
            pid = db.add_person(name, image_url, None if qualities is None else float(qualities[i]),
                                None if face_hashes is None else face_hashes[i])
            if pid <= 0:                                   # same face hash already stored: the reference skips the visit
                results["duplicate_faces"] += 1
                continue
            person_of[i] = (pid, name)
            results["new_persons"] += 1
        else:
            if lead not in person_of:                      # its founder was dropped as a stored duplicate
                results["duplicate_faces"] += 1
                continue
            pid, name = person_of[lead]
            db.update_person_stats(pid)
            results["recognized"] += 1
        db.store_visit_info(pid, visit_id, customer_id, entry_time, image_url, visit.get("saved_image_path"), sim)
        groups.append(visit_group(pid, name, visit, i, sim))
    path = None
    if output_dir is not None and groups:
        path = save_clustering_results(groups, results["processed"], results, output_dir, job_id, now)
    return {"results": results, "groups": groups, "person_ids": {i: p[0] for i, p in person_of.items()}, "json_path": path}
