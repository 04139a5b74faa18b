"""TEST INFRASTRUCTURE -- not part of the product path.

Stand-ins for the two wheels `duplicate.py` / `qdrant_manager.py` of the reference import but this image lacks,
so those files run VERBATIM (tests/golden/make_cluster_golden.py):

  * `qdrant_client` (not listed in the reference's requirements.txt at all, so unpinned) -- call sites
    qdrant_manager.py:12-14,42,48,67,126,164,202,222,240,279.  `QdrantClient(":memory:")` is qdrant-client's
    "local mode": a numpy brute-force collection.  Restated here from its published behaviour
    (qdrant_client/local/local_collection.py, distances.py):
      - Cosine collections L2-normalise every vector at upsert (float32) and the query at search;
      - scores = stored @ query (float32 dot); results ordered by descending score (`argsort(...)[::-1]`);
      - `score_threshold` keeps scores >= threshold for bigger-is-better distances; `limit` cuts the list;
      - `retrieve(with_vectors=True)` returns the stored (normalised) vector; `upsert` replaces by id;
      - `delete` takes a `PointIdsList` or a `FilterSelector` (an empty filter matches every point).
  * `insightface.app.FaceAnalysis` (requirements.txt:12, unpinned) -- duplicate.py:20,356-358.  The clustering
    decisions under test never reach the model (the harness supplies the embeddings), so the stand-in only
    satisfies the constructor / prepare() calls of `initialize_model`.

`install()` puts them into sys.modules; it never shadows a real wheel.
"""
from __future__ import annotations

import sys
import types
from types import SimpleNamespace
from typing import Dict, List

import numpy as np


# ---------------------------------------------------------------------------------------------
# qdrant_client.http.models
# ---------------------------------------------------------------------------------------------
class Distance:
    COSINE = "Cosine"
    EUCLID = "Euclid"
    DOT = "Dot"


class VectorParams:
    def __init__(self, size, distance):
        self.size, self.distance = size, distance


class PointStruct:
    def __init__(self, id, vector, payload=None):
        self.id, self.vector, self.payload = id, vector, payload or {}


class PointIdsList:
    def __init__(self, points):
        self.points = list(points)


class Filter:
    def __init__(self, must=None, should=None, must_not=None):
        self.must, self.should, self.must_not = must, should, must_not


class FilterSelector:
    def __init__(self, filter):
        self.filter = filter


class ScoredPoint:
    def __init__(self, id, score, payload, vector=None):
        self.id, self.score, self.payload, self.vector, self.version = id, score, payload, vector, 0


class Record:
    def __init__(self, id, payload, vector=None):
        self.id, self.payload, self.vector = id, payload, vector


# ---------------------------------------------------------------------------------------------
# qdrant_client.QdrantClient, local (":memory:") mode
# ---------------------------------------------------------------------------------------------
class _Collection:
    def __init__(self, size: int, distance: str):
        self.size, self.distance = size, distance
        self.ids: List = []
        self.vectors = np.zeros((0, size), np.float32)
        self.payloads: List[dict] = []

    def _row(self, pid):
        try:
            return self.ids.index(pid)
        except ValueError:
            return -1

    def upsert(self, point: PointStruct) -> None:
        v = np.asarray(point.vector, dtype=np.float32)
        if self.distance == Distance.COSINE:
            n = np.linalg.norm(v)
            v = v / (n if n != 0.0 else np.float32(1.1920929e-07))
        r = self._row(point.id)
        if r >= 0:
            self.vectors[r], self.payloads[r] = v, dict(point.payload)
        else:
            self.ids.append(point.id)
            self.vectors = np.vstack([self.vectors, v[None, :]])
            self.payloads.append(dict(point.payload))

    def search(self, query, limit, score_threshold, with_payload=True, with_vectors=False):
        q = np.asarray(query, dtype=np.float32)
        if len(self.ids) == 0:
            return []
        if self.distance == Distance.COSINE:
            n = np.linalg.norm(q)
            q = q / (n if n != 0.0 else np.float32(1.1920929e-07))
            scores = np.dot(self.vectors, q)
            order = np.argsort(scores)[::-1]
        elif self.distance == Distance.DOT:
            scores = np.dot(self.vectors, q)
            order = np.argsort(scores)[::-1]
        else:
            scores = np.linalg.norm(self.vectors - q[None, :], axis=1)
            order = np.argsort(scores)
        out = []
        for r in order:
            if len(out) >= limit:
                break
            s = float(scores[r])
            if score_threshold is not None:
                if self.distance == Distance.EUCLID:
                    if s > score_threshold:
                        break
                elif s < score_threshold:
                    break
            out.append(ScoredPoint(self.ids[r], s, dict(self.payloads[r]) if with_payload else None,
                                   self.vectors[r].tolist() if with_vectors else None))
        return out

    def delete_ids(self, ids) -> None:
        keep = [i for i, pid in enumerate(self.ids) if pid not in set(ids)]
        self.ids = [self.ids[i] for i in keep]
        self.payloads = [self.payloads[i] for i in keep]
        self.vectors = self.vectors[keep] if keep else np.zeros((0, self.size), np.float32)


class QdrantClient:
    def __init__(self, location=None, host=None, port=None, **kw):
        if location != ":memory:":
            raise ConnectionError("oracle.fakes.QdrantClient only provides the in-memory local mode")
        self._collections: Dict[str, _Collection] = {}

    def get_collections(self):
        return SimpleNamespace(collections=[SimpleNamespace(name=n) for n in self._collections])

    def create_collection(self, collection_name, vectors_config, **kw):
        self._collections[collection_name] = _Collection(vectors_config.size, vectors_config.distance)
        return True

    def upsert(self, collection_name, points, **kw):
        for p in points:
            self._collections[collection_name].upsert(p)

    def search(self, collection_name, query_vector, limit=10, score_threshold=None, with_payload=True,
               with_vectors=False, **kw):
        return self._collections[collection_name].search(query_vector, limit, score_threshold, with_payload,
                                                         with_vectors)

    def retrieve(self, collection_name, ids, with_payload=True, with_vectors=False, **kw):
        c = self._collections[collection_name]
        out = []
        for pid in ids:
            r = c._row(pid)
            if r >= 0:
                out.append(Record(pid, dict(c.payloads[r]) if with_payload else None,
                                  c.vectors[r].tolist() if with_vectors else None))
        return out

    def delete(self, collection_name, points_selector, **kw):
        c = self._collections[collection_name]
        if isinstance(points_selector, PointIdsList):
            c.delete_ids(points_selector.points)
        elif isinstance(points_selector, FilterSelector):
            c.delete_ids(list(c.ids))
        else:
            c.delete_ids(list(points_selector))

    def get_collection(self, collection_name):
        c = self._collections[collection_name]
        return SimpleNamespace(points_count=len(c.ids), vectors_count=len(c.ids), status="green",
                               config=SimpleNamespace(params=SimpleNamespace(
                                   vectors=SimpleNamespace(size=c.size, distance=c.distance))))


# ---------------------------------------------------------------------------------------------
# insightface.app.FaceAnalysis
# ---------------------------------------------------------------------------------------------
class FaceAnalysis:
    def __init__(self, name="buffalo_l", **kw):
        self.name = name

    def prepare(self, ctx_id=0, det_size=(640, 640), **kw):
        self.det_size = det_size

    def get(self, img, max_num=0):
        raise RuntimeError("oracle.fakes.FaceAnalysis has no model: the harness supplies embeddings directly")


def install() -> None:
    """Inject `qdrant_client` and `insightface` stand-ins (idempotent; never shadows real wheels)."""
    try:
        import qdrant_client  # noqa: F401
    except Exception:
        qc = types.ModuleType("qdrant_client")
        http = types.ModuleType("qdrant_client.http")
        models = types.ModuleType("qdrant_client.http.models")
        for cls in (Distance, VectorParams, PointStruct, PointIdsList, Filter, FilterSelector, ScoredPoint, Record):
            setattr(models, cls.__name__, cls)
        http.models = models
        qc.QdrantClient, qc.http, qc.__b2f_fake__ = QdrantClient, http, True
        sys.modules.update({"qdrant_client": qc, "qdrant_client.http": http, "qdrant_client.http.models": models})
    try:
        import insightface  # noqa: F401
    except Exception:
        ins = types.ModuleType("insightface")
        app = types.ModuleType("insightface.app")
        app.FaceAnalysis = FaceAnalysis
        ins.app, ins.__b2f_fake__ = app, True
        sys.modules.update({"insightface": ins, "insightface.app": app})
