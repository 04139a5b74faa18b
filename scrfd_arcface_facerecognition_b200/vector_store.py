"""GPU-resident stand-in for the reference's `QdrantManager` (qdrant_manager.py:17-300).

Same constructor argument (the application config dict with a `vector_database` section) and the same
method surface and return conventions: `add_embedding / update_embedding` (upsert by person_id),
`search_similar` (cosine top-k, score >= threshold, descending), `delete_embedding`, `get_embedding`,
`get_embedding_count`, `clear_all`, `get_collection_info`.  Qdrant's Cosine collections L2-normalise vectors at
upsert and score with a dot product; `Gallery` does the same, on the device, with the tcgen05 top-k kernel for
the coarse pass and an exact fp32 re-score of the candidates.  Failures are logged and reported through the
boolean / empty-list returns, as the reference does -- except a missing CUDA device, which raises.
"""
from __future__ import annotations

import logging
from typing import Any, Dict, List, Optional

import numpy as np
import torch

from . import _lib
from .engine import stream_ptr
from .gallery import KMAX, Gallery

__all__ = ["QdrantManager", "GalleryManager"]


class GalleryManager:
    def __init__(self, config: Dict[str, Any]):
        self.config = config.get("vector_database", {})
        self.collection_name = self.config.get("collection_name", "face_embeddings")
        self.vector_size = self.config.get("vector_size", 512)
        self.distance_metric = self.config.get("distance_metric", "Cosine")
        self.logger = logging.getLogger(__name__)
        if self.distance_metric != "Cosine":
            raise ValueError(f"only the Cosine metric of the reference config is implemented, got {self.distance_metric}")
        self.gallery = Gallery(dim=self.vector_size)
        self._row_of: Dict[Any, int] = {}

    # ---- upsert / delete (reference qdrant_manager.py:91-136, 190-212, 252-266) --------------------------
    def add_embedding(self, person_id, embedding, metadata: Dict[str, Any]) -> bool:
        try:
            vec = np.asarray(embedding, dtype=np.float32).reshape(-1)
            if vec.shape[0] != self.vector_size:
                self.logger.error(f"Vector size mismatch: expected {self.vector_size}, got {vec.shape[0]}")
                return False
            payload = {"person_id": person_id, **(metadata or {})}
            row = self._row_of.get(person_id)
            if row is None:
                self._row_of[person_id] = len(self.gallery)
                self.gallery.add(vec[None], ids=[person_id], payloads=[payload])
            else:
                self.gallery.replace_rows(torch.tensor([row], device=self.gallery.device),
                                          torch.from_numpy(vec[None]))
                self.gallery.payloads[row] = payload
            return True
        except Exception as e:  # mirrors the reference's log-and-return-False convention
            self.logger.error(f"Failed to add embedding for person {person_id}: {e}")
            return False

    def update_embedding(self, person_id, embedding, metadata: Dict[str, Any]) -> bool:
        return self.add_embedding(person_id, embedding, metadata)

    def delete_embedding(self, person_id) -> bool:
        try:
            row = self._row_of.pop(person_id, None)
            if row is not None:                         # deleting an unknown id succeeds in Qdrant too
                self.gallery.remove(row)
                for pid, r in self._row_of.items():
                    if r > row:
                        self._row_of[pid] = r - 1
            return True
        except Exception as e:
            self.logger.error(f"Failed to delete embedding for person {person_id}: {e}")
            return False

    def clear_all(self) -> bool:
        try:
            self.gallery.clear()
            self._row_of.clear()
            return True
        except Exception as e:
            self.logger.error(f"Failed to clear all embeddings: {e}")
            return False

    # ---- queries (reference qdrant_manager.py:138-188, 214-250, 268-300) ---------------------------------
    def search_similar(self, query_embedding, k: int = 5, threshold: float = 0.0) -> List[Dict[str, Any]]:
        try:
            q = np.asarray(query_embedding, dtype=np.float32).reshape(-1)
            if q.shape[0] != self.vector_size:
                self.logger.error(f"Query vector size mismatch: expected {self.vector_size}, got {q.shape[0]}")
                return []
            if k <= KMAX:
                return self.gallery.search_similar(q, k=k, threshold=threshold)
            return self._search_exhaustive(q, k, threshold)
        except Exception as e:
            self.logger.error(f"Failed to search similar faces: {e}")
            return []

    def _search_exhaustive(self, q: np.ndarray, k: int, threshold: float) -> List[Dict[str, Any]]:
        """k beyond the kernel's running top-8 (duplicate.py asks for every row): exact fp32 scores of the whole
        gallery, ordered (score desc, insertion order asc)."""
        g = self.gallery
        if len(g) == 0:
            return []
        qn = torch.from_numpy(q).to(g.device)
        qn = (qn / qn.norm().clamp_min(1e-30)).contiguous()
        scores = torch.empty(len(g), dtype=torch.float32, device=g.device)
        _lib.check(g.lib.b2f_rows_dot(g.f32.data_ptr(), len(g), g.dim, qn.data_ptr(), scores.data_ptr(), stream_ptr()),
                   "b2f_rows_dot")
        order = torch.argsort(scores, descending=True, stable=True)[:k]
        out = []
        for row, sc in zip(order.tolist(), scores[order].tolist()):
            if sc < threshold:
                break
            payload = g.payloads[row]
            out.append({"person_id": payload.get("person_id", g.ids[row]), "name": payload.get("name", "Unknown"),
                        "similarity": float(sc), "quality": payload.get("quality", 0.0), "metadata": payload})
        return out

    def get_embedding_count(self) -> int:
        return len(self.gallery)

    def get_embedding(self, person_id) -> Optional[np.ndarray]:
        row = self._row_of.get(person_id)
        if row is None:
            return None
        return np.array(self.gallery.f32[row].cpu().tolist())      # float64 list round-trip, like the reference

    def get_collection_info(self) -> Dict[str, Any]:
        return {"name": self.collection_name, "vector_size": self.vector_size, "distance_metric": self.distance_metric,
                "points_count": len(self.gallery), "status": "green"}

    # ---- duplicate.py:2726-2797 on the whole collection ---------------------------------------------------
    def find_duplicate_leaders(self, threshold: float = 0.8) -> Dict[Any, Any]:
        """{person_id: leader person_id} for every merged entry (greedy one-hop leader merge in insertion order)."""
        leader = self.gallery.merge_duplicates(threshold)
        ids = self.gallery.ids
        return {ids[i]: ids[int(l)] for i, l in enumerate(leader) if int(l) != i}


    def online_person_labels(self, grouping_threshold: float) -> Dict[Any, Any]:
        """{person_id: person it joins} under the per-visit decision of duplicate.py:1853-1949 taken in insertion order
        (join the most similar earlier person at >= grouping_threshold, else found a new person)."""
        label = self.gallery.online_clusters(grouping_threshold)
        ids = self.gallery.ids
        return {ids[i]: ids[int(l)] for i, l in enumerate(label)}

    def write_clustering_results(self, visits, grouping_threshold: float, db, output_dir: Optional[str] = None, **kw):
        """Cluster the stored embeddings (one per visit, in insertion order) on the GPU and persist the job in the
        reference's two on-disk formats (`result_store.write_online_clustering`: SQLite person / visit rows and the
        clustering_results JSON of json_storage.py:192-245)."""
        from .result_store import write_online_clustering
        if len(visits) != len(self.gallery):
            raise ValueError(f"{len(visits)} visits for {len(self.gallery)} stored embeddings")
        label = self.gallery.online_clusters(grouping_threshold)
        sim = self.gallery.online_similarities(label, float(self.config.get("similarity_threshold", 0.0)))
        return write_online_clustering(visits, label, sim, db, output_dir, **kw)


QdrantManager = GalleryManager
