"""GPU-resident gallery: cosine top-k matching and duplicate-merge clustering.

Mirrors the matching semantics of the reference:
  * `best_match`      -- strict-greater scan with threshold, lowest index on ties, "Unknown" = -1
                         (reference main.py:136-142 over `compute_similarity`, utils/helpers.py:110-123)
  * `search_similar`  -- exact cosine top-k with score >= threshold, sorted descending
                         (reference qdrant_manager.py:138-188; Cosine collection, config.json:99-100)
  * `merge_duplicates`-- greedy one-hop leader merge in ascending id order
                         (reference duplicate.py:2726-2797)
The gallery rows are L2-normalised once at insert (qdrant does the same for Cosine); queries are
normalised on the fly.  The coarse pass is a tcgen05 GEMM with a running top-k epilogue (Q x G is never
materialised); the k' best coarse candidates are re-scored exactly in fp32 before ranking.
Sharding: rows are split G/P per rank; each rank matches all queries against its shard and the
per-shard top-k (score, global index) are exchanged with one NCCL all_gather and merged locally.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import numpy as np
import torch

from . import _lib
from .engine import default_dtype, stream_ptr, torch_dtype

KMAX = 8   # kTopKMax / kRescore in the kernels


def merge_shard_topk(scores: torch.Tensor, idx: torch.Tensor, k: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """Merge per-shard top-k lists [P,Q,k] into a global [Q,k] by (score desc, index asc); -1 = empty slot.
    Pure index plumbing over a few KB per query batch (runs on whatever device holds the lists)."""
    p, q, kk = scores.shape
    s = scores.permute(1, 0, 2).reshape(q, p * kk)
    i = idx.permute(1, 0, 2).reshape(q, p * kk)
    s = torch.where(i < 0, torch.full_like(s, float("-inf")), s)
    big = torch.iinfo(torch.int64).max
    order = torch.argsort(torch.where(i < 0, torch.full_like(i, big), i), dim=1, stable=True)
    s, i = torch.gather(s, 1, order), torch.gather(i, 1, order)
    order = torch.argsort(s, dim=1, descending=True, stable=True)
    s, i = torch.gather(s, 1, order)[:, :k], torch.gather(i, 1, order)[:, :k]
    return torch.where(i < 0, torch.zeros_like(s), s), i


def merge_shard_top1(scores: torch.Tensor, idx: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """k = 1 special case of merge_shard_topk ([P,Q] -> [Q]): a max and a masked min instead of two sorts.
    Winner = highest score, lowest global index among equals; empty slots (-1) never win."""
    s = torch.where(idx < 0, torch.full_like(scores, float("-inf")), scores)
    best = s.max(dim=0).values
    big = torch.iinfo(torch.int64).max
    win = torch.where((s == best[None, :]) & (idx >= 0), idx, torch.full_like(idx, big)).min(dim=0).values
    win = torch.where(win == big, torch.full_like(win, -1), win)
    return torch.where(win >= 0, best, torch.zeros_like(best)), win


def exchange_shard_topk(s: torch.Tensor, i: torch.Tensor, k: int, world_size: int, group=None):
    """The one collective of sharded matching: every rank contributes its shard's top-k (score [Q,k] f32, global
    index [Q,k] i64, -1 = empty) packed as float64 pairs (indices < 2^53 are exact), then merges locally by
    (score desc, index asc).  Identical on every rank and to the unsharded answer."""
    import torch.distributed as dist
    mine = torch.stack([s.double(), i.double()], dim=-1).contiguous()
    flat = torch.empty((world_size * mine.shape[0],) + tuple(mine.shape[1:]), dtype=mine.dtype, device=mine.device)
    dist.all_gather_into_tensor(flat, mine, group=group)          # concatenated along dim 0 (gloo and NCCL agree on this)
    both = flat.view((world_size,) + tuple(mine.shape))
    gs, gi = both[..., 0].float(), both[..., 1].long()
    if k == 1:
        ms, mi = merge_shard_top1(gs[..., 0], gi[..., 0])
        return ms[:, None], mi[:, None]
    return merge_shard_topk(gs, gi, k)


def pack_top1_keys(s: torch.Tensor, i: torch.Tensor) -> torch.Tensor:
    """(score f32, global index i64 or -1) -> int64 keys whose signed maximum is the winner by (score desc, index asc):
    high word = the score's IEEE bits made monotone, low word = 0xFFFFFFFF - index, top bit flipped; empty = INT64_MIN.
    CUDA tensors go through `b2f_topk_pack_keys`; host tensors (gloo tests of the exchange logic) use the same
    arithmetic in torch."""
    s, i = s.reshape(-1).contiguous(), i.reshape(-1).contiguous()
    if s.is_cuda:
        keys = torch.empty(s.shape[0], dtype=torch.int64, device=s.device)
        _lib.check(_lib.lib().b2f_topk_pack_keys(s.data_ptr(), i.data_ptr(), s.shape[0], keys.data_ptr(), stream_ptr()),
                   "b2f_topk_pack_keys")
        return keys
    u = s.view(torch.int32).to(torch.int64) & 0xFFFFFFFF
    ob = torch.where(u >> 31 != 0, u ^ 0xFFFFFFFF, u ^ 0x80000000)
    key = ((ob - 0x80000000) << 32) | (0xFFFFFFFF - i.clamp(min=0))          # subtracting 2^31 from the high word == flipping bit 63
    return torch.where(i < 0, torch.full_like(key, torch.iinfo(torch.int64).min), key)


def unpack_top1_keys(keys: torch.Tensor):
    if keys.is_cuda:
        s = torch.empty(keys.shape[0], dtype=torch.float32, device=keys.device)
        i = torch.empty(keys.shape[0], dtype=torch.int64, device=keys.device)
        _lib.check(_lib.lib().b2f_topk_unpack_keys(keys.data_ptr(), keys.shape[0], s.data_ptr(), i.data_ptr(), stream_ptr()),
                   "b2f_topk_unpack_keys")
        return s, i
    none = keys == torch.iinfo(torch.int64).min
    ob = ((keys >> 32) + 0x80000000) & 0xFFFFFFFF
    u = torch.where(ob >> 31 != 0, ob ^ 0x80000000, ob ^ 0xFFFFFFFF)
    u = torch.where(u >= 0x80000000, u - (1 << 32), u).to(torch.int32)
    s = torch.where(none, torch.zeros_like(u, dtype=torch.float32), u.view(torch.float32))
    i = torch.where(none, torch.full_like(keys, -1), 0xFFFFFFFF - (keys & 0xFFFFFFFF))
    return s, i


def exchange_shard_top1(s: torch.Tensor, i: torch.Tensor, group=None):
    """k = 1 exchange as ONE collective: every rank packs its shard's (score [Q] f32, global index [Q] i64, -1 = none)
    into int64 keys (`pack_top1_keys`), the ranks MAX all-reduce the keys, and everyone unpacks the same answer --
    identical to `merge_shard_top1` of the gathered lists.  No host synchronisation: legal inside a CUDA graph capture."""
    import torch.distributed as dist
    keys = pack_top1_keys(s, i)
    dist.all_reduce(keys, op=dist.ReduceOp.MAX, group=group)
    return unpack_top1_keys(keys)


class Gallery:
    def __init__(self, dim: int = 512, device: Optional[torch.device] = None, dtype: Optional[int] = None,
                 rank: int = 0, world_size: int = 1, process_group=None):
        if not torch.cuda.is_available():
            raise _lib.B2FError("Gallery needs a CUDA device: there is no CPU fallback")
        self.dim = dim
        self.device = device or torch.device("cuda", torch.cuda.current_device())
        self.dtype = default_dtype() if dtype is None else dtype
        self.rank, self.world_size, self.group = rank, world_size, process_group
        self.lib = _lib.lib()
        self.f32 = torch.empty((0, dim), dtype=torch.float32, device=self.device)      # unit rows (this shard)
        self.h16 = torch.empty((0, dim), dtype=torch_dtype(self.dtype), device=self.device)
        self.ids: List = []
        self.payloads: List[dict] = []
        self.idx_base = 0            # global index of this shard's first row
        self._scratch = {}
        # bumped whenever the row count, the base index or the storage pointers change: a CUDA graph that captured a
        # match against this gallery is valid for one version only (FacePipeline.capture keys its cache on it)
        self.version = 0

    # ---- population ------------------------------------------------------------------------------
    def __len__(self) -> int:
        return int(self.f32.shape[0])

    def _normalise(self, x: torch.Tensor):
        x = x.to(self.device, torch.float32).contiguous().reshape(-1, self.dim)
        n = x.shape[0]
        f32 = torch.empty_like(x)
        h16 = torch.empty((n, self.dim), dtype=torch_dtype(self.dtype), device=self.device)
        _lib.check(self.lib.b2f_l2norm_rows(x.data_ptr(), n, self.dim, f32.data_ptr(), h16.data_ptr(), self.dtype,
                                            None, stream_ptr()), "b2f_l2norm_rows")
        return f32, h16

    def add(self, embeddings, ids: Optional[List] = None, payloads: Optional[List[dict]] = None) -> None:
        if isinstance(embeddings, np.ndarray):
            embeddings = torch.from_numpy(np.ascontiguousarray(embeddings, dtype=np.float32))
        f32, h16 = self._normalise(embeddings)
        start = len(self)
        n = f32.shape[0]
        # rows live in capacity-doubling stores, so one-by-one enrolment (reference main.py:78-105,
        # qdrant_manager.py:91-136) does not re-copy the gallery on every insert; f32 / h16 are views of them
        cap = getattr(self, "_cap", 0)
        if start + n > cap or getattr(self, "_store32", None) is None or self._store32.data_ptr() != self.f32.data_ptr():
            cap = max(start + n, 2 * cap, 1024)
            s32 = torch.empty((cap, self.dim), dtype=torch.float32, device=self.device)
            s16 = torch.empty((cap, self.dim), dtype=torch_dtype(self.dtype), device=self.device)
            s32[:start], s16[:start] = self.f32, self.h16
            self._store32, self._store16, self._cap = s32, s16, cap
        self._store32[start:start + n], self._store16[start:start + n] = f32, h16
        self.f32, self.h16 = self._store32[:start + n], self._store16[:start + n]
        self.ids.extend(ids if ids is not None else range(start, start + n))
        self.payloads.extend(payloads if payloads is not None else [{} for _ in range(n)])
        self.version += 1

    def set_shard(self, embeddings, idx_base: int) -> None:
        """Replace the contents with one shard of a larger gallery whose first row has global index idx_base."""
        self.f32 = torch.empty((0, self.dim), dtype=torch.float32, device=self.device)
        self.h16 = torch.empty((0, self.dim), dtype=torch_dtype(self.dtype), device=self.device)
        self.ids, self.payloads = [], []
        self._store32 = self._store16 = None
        self._cap = 0
        self.add(embeddings)
        self.idx_base = int(idx_base)
        self.version += 1

    def replace_rows(self, rows: torch.Tensor, embeddings: torch.Tensor) -> None:
        """Overwrite local rows (int64 indices into this shard) with new embeddings (upsert of existing ids)."""
        f32, h16 = self._normalise(embeddings)
        self.f32[rows] = f32
        self.h16[rows] = h16

    def remove(self, row: int) -> None:
        """Delete one row; later rows keep their order (indices above `row` shift down by one).  The tail moves down in
        bounded chunks (front to back, so a chunk never overwrites rows it has yet to read): no O(G) temporary."""
        n = len(self)
        chunk = 1 << 16
        for lo in range(row, n - 1, chunk):
            hi = min(lo + chunk, n - 1)
            self.f32[lo:hi] = self.f32[lo + 1:hi + 1].clone()
            self.h16[lo:hi] = self.h16[lo + 1:hi + 1].clone()
        self.f32, self.h16 = self.f32[:n - 1], self.h16[:n - 1]
        del self.ids[row], self.payloads[row]
        self.version += 1

    def clear(self) -> None:
        self.set_shard(torch.empty((0, self.dim)), 0)

    # ---- matching ----------------------------------------------------------------------------------
    def _scratch_for(self, q: int, splits: int):
        key = (q, splits)
        if key not in self._scratch:
            self._scratch[key] = dict(
                ps=torch.empty((q, 2 * splits, KMAX), dtype=torch.float32, device=self.device),
                pi=torch.empty((q, 2 * splits, KMAX), dtype=torch.int32, device=self.device),
                os=torch.empty((q, KMAX), dtype=torch.float32, device=self.device),
                oi=torch.empty((q, KMAX), dtype=torch.int64, device=self.device))
        return self._scratch[key]

    def match_local(self, queries: torch.Tensor, k: int = 1, threshold: float = float("-inf"),
                    strict: bool = False, splits: Optional[int] = None, causal_base: Optional[int] = None):
        """Top-k of this shard for raw (un-normalised) fp32 queries [Q,dim] on the device.
        Returns (scores [Q,k] f32, global indices [Q,k] int64 with -1 for empty slots).
        `causal_base`: query i is matched only against the rows before local row causal_base + i (a prefix search: the
        "best earlier person" of the online loop, reference duplicate.py:1853-1855, for all rows in one pass)."""
        assert 1 <= k <= KMAX
        q = int(queries.shape[0])
        g = len(self)
        if q == 0 or g == 0:
            return (torch.zeros((q, k), dtype=torch.float32, device=self.device),
                    torch.full((q, k), -1, dtype=torch.int64, device=self.device))
        qf, qh = self._normalise(queries)
        # gallery ranges per query tile: the library's plan for this (q, g), or the caller's request rounded to one that
        # tiles the gallery evenly
        if causal_base is not None:
            splits = 1                                   # (the causal limit already shortens every item's gallery range)
        else:
            splits = int(self.lib.b2f_match_plan(q, g)) if splits is None else int(self.lib.b2f_match_splits(g, splits))
        sc = self._scratch_for(q, splits)
        # coarse lists are KMAX wide; only k + 2 candidates per list are tracked (two spare ones for the exact re-score to
        # re-order what 16-bit operands may have swapped): the running list length is what the GEMM's epilogue costs
        keep = min(k + 2, KMAX)
        if causal_base is None:
            _lib.check(self.lib.b2f_match_partial_keep(qh.data_ptr(), q, self.h16.data_ptr(), g, self.dim, self.dtype, None,
                                                       None, KMAX, keep, splits, sc["ps"].data_ptr(), sc["pi"].data_ptr(),
                                                       stream_ptr()), "b2f_match_partial_keep")
        else:
            _lib.check(self.lib.b2f_match_partial_causal(qh.data_ptr(), q, self.h16.data_ptr(), g, self.dim, self.dtype,
                                                         KMAX, keep, splits, int(causal_base), sc["ps"].data_ptr(),
                                                         sc["pi"].data_ptr(), stream_ptr()), "b2f_match_partial_causal")
        thr = float(threshold) if np.isfinite(threshold) else -3.0e38
        _lib.check(self.lib.b2f_match_merge(sc["ps"].data_ptr(), sc["pi"].data_ptr(), q, 2 * splits * KMAX,
                                            qf.data_ptr(), self.f32.data_ptr(), self.dim, KMAX, thr,
                                            1 if strict else 0, self.idx_base, sc["os"].data_ptr(),
                                            sc["oi"].data_ptr(), stream_ptr()), "b2f_match_merge")
        return sc["os"][:, :k], sc["oi"][:, :k]

    def match(self, queries: torch.Tensor, k: int = 1, threshold: float = float("-inf"), strict: bool = False):
        """Global top-k across all shards: local match, one all_gather of (score, index), local merge."""
        s, i = self.match_local(queries, k, threshold, strict)
        if self.world_size == 1:
            return s, i
        if k == 1:
            ms, mi = exchange_shard_top1(s, i, self.group)
            return ms[:, None], mi[:, None]
        return exchange_shard_topk(s, i, k, self.world_size, self.group)

    def match_sharded_queries(self, local_queries: torch.Tensor, threshold: float = float("-inf"), strict: bool = False):
        """Top-1 for THIS rank's queries against the whole row-sharded gallery: all ranks all-gather their
        raw fp32 queries (every rank must see every query), match them against their shard, exchange the per-shard
        winners with one MAX all-reduce of packed keys, and keep their own slice.  Two collectives, no host sync.
        Returns (scores [q_local,1], global indices [q_local,1])."""
        if self.world_size == 1:
            return self.match_local(local_queries, 1, threshold, strict)
        import torch.distributed as dist
        n = local_queries.shape[0]
        allq = torch.empty((self.world_size * n, local_queries.shape[1]), dtype=local_queries.dtype, device=self.device)
        dist.all_gather_into_tensor(allq, local_queries.contiguous(), group=self.group)
        s, i = self.match_local(allq, 1, threshold, strict)
        ms, mi = exchange_shard_top1(s, i, self.group)
        lo = self.rank * n
        return ms[lo:lo + n, None], mi[lo:lo + n, None]

    # ---- reference-shaped conveniences ------------------------------------------------------------
    def best_match(self, embedding: np.ndarray, similarity_thresh: float) -> Tuple[int, float]:
        """(index or -1, similarity) with the strict '>' semantics of reference main.py:136-142
        (initial best 0, so a match must also be positive)."""
        q = torch.from_numpy(np.asarray(embedding, np.float32).reshape(1, -1)).to(self.device)
        s, i = self.match(q, 1, max(float(similarity_thresh), 0.0), strict=True)
        idx = int(i[0, 0].item())
        return (idx, float(s[0, 0].item())) if idx >= 0 else (-1, 0.0)

    def search_similar(self, query_embedding, k: int = 5, threshold: float = 0.0) -> List[dict]:
        """[{person_id, name, similarity, quality, metadata}] like reference qdrant_manager.py:138-188."""
        q = torch.from_numpy(np.asarray(query_embedding, np.float32).reshape(1, -1)).to(self.device)
        if q.shape[1] != self.dim:
            return []
        out: List[dict] = []
        kk = min(k, KMAX)
        s, i = self.match(q, kk, threshold)
        for sc, ix in zip(s[0].tolist(), i[0].tolist()):
            if ix < 0:
                continue
            row = ix - self.idx_base
            payload = self.payloads[row] if 0 <= row < len(self.payloads) else {}
            pid = self.ids[row] if 0 <= row < len(self.ids) else ix
            out.append({"person_id": payload.get("person_id", pid), "name": payload.get("name", "Unknown"),
                        "similarity": float(sc), "quality": payload.get("quality", 0.0), "metadata": payload})
        return out

    # ---- duplicate merge ---------------------------------------------------------------------------
    def duplicate_pairs(self, threshold: float, row_begin: int = 0, row_end: Optional[int] = None,
                        max_pairs: Optional[int] = None) -> torch.Tensor:
        """Sorted int64 keys (i<<32 | j), i<j, cos(i,j) >= threshold, for i in [row_begin,row_end)."""
        n = len(self)
        row_end = n if row_end is None else row_end
        if n < 2 or row_begin >= row_end:
            return torch.empty(0, dtype=torch.int64, device=self.device)
        cap = int(max_pairs or max(1 << 20, 64 * (row_end - row_begin)))
        while True:
            pairs = torch.empty(cap, dtype=torch.int64, device=self.device)
            count = torch.zeros(1, dtype=torch.int64, device=self.device)
            _lib.check(self.lib.b2f_pairs_threshold(self.h16.data_ptr(), n, self.dim, self.dtype, row_begin, row_end,
                                                    float(threshold), self.f32.data_ptr(), pairs.data_ptr(), cap,
                                                    count.data_ptr(), stream_ptr()), "b2f_pairs_threshold")
            found = int(count.item())
            if found <= cap:
                return torch.sort(pairs[:found]).values
            cap = found

    def merge_duplicates(self, threshold: float) -> np.ndarray:
        """leader[i] for every row (leader[i] == i for survivors), reference duplicate.py:2726-2797 semantics.
        With world_size > 1 the upper triangle is block-partitioned by rows -- equal-area row ranges, one per rank
        (`triangle_range`) -- and the pair lists are exchanged with all_gather before the (cheap, order-dependent)
        resolve, which every rank runs on the same sorted list."""
        n = len(self)
        if self.world_size == 1:
            pairs = self.duplicate_pairs(threshold)
        else:
            import torch.distributed as dist
            # one contiguous row range per rank, cut so that every rank owns the same AREA of the upper triangle
            # (row i meets n - 1 - i columns): one launch per rank instead of one per 4096-row block
            b, e = triangle_range(n, self.rank, self.world_size)
            local = self.duplicate_pairs(threshold, b, e) if e > b else torch.empty(0, dtype=torch.int64, device=self.device)
            cnt = torch.tensor([local.numel()], dtype=torch.int64, device=self.device)
            cnts = [torch.zeros_like(cnt) for _ in range(self.world_size)]
            dist.all_gather(cnts, cnt, group=self.group)
            mx = int(max(c.item() for c in cnts))
            padded = torch.full((max(mx, 1),), -1, dtype=torch.int64, device=self.device)
            padded[:local.numel()] = local
            bufs = [torch.empty_like(padded) for _ in range(self.world_size)]
            dist.all_gather(bufs, padded, group=self.group)
            pairs = torch.sort(torch.cat([b[:int(c.item())] for b, c in zip(bufs, cnts)])).values
        return self.resolve_pairs(pairs)

    def online_clusters(self, grouping_threshold: float, duplicate_threshold: Optional[float] = None,
                        search_threshold: float = 0.0) -> np.ndarray:
        """person (leader row) of every row under the reference's online decision (duplicate.py:1853-1949, processed in
        row order): a row joins the most similar earlier person when cos >= threshold, else founds a new person.
        The founders are exactly the survivors of the greedy leader merge (a row founds a person iff no earlier
        founder reaches the threshold), so: thresholded pairs (tcgen05 GEMM) -> founders (cluster_resolve) -> every
        other row picks its most similar founder among its pairs (exact fp32 dot, earliest founder on ties).
        `duplicate_threshold` (config `duplicate_similarity_threshold`): a row whose best cosine to an earlier person
        reaches it is dropped as a duplicate image (`is_duplicate_image`, duplicate.py:2618-2652) -- label -1.  The
        reference only reaches that check when its database already has the `low_similarity_images` table; None (the
        default) is the fresh-database behaviour.  `search_threshold` is `search_person`'s score floor
        (duplicate.py:1619-1643); a grouping threshold below it cannot admit a row the search did not return."""
        n = len(self)
        if n == 0:
            return np.empty(0, np.int64)
        thr = max(float(grouping_threshold), float(search_threshold))
        pairs = self.duplicate_pairs(thr)
        lowest = torch.from_numpy(self.resolve_pairs(pairs).astype(np.int64)).to(self.device)
        rows = torch.arange(n, device=self.device)
        founder = lowest == rows
        a, b = pairs >> 32, pairs & 0xFFFFFFFF                          # a < b, cos(a, b) >= threshold
        keep = founder[a] & ~founder[b]
        a, b = a[keep], b[keep]
        label = rows.clone()
        if a.numel():
            sim = (self.f32[a] * self.f32[b]).sum(dim=1)
            # per b: highest similarity, then lowest founder index
            order = torch.argsort(a, stable=True)
            order = order[torch.argsort(sim[order], descending=True, stable=True)]
            order = order[torch.argsort(b[order], stable=True)]
            b_s, a_s = b[order], a[order]
            first = torch.ones_like(b_s, dtype=torch.bool)
            first[1:] = b_s[1:] != b_s[:-1]
            label[b_s[first]] = a_s[first]
            if duplicate_threshold is not None:
                label[b_s[first][sim[order][first] >= float(duplicate_threshold)]] = -1
        return label.cpu().numpy()

    def online_similarities(self, label: np.ndarray, search_threshold: float = 0.0) -> np.ndarray:
        """The cosine each online decision was taken on, as the reference records it per visit
        (duplicate.py:1854-1855 `search_results[0]['similarity'] if search_results else 0.0`): a joining row's similarity
        to its person, a founding row's best similarity to the persons founded before it (0 when none reaches
        `search_threshold`, 1.0 for the very first person).  Exact fp32 dots of the stored unit rows; the
        founder-vs-earlier-founder maxima come from one causal pass of the match kernel (`match_local(causal_base=0)`)."""
        n = len(self)
        label_t = torch.as_tensor(np.asarray(label, np.int64), device=self.device)
        rows = torch.arange(n, device=self.device)
        sim = (self.f32 * self.f32[label_t.clamp(min=0)]).sum(dim=1)
        sim[label_t < 0] = 0.0                                                      # skipped as duplicate images
        founders = rows[label_t == rows]
        best = torch.zeros(len(founders), dtype=torch.float32, device=self.device)
        if len(founders) > 1:
            # founder i against the founders before it: one causal (prefix) top-1 pass of the tcgen05 match kernel over a
            # scratch gallery of the founders' unit rows, exact fp32 re-score of the winner included
            scratch = Gallery(self.dim, self.device, self.dtype)
            scratch.add(self.f32[founders])
            s, i = scratch.match_local(scratch.f32, 1, float(search_threshold), causal_base=0)
            best = torch.where(i[:, 0] >= 0, s[:, 0], torch.zeros_like(s[:, 0])).clone()
        if len(founders):
            best[0] = 1.0                                                           # first person: duplicate.py:1826
        sim[founders] = best
        return sim.cpu().numpy()

    def resolve_pairs(self, pairs: torch.Tensor) -> np.ndarray:
        """Greedy one-hop leader merge (ascending id order) over a sorted pair list -> leader[i] per row."""
        n = len(self)
        leader = torch.empty(3 * n + 8, dtype=torch.int32, device=self.device)
        _lib.check(self.lib.b2f_cluster_resolve(pairs.data_ptr(), pairs.numel(), n, leader.data_ptr(), stream_ptr()),
                   "b2f_cluster_resolve")
        return leader[:n].cpu().numpy()


def row_blocks(n: int, world_size: int, block: int = 4096) -> List[Tuple[int, int, int]]:
    """(rank, row_begin, row_end) for the block-partitioned upper triangle: early row blocks see more
    columns than late ones, so blocks are dealt cyclically for balance."""
    out = []
    for bi, start in enumerate(range(0, n, block)):
        out.append((bi % world_size, start, min(start + block, n)))
    return out


def triangle_range(n: int, rank: int, world_size: int, align: int = 256) -> Tuple[int, int]:
    """[begin, end) rows of the all-pairs upper triangle owned by `rank`: row i is compared with the n - 1 - i rows
    after it, so equal work means equal area, begin_r = n (1 - sqrt(1 - r / P)), rounded to the 256-row tiles of the
    pair kernel.  The ranges of all ranks cover [0, n) exactly once."""
    def cut(r: int) -> int:
        if r <= 0:
            return 0
        if r >= world_size:
            return n
        x = n * (1.0 - (1.0 - r / world_size) ** 0.5)
        return min(n, int(round(x / align)) * align)
    return cut(rank), cut(rank + 1)


def shard_range(total: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous [begin, end) slice of `total` items owned by `rank` (frames, crops, gallery rows)."""
    per = (total + world_size - 1) // world_size
    return min(rank * per, total), min((rank + 1) * per, total)
