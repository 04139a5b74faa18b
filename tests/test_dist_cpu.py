"""CPU, world_size 2 over gloo: the host-side sharding logic of the multi-GPU path --
gallery row shards -> per-shard top-k -> all_gather -> merge must equal the single-shard answer,
and the clustering row blocks must cover the upper triangle exactly once."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from scrfd_arcface_facerecognition_b200.gallery import exchange_shard_topk, merge_shard_topk, row_blocks, shard_range
from tests.golden import inputs


def _topk_rows(q, g, k, base):
    s = q @ g.T
    sc, ix = torch.sort(s, dim=1, descending=True, stable=True)
    return sc[:, :k].contiguous(), (ix[:, :k] + base).contiguous()


def _worker(rank, world, port, q_np, g_np, k, ret):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    q, g = torch.from_numpy(q_np), torch.from_numpy(g_np)
    b, e = shard_range(len(g), rank, world)
    s, i = _topk_rows(q, g[b:e], k, b)
    ms, mi = exchange_shard_topk(s, i, k, world)              # the collective Gallery.match runs (NCCL on the GPU box)
    if rank == 0:
        ret["s"], ret["i"] = ms.numpy(), mi.numpy()
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_sharded_topk_equals_single_shard():
    g = inputs.embeddings(70, 301)
    g /= np.linalg.norm(g, axis=1, keepdims=True)
    q, ids = inputs.planted_queries(g, 71, 9)
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    k = 5
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), q, g, k, ret), nprocs=2, join=True)
    s, i = _topk_rows(torch.from_numpy(q), torch.from_numpy(g), k, 0)
    np.testing.assert_array_equal(ret["i"], i.numpy())
    np.testing.assert_array_equal(ret["s"], s.numpy())
    np.testing.assert_array_equal(ret["i"][:, 0], ids)
    ret1 = mgr.dict()                                   # k = 1 goes through the max / masked-min merge
    mp.spawn(_worker, args=(2, _free_port(), q, g, 1, ret1), nprocs=2, join=True)
    np.testing.assert_array_equal(ret1["i"][:, 0], ids)
    np.testing.assert_array_equal(ret1["s"][:, 0], s.numpy()[:, 0])


def test_merge_handles_empty_slots_and_ties():
    s = torch.tensor([[[0.9, 0.5], [0.9, 0.0]]]).permute(1, 0, 2)          # [P=2, Q=1, k=2]
    i = torch.tensor([[[7, 3], [2, -1]]]).permute(1, 0, 2)
    ms, mi = merge_shard_topk(s, i, 2)
    assert mi.tolist() == [[2, 7]] and torch.allclose(ms, torch.tensor([[0.9, 0.9]]))          # equal scores: lower index first
    ms, mi = merge_shard_topk(s, i, 4)
    assert mi.tolist() == [[2, 7, 3, -1]]


def test_top1_merge_equals_general_merge():
    from scrfd_arcface_facerecognition_b200.gallery import merge_shard_top1, merge_shard_topk
    g = torch.Generator().manual_seed(5)
    p, q = 8, 500
    s = torch.rand((p, q), generator=g)
    i = torch.randint(0, 10_000, (p, q), generator=g)
    s[:, 10:60] = s[0, 10:60]                           # equal scores across shards: lowest index wins
    i[torch.rand((p, q), generator=g) < 0.2] = -1       # empty slots never win
    i[:, 100:110] = -1                                  # queries with no match anywhere
    s1, i1 = merge_shard_top1(s, i)
    sk, ik = merge_shard_topk(s[:, :, None], i[:, :, None], 1)
    assert torch.equal(i1, ik[:, 0]) and torch.equal(s1, sk[:, 0])
    assert (i1[100:110] == -1).all() and (s1[100:110] == 0).all()


def test_row_blocks_partition():
    for n, world in ((10000, 2), (4096 * 3 + 5, 4), (100, 8)):
        blocks = row_blocks(n, world)
        covered = np.zeros(n, int)
        for r, b, e in blocks:
            assert 0 <= r < world
            covered[b:e] += 1
        assert (covered == 1).all()
    from scrfd_arcface_facerecognition_b200.gallery import triangle_range
    for n, world in ((200_000, 8), (12_000, 2), (300, 4), (5, 8)):
        cuts = [triangle_range(n, r, world) for r in range(world)]
        assert cuts[0][0] == 0 and cuts[-1][1] == n and all(cuts[i][1] == cuts[i + 1][0] for i in range(world - 1))
        if n >= 100_000:                                   # equal areas of the upper triangle within the 256-row rounding (a few %)
            area = [sum(n - 1 - i for i in range(b, e, 64)) for b, e in cuts]
            assert max(area) <= 1.06 * min(area)
    assert shard_range(10, 0, 4) == (0, 3) and shard_range(10, 3, 4) == (9, 10) and shard_range(2, 3, 4) == (2, 2)


def _worker_top1(rank, world, port, s_np, i_np, ret):
    from scrfd_arcface_facerecognition_b200.gallery import exchange_shard_top1
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ms, mi = exchange_shard_top1(torch.from_numpy(s_np[rank]), torch.from_numpy(i_np[rank]))   # ONE MAX all-reduce of packed keys
    ret[rank] = (ms.numpy(), mi.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_top1_exchange_is_one_max_allreduce_of_packed_keys():
    """the k = 1 collective of the sharded match (what bench.py runs under NCCL): packed (score, index) keys, MAX
    all-reduce, unpack == merge by (score desc, index asc) of the gathered lists, on every rank"""
    from scrfd_arcface_facerecognition_b200.gallery import merge_shard_top1, pack_top1_keys, unpack_top1_keys
    g = torch.Generator().manual_seed(6)
    p, q = 2, 300
    s = torch.rand((p, q), generator=g) * 2 - 1
    i = torch.randint(0, 1_000_000, (p, q), generator=g)
    s[1, 10:60] = s[0, 10:60]                           # equal scores across shards: lowest index wins
    i[torch.rand((p, q), generator=g) < 0.2] = -1       # empty slots never win
    i[:, 100:110] = -1
    i[0, 0], i[1, 0] = 2 ** 32 - 2, 5                   # the largest representable global index
    rs, ri = unpack_top1_keys(pack_top1_keys(s[0], i[0]))
    assert torch.equal(ri, i[0]) and torch.equal(rs[i[0] >= 0], s[0][i[0] >= 0])
    ret = mp.Manager().dict()
    mp.spawn(_worker_top1, args=(p, _free_port(), s.numpy(), i.numpy(), ret), nprocs=p, join=True)
    want_s, want_i = merge_shard_top1(s, i)
    for rank in range(p):
        np.testing.assert_array_equal(ret[rank][1], want_i.numpy())
        np.testing.assert_array_equal(ret[rank][0], want_s.numpy())
