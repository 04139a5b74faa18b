"""Multi-GPU parity check, launched under torchrun (not collected by pytest: it needs N > 1 GPUs of one box):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu_check.py
Every rank checks, over NCCL:
  * row-sharded gallery: `Gallery.match_sharded_queries` / `Gallery.match` (k = 1 and k = 5) == the unsharded oracle answer
    (restate.search_similar / best_match semantics: score desc, index asc), identical on every rank;
  * block-partitioned clustering: `Gallery.merge_duplicates` on N ranks == oracle.restate.merge_duplicates (and == the
    reference-run golden leaders of tests/golden/cluster_outputs.npz).
Writes gpurun_out/multi_gpu_check.json on rank 0 and exits non-zero on any mismatch."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("B2F_SYNTHETIC_WEIGHTS", "1")

from oracle import restate  # noqa: E402
from scrfd_arcface_facerecognition_b200.gallery import Gallery, shard_range  # noqa: E402
from tests.golden import inputs  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    report = {"world": world}

    # ---- sharded matching ---------------------------------------------------------------------------------
    gal = inputs.embeddings(20, 70_001)
    qs, ids = inputs.planted_queries(gal, 21, 260)
    gn, qn = restate.normalize_rows(gal).astype(np.float64), restate.normalize_rows(qs).astype(np.float64)
    sims = qn @ gn.T
    order = np.argsort(-sims, axis=1, kind="stable")
    b, e = shard_range(len(gal), rank, world)
    G = Gallery(rank=rank, world_size=world)
    G.set_shard(torch.from_numpy(gal[b:e]).to(dev), b)
    s1, i1 = G.match(torch.from_numpy(qs).to(dev), 1, 0.4, strict=True)
    ok1 = bool((i1[:, 0].cpu().numpy() == ids).all() and (i1[:, 0].cpu().numpy() == order[:, 0]).all())
    ok1 = ok1 and float(np.abs(s1[:, 0].cpu().numpy() - sims[np.arange(len(qs)), order[:, 0]]).max()) <= 2e-6
    s5, i5 = G.match(torch.from_numpy(qs).to(dev), 5)
    top = np.take_along_axis(sims, order[:, :6], 1)
    gap_ok = (top[:, :5] - top[:, 1:6]) >= 1e-3
    gap_ok[:, 1:] &= gap_ok[:, :-1]
    ok5 = bool((i5.cpu().numpy() == order[:, :5])[gap_ok].all())
    # each rank owns a slice of the queries (frames are sharded): own slice back, global ids
    per = len(qs) // world
    mine = torch.from_numpy(qs[rank * per:(rank + 1) * per]).to(dev)
    sm, im = G.match_sharded_queries(mine, 0.4, strict=True)
    okq = bool((im[:, 0].cpu().numpy() == ids[rank * per:(rank + 1) * per]).all())
    report["match"] = dict(top1=ok1, top5=ok5, own_queries=okq)

    # ---- block-partitioned clustering -----------------------------------------------------------------------
    golden = dict(np.load(os.path.join(ROOT, "tests", "golden", "cluster_outputs.npz")))
    thr = float(golden["thresholds"][3])
    okc = {}
    for name, rows in inputs.cluster_cases():
        C = Gallery(rank=rank, world_size=world)
        C.set_shard(torch.from_numpy(rows).to(dev), 0)
        leader = C.merge_duplicates(thr)
        okc[name] = bool((leader == golden[f"merge_{name}_leader"]).all())
    big = inputs.clustered(30, 3000, 4)                   # 12 000 rows: three 4096-row blocks dealt over the ranks
    C = Gallery(rank=rank, world_size=world)
    C.set_shard(torch.from_numpy(big).to(dev), 0)
    leader = C.merge_duplicates(0.8)
    okc["12k_rows_vs_oracle"] = bool((leader == restate.merge_duplicates(big, 0.8)).all())
    report["cluster"] = okc

    flat = [v for d in (report["match"], report["cluster"]) for v in d.values()]
    ok = torch.tensor([int(all(flat))], device=dev)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    report["all_ranks_ok"] = bool(ok.item())
    if rank == 0:
        print(json.dumps(report))
        out_dir = os.path.join(ROOT, "gpurun_out")
        if os.path.isdir(out_dir):
            json.dump(report, open(os.path.join(out_dir, "multi_gpu_check.json"), "w"), indent=1)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if report["all_ranks_ok"] else 1)


if __name__ == "__main__":
    main()
