"""Times the matching / all-pairs kernels alone (CUDA events, inputs resident): python tools/match_bench.py [--gen 0|1]"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scrfd_arcface_facerecognition_b200 import _lib  # noqa: E402
from scrfd_arcface_facerecognition_b200.gallery import Gallery  # noqa: E402


def ms_of(fn, reps):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gen", type=int, default=1)
    ap.add_argument("--pairs-rows", type=int, default=200_000)
    a = ap.parse_args()
    _lib.call("b2f_set_tuning", 17, a.gen)
    gen = torch.Generator(device="cuda").manual_seed(1)
    out = {"generation": a.gen}
    for q, g in ((1024, 1_000_000), (2048, 500_000), (8192, 125_000), (100_000, 125_000)):
        G = Gallery()
        G.set_shard(torch.randn((g, 512), generator=gen, device="cuda"), 0)
        qs = torch.randn((q, 512), generator=gen, device="cuda")
        qf, qh = G._normalise(qs)
        splits = int(G.lib.b2f_match_plan(q, g))
        sc = G._scratch_for(q, splits)

        def gemm():
            _lib.check(G.lib.b2f_match_partial_keep(qh.data_ptr(), q, G.h16.data_ptr(), g, 512, G.dtype, None, None, 8, 3, splits,
                                                    sc["ps"].data_ptr(), sc["pi"].data_ptr(), torch.cuda.current_stream().cuda_stream))
        t_gemm = ms_of(gemm, 10 if q < 50_000 else 3)
        t_all = ms_of(lambda: G.match_local(qs, 1, 0.4, strict=True), 10 if q < 50_000 else 3)
        fl = 2.0 * q * g * 512
        out[f"{q}x{g}"] = dict(splits=splits, gemm_ms=t_gemm, gemm_tflops=fl / t_gemm / 1e9, match_ms=t_all, match_tflops=fl / t_all / 1e9)
        del G
    n = a.pairs_rows
    c = torch.nn.functional.normalize(torch.randn((n // 4, 512), generator=gen, device="cuda"), dim=1).repeat_interleave(4, dim=0)
    x = c + 0.35 * torch.nn.functional.normalize(torch.randn((n, 512), generator=gen, device="cuda"), dim=1)
    G = Gallery()
    G.set_shard(x[torch.randperm(n, generator=gen, device="cuda")], 0)
    t = ms_of(lambda: G.duplicate_pairs(0.8), 3)
    out[f"pairs_{n}"] = dict(ms=t, tflops=n * (n - 1) * 512.0 / t / 1e9)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
