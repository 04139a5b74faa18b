"""Diagnostic: determinism of the batched pipeline and correctness of 1M-row top-1 on planted rows."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from models import SCRFD, ArcFace
from scrfd_arcface_facerecognition_b200.gallery import Gallery
from scrfd_arcface_facerecognition_b200.pipeline import FacePipeline
from scrfd_arcface_facerecognition_b200 import _lib

dev = torch.device("cuda", 0)
G = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
# 1. pure match test with synthetic queries
gal = Gallery()
gen = torch.Generator(device=dev).manual_seed(2)
gal.set_shard(torch.randn((G, 512), generator=gen, device=dev), 0)
q = torch.randn((1024, 512), generator=gen, device=dev) * 5
rows = torch.from_numpy(np.random.default_rng(7).permutation(G)[:1024]).to(dev)
gal.replace_rows(rows, q)
for trial in range(3):
    s, i = gal.match(q, 1, 0.4, strict=True)
    torch.cuda.synchronize()
    ok = (i.reshape(-1) == rows)
    print(f"synthetic queries trial {trial}: top1 correct {int(ok.sum())}/1024, min score {float(s.min()):.4f}")
    if not ok.all():
        bad = torch.nonzero(~ok).reshape(-1)[:8]
        for b in bad.tolist():
            print("   query", b, "expected row", int(rows[b]), "(tile", int(rows[b]) // 256, "col", int(rows[b]) % 256, ") got", int(i[b, 0]), "score", float(s[b, 0]))
        r = rows[~ok].cpu().numpy()
        print("   failing rows: tile%2 hist", np.bincount((r // 256) % 2, minlength=2), " split", np.bincount((r // 256) // 106, minlength=37))
# 2. pipeline determinism
det, rec = SCRFD("weights/det_10g.onnx"), ArcFace("weights/w600k_r50.onnx")
pipe = FacePipeline(det, rec, None, max_num=16)
frames = torch.from_numpy(np.random.default_rng(1000).integers(0, 256, (16, 1080, 1920, 3), dtype=np.uint8)).to(dev)
outs = []
for mode in (1, 1, 0):
    _lib.check(_lib.lib().b2f_set_tuning(2, mode))
    o = pipe.process(frames)
    torch.cuda.synchronize()
    outs.append({k: v.clone() for k, v in o.items()})
for a, b, name in ((0, 1, "persistent vs persistent"), (0, 2, "persistent vs per-tile kernel")):
    for k in ("det", "kps", "emb"):
        d = (outs[a][k] - outs[b][k]).abs().max().item()
        print(f"{name}: {k} max abs diff {d:.3e}")
e = outs[0]["emb"]
en = e / e.norm(dim=1, keepdim=True)
sim = en @ en.T
sim.fill_diagonal_(0)
print("max off-diagonal cosine between face embeddings:", float(sim.max()), " counts", outs[0]["counts"][:, 0].tolist())
