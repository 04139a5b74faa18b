"""Time the first ops of the ArcFace engine in sequence vs each repeated back to back (same buffers)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from scrfd_arcface_facerecognition_b200 import ArcFace, SCRFD, _lib
from scrfd_arcface_facerecognition_b200.engine import stream_ptr

which = sys.argv[1] if len(sys.argv) > 1 else "rec"
if which == "rec":
    m = ArcFace("weights/w600k_r50.onnx"); eng = m._engine; n = 1024
else:
    m = SCRFD("weights/det_10g.onnx"); eng = m._engine_for(640, 640); n = 64
eng.input_buffer(n).normal_()
for _ in range(2):
    eng.run(n)
torch.cuda.synchronize()
bound = eng._bound[n][0]
sp = stream_ptr()
ev = lambda: torch.cuda.Event(enable_timing=True)
nops = int(sys.argv[2]) if len(sys.argv) > 2 else 10
seq = []
tl = []
eng.run(n, timings=tl)
torch.cuda.synchronize()
for i, kind, e0, e1 in tl[:nops]:
    seq.append(e0.elapsed_time(e1))
for i in range(nops):
    b = bound[i]
    for _ in range(3):
        b.fn(*b.args, sp)
    e0, e1 = ev(), ev()
    e0.record()
    for _ in range(10):
        b.fn(*b.args, sp)
    e1.record()
    torch.cuda.synchronize()
    at = eng.plan.ops[i].attrs
    print(f"{which} op {i} {eng.plan.ops[i].kind} {at.get('cin',0)}->{at.get('cout',0)} k{at.get('kh',0)} s{at.get('stride',0)} {at.get('h',0)}: in sequence {seq[i]*1e3:8.1f} us   repeated {e0.elapsed_time(e1)*100:8.1f} us")
