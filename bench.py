#!/usr/bin/env python
"""bench.py -- benchmarks of the detect -> align -> embed -> match hot path (BASELINE.json configs[1..4]).

    python bench.py --gpus N --steps K --warmup W              headline: configs[1] (--config 2, the default)
    python bench.py --config 3|4|5 --gpus N ...                configs[2..4]
    python bench.py --impl reference [--config C] ...          the reference's CPU path on the host cores
N > 1: launched under torchrun, one rank per GPU (RANK / LOCAL_RANK / WORLD_SIZE / MASTER_* from the environment).

  --config 2  SCRFD-10G + ArcFace-R50, 64 synthetic 1920x1080 frames per GPU per step, max_num = 16 (1024 faces),
              top-1 against a 1M x 512 gallery (rows sharded over the ranks).                      metric: faces/s
  --config 3  ArcFace-R50 embedding of 100k synthetic aligned 112x112 crops (split over the ranks) + cosine top-1
              against a 1M x 512 gallery (rows sharded).                         metric: faces/s; match TFLOP/s at Q = 100k
  --config 4  duplicate.py-style all-pairs cosine clustering of 200k embeddings at 0.8, upper-triangle row blocks
              dealt to the ranks.                                                                  metric: pairs/s
  --config 5  SCRFD-2.5G + ArcFace-R50 video loop, 1080p frames dealt to the ranks, max_num = 50, a 4096-row target
              gallery.                                                     metric: faces/s (+ frames/s against 24 fps)

Every run prints ONE JSON line (rank 0).  `value` is device-resident throughput (inputs already in HBM), `e2e.value`
goes through pinned host buffers with the H2D / D2H copies inside the timed region.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# weights/*.onnx cannot be fetched offline: the benchmark opts in to seeded random weights of the same architectures
# (the product raises FileNotFoundError on a missing model file otherwise) and says so in `config.weights`
os.environ.setdefault("B2F_SYNTHETIC_WEIGHTS", "1")

METRICS = {
    2: ("end-to-end faces/sec SCRFD-10G+ArcFace-R50 (detect->align->embed->match vs 1M gallery)", "faces/s"),
    3: ("faces/sec ArcFace-R50 embed of 100k aligned crops + cosine top-1 vs 1M gallery (match TFLOP/s reported beside it)", "faces/s"),
    4: ("all-pairs cosine clustering of 200k embeddings (duplicate.py find_and_merge_duplicates): pair comparisons/sec", "pairs/s"),
    5: ("end-to-end faces/sec SCRFD-2.5G+ArcFace-R50 video loop, 1080p, max_num=50 (frames/s vs 24 fps reported beside it)", "faces/s"),
}
DEFAULTS = {   # per-config defaults of the size flags
    2: dict(frames=64, faces=16, gallery=1_000_000, det="weights/det_10g.onnx", steps=10),
    3: dict(crops=100_000, gallery=1_000_000, steps=3),
    4: dict(rows=200_000, steps=3),
    5: dict(frames=32, faces=50, gallery=4096, det="weights/det_2.5g.onnx", steps=10),
}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, default=2, choices=[2, 3, 4, 5], help="BASELINE.json configs[config-1]")
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--frames", type=int, default=None, help="frames per GPU per step (configs 2, 5)")
    ap.add_argument("--faces", type=int, default=None, help="max_num: face slots per frame (configs 2, 5)")
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--gallery", type=int, default=None, help="total gallery rows (sharded over ranks in configs 2, 3)")
    ap.add_argument("--crops", type=int, default=None, help="config 3: aligned crops in the whole job")
    ap.add_argument("--rows", type=int, default=None, help="config 4: embeddings to cluster")
    ap.add_argument("--det", default=None)
    ap.add_argument("--rec", default="weights/w600k_r50.onnx")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--tail", default="inline", choices=["overlap", "inline"],
                    help="N > 1, config 2: run step i's sharded match in line (default; measured faster: the persistent kernels of the "
                         "nets and of the match each want every SM) or on a side stream under step i+1's nets")
    ap.add_argument("--layer-report", default=None, help="write a per-launch table of the conv nets to this path")
    a = ap.parse_args()
    for k, v in DEFAULTS[a.config].items():
        if getattr(a, k, None) is None:
            setattr(a, k, v)
    return a


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return dict(tensor=d.get("bf16_tflops_sustained", 1388.5), tensor_burst=d.get("bf16_tflops", 1636.6),
                    hbm=d.get("hbm_gbs", 6546.6),
                    tensor_src="measured (MEASURED_PEAKS.json bf16_tflops_sustained: kernel timed inside a long step)",
                    hbm_src="measured (MEASURED_PEAKS.json hbm_gbs)")
    return dict(tensor=1400.0, tensor_burst=1650.0, hbm=6500.0, tensor_src="fallback (B200_PROFILING.md: ~1.4 PFLOP/s sustained)",
                hbm_src="fallback (B200_PROFILING.md)")


# ---------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi during the timed region)
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock + throttle reasons sampled DURING the timed region: NVML polled every 5 ms from a thread
    (nvidia-smi -lms as the fallback when the NVML binding is unavailable)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []
        self.samples, self.max_mhz, self.reason_bits, self._stop, self._thr, self.nvml = [], None, 0, False, None, None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(self._physical_index())
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self._thr = threading.Thread(target=self._poll, daemon=True)
            self._thr.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _physical_index(self):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            ids = [v.strip() for v in vis.split(",") if v.strip()]
            if self.index < len(ids) and ids[self.index].isdigit():
                return int(ids[self.index])
        return self.index

    def _poll(self):
        n = self.nvml
        while not self._stop:
            try:
                self.samples.append(float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)))
                self.reason_bits |= int(n.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
            except Exception:
                try:
                    self.reason_bits |= int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
                except Exception:
                    pass
            time.sleep(0.005)

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.nvml is not None:
            self._stop = True
            self._thr.join(timeout=1.0)
            names = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}
            reasons = sorted(v for k, v in names.items() if self.reason_bits & k)
            sm = self.samples
            return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": self.max_mhz, "reasons": reasons,
                    "samples": len(sm), "source": "nvml"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx = float(parts[1])
            except ValueError:
                continue
            for nm, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm), "source": "nvidia-smi"}


# ---------------------------------------------------------------------------------------------
# distributed plumbing shared by every config
# ---------------------------------------------------------------------------------------------
class Job:
    """One process per GPU: device, rank / world, barrier + CUDA-event timing with the maximum over the ranks."""

    def __init__(self, a):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if a.gpus > 1 and self.world != a.gpus:
            raise SystemExit(f"--gpus {a.gpus} needs torchrun with {a.gpus} ranks (WORLD_SIZE={self.world})")
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            # NCCL announces its version on stdout when the first communicator comes up: keep stdout for the JSON line
            sys.stdout.flush()
            saved = os.dup(1)
            os.dup2(2, 1)
            try:
                dist.init_process_group("nccl", device_id=self.dev)
                dist.all_reduce(torch.zeros(1, device=self.dev))
                torch.cuda.synchronize()
            finally:
                sys.stdout.flush()
                os.dup2(saved, 1)
                os.close(saved)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def timed(self, fn, steps, finish=None):
        """milliseconds for `steps` calls of fn(i) (+ finish()), device-timed, max over ranks, barrier + sync both sides"""
        torch = self.torch
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        if finish is not None:
            finish()
        e1.record()
        self.barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(ms, op=self.dist.ReduceOp.MAX)
        return float(ms.item())

    def max_over_ranks(self, v: float) -> float:
        if self.world == 1:
            return v
        t = self.torch.tensor([v], device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def finish(self):
        if self.world > 1:
            self.dist.barrier()
            self.dist.destroy_process_group()


def event_ms(torch, fn, reps=1):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def memory_rooflines(spans, pk):
    """[(kernel, algorithmic bytes, ms)] -> per-kernel HBM roofline entries (CUDA events inside this run)."""
    agg = {}
    for name, nbytes, ms in spans:
        b, t, n = agg.get(name, (0.0, 0.0, 0))
        agg[name] = (b + nbytes, t + ms, n + 1)
    out = []
    for name, (b, t, n) in agg.items():
        gbs = b / (t / 1e3) / 1e9 if t > 0 else 0.0
        out.append({"kernel": name, "bound": "hbm", "launches": n, "bytes_per_launch": b / n, "us_per_launch": t / n * 1e3,
                    "achieved": gbs, "peak": pk["hbm"], "unit": "GB/s", "frac": gbs / pk["hbm"]})
    return out


def net_launch_rows(torch, eng, n, first_launch=None, start=0):
    """One eager pass of a conv net with CUDA events around every launch -> [(op index, kind, flops, bytes, ms)]."""
    tl = []
    # the host needs ~10-20 us per launch (ctypes call, plan, tensor maps) and many layers run for less: a few ms of spin
    # on the stream first lets the host queue the whole pass ahead of the device, so an event pair brackets the kernel's
    # execution and not the wait for its launch to arrive
    torch.cuda._sleep(int(6e6))
    if first_launch is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        first_launch()
        e1.record()
        tl.append((1, "conv", e0, e1))
    eng.run(n, timings=tl, start=start)
    torch.cuda.synchronize()
    rows = []
    for i, kind, e0, e1 in tl:
        op = eng.plan.ops[i]
        si, so = eng.plan.tensors[op.src], eng.plan.tensors[op.dst]
        nbytes = n * (si.h * si.w * si.cp * 2 + so.h * so.w * so.cp * (4 if so.f32 else 2))
        rows.append((i, kind, eng.op_flops(i, n), nbytes, e0.elapsed_time(e1)))
    return rows


def conv_roofline(a, nets, pk):
    """Time every launch of the tensor-core conv kernels in eager passes of the nets (CUDA events on the launching
    stream, launches queued ahead of the device, fastest of three repetitions per launch) and relate their algorithmic FLOPs to the measured cuBLAS bf16 peak.
    nets: [(tag, engine, batch, first_launch or None, start op)].  Also returns the HBM entries of the nets' pooling ops."""
    import torch
    per_net = []
    for rep in range(3):                            # per launch, the fastest of three repetitions
        cur = [(tag, eng, n, net_launch_rows(torch, eng, n, first, start)) for tag, eng, n, first, start in nets]
        if per_net:
            cur = [(tag, eng, n, [r if r[4] <= q[4] else q for r, q in zip(rows, prev)])
                   for (tag, eng, n, rows), (_, _, _, prev) in zip(cur, per_net)]
        per_net = cur
    conv = [(fl, ms) for _, _, _, rows in per_net for _, kind, fl, _, ms in rows if kind == "conv"]
    pools = [(f"{kind}_kernel ({tag} op {i})", nb, ms) for tag, _, _, rows in per_net for i, kind, _, nb, ms in rows
             if kind in ("pool", "dwconv", "eltwise", "im2col")]
    if a.layer_report:
        with open(a.layer_report, "w") as f:
            f.write("net,op,kind,cin,cout,k,stride,h,w,ms,tflops,gbs\n")
            for tag, eng, n, rows in per_net:
                for i, kind, fl, nb, ms in rows:
                    at = eng.plan.ops[i].attrs
                    f.write(f"{tag},{i},{kind},{at.get('cin', at.get('c', 0))},{at.get('cout', 0)},{at.get('kh', at.get('k', 0))},"
                            f"{at.get('stride', 0)},{at.get('h', 0)},{at.get('w', 0)},{ms:.4f},{fl / ms / 1e9 if ms else 0:.1f},"
                            f"{nb / ms / 1e6 if ms else 0:.0f}\n")
    flops, ms = sum(r[0] for r in conv), sum(r[1] for r in conv)
    achieved = flops / (ms / 1e3) / 1e12
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "conv_traffic.json")
    if os.path.exists(tpath):                       # dram bytes per conv launch from the committed ncu capture
        t = json.load(open(tpath))
        traffic = t.get("dram_bytes_per_launch")
        traffic_src = t.get("source", "profiles/conv_traffic.json: static, from a committed `ncu --set full` capture of this workload (not measured in this run)")
    roof = {"bound": "tensor", "kernel": "conv_tile_kernel<CG2> + umma_conv_persistent_kernel (all conv/FC launches of " +
            " + ".join(t for t, *_ in nets) + ")",
            "achieved": achieved, "peak": pk["tensor"], "unit": "TFLOP/s", "frac": achieved / pk["tensor"], "traffic": traffic,
            "traffic_source": traffic_src, "flops_per_launch": flops / max(len(conv), 1), "ms_per_launch": ms / max(len(conv), 1),
            "peak_source": pk["tensor_src"], "launches_per_step": len(conv), "flops_per_step": flops, "conv_ms_per_step": ms}
    return roof, pools


# ---------------------------------------------------------------------------------------------
# reference CPU path (the reference's own Python when its files are present -- /root/reference here, baseline/_ref on
# the GPU box -- else the restated oracle port)
# ---------------------------------------------------------------------------------------------
CPU_NOTE = "reference CPU path, ORT CPUExecutionProvider substituted by torch-CPU fp32"


def make_cpu_state(a, gallery_rows: int):
    import cv2
    import torch
    from oracle import ref_loader, restate
    from oracle.torch_exec import TorchGraph
    from scrfd_arcface_facerecognition_b200 import archs
    cores = len(os.sched_getaffinity(0))
    torch.set_num_threads(cores)
    cv2.setNumThreads(cores)
    ref = ref_loader.load()
    rng = np.random.default_rng(0)
    frames = [rng.integers(0, 256, (a.height, a.width, 3), dtype=np.uint8) for _ in range(2)]
    gal = np.random.default_rng(2).standard_normal((gallery_rows, 512), dtype=np.float32) if gallery_rows else None
    if gal is not None:
        gal /= np.linalg.norm(gal, axis=1, keepdims=True)
    faces = getattr(a, "faces", None) or 16
    if ref is not None:                     # the reference's own classes, verbatim, over the torch-CPU session shim
        r = ref.ArcFace(a.rec)
        rec = lambda f, k: r(f, k)
        get_feat = lambda crops: r.get_feat(crops)
        det = None
        if a.det:
            d = ref.SCRFD(a.det)
            det = lambda f: d.detect(f, max_num=faces)
        kind = "reference"
    else:                                    # no reference files -> restated oracle around the same net
        rg = TorchGraph(archs.build_arch(archs.arch_for_path(a.rec)))
        det = None
        if a.det:
            dg = TorchGraph(archs.build_arch(archs.arch_for_path(a.det)))

            def det(f):
                canvas, ds = restate.letterbox_u8(f, 640, 640)
                out = dg.run(restate.blob_from_bgr(canvas, 1 / 128, 127.5))
                return restate.scrfd_postprocess([out[n] for n in dg.output_names], 640, 640, ds, 0.5, 0.4, faces, "max",
                                                 f.shape[:2])

        def rec(f, k):
            M = restate.estimate_norm_closed_form(k)
            crop = cv2.warpAffine(f, M, (112, 112), borderValue=0.0)
            return rg.run(restate.blob_from_bgr(crop, 1 / 127.5, 127.5))[rg.output_names[0]].reshape(-1)

        def get_feat(crops):
            return rg.run(restate.blob_from_bgr(np.stack(crops), 1 / 127.5, 127.5))[rg.output_names[0]].reshape(len(crops), -1)
        kind = "port"
    return dict(det=det, rec=rec, get_feat=get_feat, gallery=gal, frames=frames, cores=cores, kind=kind, ref=ref)


def cpu_match(e, gal_n):
    """cosine top-1 of the reference loop (main.py:136-142) as one BLAS product per batch of faces"""
    e = e / np.linalg.norm(e, axis=1, keepdims=True)
    sims = e @ gal_n.T
    return sims.argmax(1), sims.max(1)


def cpu_pipeline_step(a, state, n_frames: int):
    """configs 2 / 5: n_frames frames -> detect -> align -> embed (one call per face, main.py:134) -> match."""
    faces = 0
    for i in range(n_frames):
        f = state["frames"][i % len(state["frames"])]
        boxes, kpss = state["det"](f)
        embs = [state["rec"](f, k) for k in kpss]
        if embs:
            cpu_match(np.stack(embs), state["gallery"])
        faces += len(embs)
    return faces


def cpu_sample(a, state):
    """(units done, description) of ONE bounded sample of the configured workload on the host cores."""
    if a.config in (2, 5):
        n = cpu_pipeline_step(a, state, 1)
        nets = "SCRFD-10G" if a.config == 2 else "SCRFD-2.5G"
        return n, (f"1 frame {a.width}x{a.height} per step ({nets} torch-CPU fp32 + up to {a.faces} ArcFace-R50 faces, one call "
                   f"per face + cosine top-1 vs {len(state['gallery'])} gallery rows)")
    if a.config == 3:
        crops = state.setdefault("crops", [np.random.default_rng(5 + i).integers(0, 256, (112, 112, 3), dtype=np.uint8)
                                            for i in range(16)])
        e = state["get_feat"](crops)
        cpu_match(e, state["gallery"])
        return len(crops), (f"16 aligned crops per step (ArcFace-R50 torch-CPU fp32, get_feat batch of 16) + cosine top-1 vs "
                            f"{len(state['gallery'])} gallery rows")
    # config 4: the reference's find_and_merge_duplicates is N searches of N (restated: oracle.restate.merge_duplicates)
    from oracle import restate
    from tests.golden import inputs
    n = 3000
    emb = state.setdefault("emb4", inputs.clustered(3, n // 4, 4))
    restate.merge_duplicates(emb, 0.8)
    return n * (n - 1) // 2, (f"{n} x 512 embeddings per step ({n * (n - 1) // 2} pair comparisons; greedy leader merge at 0.8, "
                              "numpy fp64 products -- the reference issues one Qdrant search per person)")


def cpu_state_for(a):
    if a.config == 4:
        import torch
        cores = len(os.sched_getaffinity(0))
        torch.set_num_threads(cores)
        return dict(cores=cores, kind="port")
    gallery_rows = a.gallery
    if a.config == 3:
        a.det = None
    return make_cpu_state(a, gallery_rows)


def cpu_baseline(a, budget_s=12.0, max_samples=8):
    state = cpu_state_for(a)
    cpu_sample(a, state)
    t0 = time.perf_counter()
    n, units, what = 0, 0, ""
    while time.perf_counter() - t0 < budget_s and n < max_samples:
        u, what = cpu_sample(a, state)
        units += u
        n += 1
    dt = time.perf_counter() - t0
    return {"value": units / dt, "unit": METRICS[a.config][1], "cores": state["cores"], "kind": state["kind"],
            "sample": f"{n} x [{what}]", "note": CPU_NOTE}


def run_reference_arm(a):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    metric, unit = METRICS[a.config]
    state = cpu_state_for(a)
    what = ""
    for _ in range(max(min(a.warmup, 2), 1)):
        _, what = cpu_sample(a, state)
    t0 = time.perf_counter()
    units = 0
    for _ in range(a.steps):
        units += cpu_sample(a, state)[0]
    dt = time.perf_counter() - t0
    v = units / dt
    cfg = workload_config(a, a.gpus)
    cfg["reference_arm_sample"] = what            # what one step of THIS arm ran: a bounded sample of the workload above
    cfg["reference_arm_frames_per_step"] = 1 if a.config in (2, 5) else None
    print(json.dumps({
        "impl": "reference", "metric": metric, "value": v, "unit": unit, "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": dt / a.steps * 1e3, "higher_is_better": True,
        "scaling": "weak" if a.config in (2, 5) else "strong",
        "vs_baseline": None, "dtype": "fp32", "data": "synthetic", "config": cfg,
        "cpu_baseline": {"value": v, "unit": unit, "cores": state["cores"], "kind": state["kind"],
                         "sample": f"{a.steps} x [{what}]", "note": CPU_NOTE},
        "e2e": {"value": v, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


def workload_config(a, n_gpus):
    w = "random-init synthetic (weights/*.onnx absent offline; B2F_SYNTHETIC_WEIGHTS=1)"
    if a.config in (2, 5):
        nets = "SCRFD-10G + ArcFace R50" if a.config == 2 else "SCRFD-2.5G + ArcFace R50 video loop"
        sharded = a.config == 2
        return {
            "workload": (f"configs[{a.config - 1}]: {nets}, {a.frames} synthetic {a.width}x{a.height} frames per GPU per step, "
                         f"max_num={a.faces} ({a.frames * a.faces} face slots), top-1 vs {a.gallery} x 512 gallery"),
            "frames_per_gpu": a.frames, "faces_per_frame": a.faces, "frame_hw": [a.height, a.width],
            "gallery_rows": a.gallery, "global_frames": a.frames * n_gpus,
            "l2_policy": f"inputs larger than L2: {a.frames * a.height * a.width * 3 / 1e6:.0f} MB of frames per step",
            "parallelism": ("single GPU" if n_gpus == 1 else
                            (f"frames split {a.frames}/rank; gallery rows split {a.gallery // n_gpus}/rank; NCCL all_gather of "
                             f"embeddings + one MAX all_reduce of packed top-1 keys, {a.tail} with the next step's nets")
                            if sharded else f"frames dealt {a.frames}/rank per step; {a.gallery}-row target gallery replicated; no collective"),
            "weights": w}
    if a.config == 3:
        return {"workload": (f"configs[2]: ArcFace R50 embedding of {a.crops} synthetic aligned 112x112 crops + cosine top-1 vs "
                             f"{a.gallery} x 512 gallery"),
                "crops": a.crops, "crops_per_gpu": -(-a.crops // n_gpus), "gallery_rows": a.gallery,
                "l2_policy": f"inputs larger than L2: {a.crops // n_gpus * 37632 / 1e6:.0f} MB of crops per GPU per step",
                "parallelism": ("single GPU" if n_gpus == 1 else
                                f"crops split {-(-a.crops // n_gpus)}/rank; gallery rows split {a.gallery // n_gpus}/rank; NCCL "
                                "all_gather of embeddings + one MAX all_reduce of packed top-1 keys"),
                "weights": w}
    return {"workload": (f"configs[3]: all-pairs cosine clustering of {a.rows} x 512 embeddings ({a.rows // 4} planted identities "
                         f"x 4), threshold 0.8, greedy one-hop leader merge in id order (duplicate.py:2726-2797)"),
            "rows": a.rows, "pairs": a.rows * (a.rows - 1) // 2,
            "l2_policy": f"inputs larger than L2: {a.rows * 512 * 2 / 1e6:.0f} MB fp16 + {a.rows * 512 * 4 / 1e6:.0f} MB fp32 rows",
            "parallelism": ("single GPU" if n_gpus == 1 else
                            f"upper triangle block-partitioned into {n_gpus} equal-area row ranges, one per rank (embeddings replicated); "
                            "NCCL all_gather of pair lists; resolve replicated"),
            "weights": "n/a"}


# ---------------------------------------------------------------------------------------------
# configs 2 and 5: frames -> detect -> align -> embed -> match
# ---------------------------------------------------------------------------------------------
def run_pipeline(a):
    job = Job(a)
    torch, dist, dev, world, rank = job.torch, job.dist, job.dev, job.world, job.rank
    from models import SCRFD, ArcFace
    from scrfd_arcface_facerecognition_b200 import _lib
    from scrfd_arcface_facerecognition_b200.gallery import Gallery, shard_range
    from scrfd_arcface_facerecognition_b200.pipeline import FacePipeline
    pk = peaks()
    metric, unit = METRICS[a.config]
    sharded = a.config == 2 and world > 1          # config 5 replicates its small target gallery: no collective

    B, F, H, W = a.frames, a.faces, a.height, a.width
    det = SCRFD(a.det)
    rec = ArcFace(a.rec)
    gal = Gallery(rank=rank, world_size=world if a.config == 2 else 1)
    g0, g1 = shard_range(a.gallery, rank, world) if a.config == 2 else (0, a.gallery)
    gen = torch.Generator(device=dev).manual_seed(2)
    rows = torch.randn((g1 - g0, 512), generator=gen, device=dev)      # each rank draws its own shard
    gal.set_shard(rows, g0)
    del rows
    pipe = FacePipeline(det, rec, None if sharded else gal, max_num=F, similarity_thresh=0.4)

    # synthetic frames: two alternating batches, resident in HBM (398 MB each at 64 x 1080p)
    rng = np.random.default_rng(1000 + rank)
    host = [torch.from_numpy(rng.integers(0, 256, (B, H, W, 3), dtype=np.uint8)).pin_memory() for _ in range(2)]
    resident = [h.to(dev) for h in host]

    # two input slots, each with its own captured graph: the resident run alternates between two batches that live in
    # HBM, the e2e run uploads batch i+1 straight into the other slot while batch i runs (no staging copy in either)
    use_graph = not a.no_graph
    if use_graph:
        caps = [pipe.capture(B, H, W, slot) for slot in range(2)]
        statics, graphs = [c[0] for c in caps], [c[2] for c in caps]
        outs, kernels_per_step = caps[0][1], caps[0][3]
    else:
        statics = [torch.empty((B, H, W, 3), dtype=torch.uint8, device=dev) for _ in range(2)]
        graphs = None
        before = _lib.launch_count()
        outs = pipe.process(statics[0])
        kernels_per_step = _lib.launch_count() - before
    for slot in range(2):
        statics[slot].copy_(resident[slot])
    static = statics[0]

    # ---- multi-GPU tail (config 2): every rank needs every query against its gallery shard.  The embeddings of step i
    # are copied out (2 MB) and their exchange + match + key all-reduce run on a side stream while the main stream
    # already runs step i+1's nets: the collectives keep the ranks loosely coupled instead of in lockstep per step.
    side = torch.cuda.Stream(device=dev) if sharded else None
    emb_copy = [torch.empty((B * F, 512), dtype=torch.float32, device=dev) for _ in range(2)] if sharded else None
    tail_done = [torch.cuda.Event() for _ in range(2)] if sharded else None
    tail_out = [None, None]
    tail_launches = 0

    def run_tail(i):
        nonlocal tail_launches
        cur = torch.cuda.current_stream()
        slot = i & 1
        cur.wait_event(tail_done[slot])                       # the tail that last used this slot has finished
        emb_copy[slot].copy_(outs["emb"], non_blocking=True)
        if a.tail == "inline":
            tail_out[slot] = gal.match_sharded_queries(emb_copy[slot], 0.4, strict=True)
            tail_done[slot].record(cur)
            return tail_out[slot]
        ready = torch.cuda.Event()
        ready.record(cur)
        with torch.cuda.stream(side):
            side.wait_event(ready)
            tail_out[slot] = gal.match_sharded_queries(emb_copy[slot], 0.4, strict=True)
            tail_done[slot].record(side)
        return tail_out[slot]

    def step_resident(i):
        nonlocal outs
        if use_graph:                                           # inputs already in HBM: batch i & 1, 398 MB each (> L2)
            graphs[i & 1].replay()
        else:
            outs = pipe.process(statics[i & 1])
        if sharded:
            return run_tail(i)
        return outs["match_score"], outs["match_idx"]

    def drain():
        if sharded and a.tail == "overlap":
            torch.cuda.current_stream().wait_stream(side)

    if sharded:
        for e in tail_done:
            e.record(torch.cuda.current_stream())

    # enrolment (reference build_targets, main.py:78-105): the faces of both synthetic batches are planted at
    # known rows of this rank's gallery shard, so top-1 has a ground truth among the random rows
    n_slots = B * F
    assert g1 - g0 >= 2 * n_slots, "gallery shard smaller than the faces planted in it"
    plant_rows = torch.from_numpy(np.random.default_rng(7 + rank).permutation(g1 - g0)[:2 * n_slots]).to(dev)
    for bi in range(2):
        step_resident(bi)
        drain()
        torch.cuda.synchronize()
        gal.replace_rows(plant_rows[bi * n_slots:(bi + 1) * n_slots], outs["emb"].clone())
    # validation step: how many face slots are real detections (bench counts only those), and is top-1 right
    s, idx = step_resident(0)
    drain()
    torch.cuda.synchronize()
    counts = outs["counts"].cpu().numpy()
    faces_per_step = int(counts[:, 0].sum())
    overflow = int((counts[:, 3] != 0).sum())
    matched = int((idx >= 0).sum().item())
    # ground truth: exact fp32 cosine against the planted rows (everything else in the gallery is ~0.2 away);
    # bit-identical embeddings (e.g. all-border crops from detections in the letterbox padding) tie, and the
    # reference's strict '>' scan keeps the lowest index (main.py:140)
    qn = torch.nn.functional.normalize(outs["emb"].float(), dim=1)
    planted = gal.f32[plant_rows]
    planted_ids = plant_rows + g0
    if sharded:                                       # identical embeddings can be planted on another rank
        pl = [torch.empty_like(planted) for _ in range(world)]
        pi = [torch.empty_like(planted_ids) for _ in range(world)]
        dist.all_gather(pl, planted.contiguous())
        dist.all_gather(pi, planted_ids.contiguous())
        planted, planted_ids = torch.cat(pl), torch.cat(pi)
    sims = qn @ planted.T                             # validation only (library matmul, outside every timed region)
    best = sims.max(dim=1, keepdim=True).values
    big = torch.iinfo(torch.int64).max
    expect = torch.where(sims >= best - 1e-6, planted_ids[None, :], torch.full_like(planted_ids[None, :], big)).min(dim=1).values
    top1_correct = int((idx.reshape(-1) == expect).sum().item())
    del sims

    for i in range(a.warmup):
        step_resident(i)
    drain()
    sampler = ClockSampler(job.local)
    if rank == 0:
        sampler.start()
    launches0 = _lib.launch_count()
    ms_total = job.timed(step_resident, a.steps, finish=drain)
    eager_launches = _lib.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    ms_step = ms_total / a.steps
    value = faces_per_step * world / (ms_step / 1e3)

    # ---- e2e: pinned host frames in, results out; batch i+1 is uploaded on a copy stream into the other input slot
    #      while batch i runs -------------------------------------------------------------------------------------------
    copy_stream = torch.cuda.Stream(device=dev)
    res_host = {k: torch.empty((B, F), dtype=dt).pin_memory() for k, dt in (("score", torch.float32), ("idx", torch.int64))}
    det_host = torch.empty((B, F, 5), dtype=torch.float32).pin_memory()
    ready = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]
    h2d = B * H * W * 3
    d2h = B * F * (4 + 8) + B * F * 5 * 4

    def upload(i):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[i & 1])            # the step that last read this slot has finished
            statics[i & 1].copy_(host[i & 1], non_blocking=True)
            ready[i & 1].record(copy_stream)

    def step_e2e(i):
        nonlocal outs
        cur = torch.cuda.current_stream()
        if i == 0:
            upload(0)
        upload(i + 1)                                           # next step's frames ride under this step's compute
        cur.wait_event(ready[i & 1])
        if use_graph:
            graphs[i & 1].replay()
        else:
            outs = pipe.process(statics[i & 1])
        consumed[i & 1].record(cur)
        det_host.copy_(outs["det"], non_blocking=True)
        if sharded:
            sc, ix = run_tail(i)
            on = side if a.tail == "overlap" else cur          # the result copies follow the tail on its stream
            with torch.cuda.stream(on):
                res_host["score"].copy_(sc.reshape(B, F), non_blocking=True)
                res_host["idx"].copy_(ix.reshape(B, F), non_blocking=True)
                tail_done[i & 1].record(on)
        else:
            res_host["score"].copy_(outs["match_score"].reshape(B, F), non_blocking=True)
            res_host["idx"].copy_(outs["match_idx"].reshape(B, F), non_blocking=True)

    for e in consumed:
        e.record(torch.cuda.current_stream())
    for i in range(max(2, a.warmup)):
        step_e2e(i)
    drain()
    torch.cuda.synchronize()
    for e in consumed:
        e.record(torch.cuda.current_stream())
    ms_e2e = job.timed(step_e2e, a.steps, finish=drain) / a.steps
    e2e_value = faces_per_step * world / (ms_e2e / 1e3)
    # the upload alone (pinned host -> HBM, nothing else running): when it is as long as a step, e2e is bound by the host link
    h2d_alone_ms = job.max_over_ranks(event_ms(torch, lambda: statics[0].copy_(host[0], non_blocking=True), 3))

    # ---- the same resident step held for about two seconds: a B200 running this path settles under its power cap
    # (sw_power_cap, SM clocks ~1.45 GHz); reported beside `value`, which is the K steps the caller asked for
    sus_steps = max(100, 2 * a.steps)
    ms_sus = job.timed(step_resident, sus_steps, finish=drain) / sus_steps
    sustained = {"steps": sus_steps, "value": faces_per_step * world / (ms_sus / 1e3), "ms_per_step": ms_sus,
                 "what": "device-resident step repeated after the timed runs (power-capped steady state)"}

    # ---- rooflines: the conv launches (tensor), and the memory-bound kernels around them (HBM), per-launch CUDA
    #      events in eager passes on rank 0 ------------------------------------------------------------------------------
    roofline, mem = None, None
    if rank == 0:
        eng_d = det._engine_for(det.input_size[1], det.input_size[0])
        st8 = rec._engine.stem8(B * F) if getattr(rec, "stem8", False) else None
        det_start = 2 if (det.fuse_stem and det.fuse_conv1 and eng_d.stem_fused(B) is not None) else \
            (1 if eng_d.patch_buffer(B) is not None else 0)        # the first layer may live in the preprocess kernel
        nets = [("SCRFD-10G" if a.config == 2 else "SCRFD-2.5G", eng_d, B, None, det_start),
                ("R50", rec._engine, B * F, st8[1] if st8 is not None else None, 2 if st8 is not None else 0)]
        roofline, pool_spans = conv_roofline(a, nets, pk)
        roofline["share_of_step"] = roofline.pop("conv_ms_per_step") / ms_step if ms_step else None
        local_pipe = FacePipeline(det, rec, gal if not sharded else None, max_num=F, similarity_thresh=0.4)
        spans = []
        for rep in range(3):                                    # third repetition kept
            _lib.profile_begin()
            local_pipe.process(static)
            spans = _lib.profile_end()
        # pooling ops are kernels of the nets' launch lists: the three largest are reported with the stage kernels
        pool_spans.sort(key=lambda r: -r[2])
        mem = memory_rooflines(spans + pool_spans[:3], pk)
        for m in mem:
            if m["kernel"].startswith("decode_nms"):
                m["note"] = "one CTA per frame: latency-bound (sort + sequential NMS), not bandwidth-bound"

    # ---- the match stage alone (BASELINE metric, second half: "match TFLOP/s") -- exactly what a step runs: at N > 1 every
    #      rank matches the queries of ALL ranks (world x B*F) against its own shard
    q_match = outs["emb"].clone()
    if sharded:
        q_match = q_match.repeat(world, 1)
    for _ in range(3):
        gal.match_local(q_match, 1, 0.4, strict=True)
    torch.cuda.synchronize()
    match_ms = job.max_over_ranks(event_ms(torch, lambda: gal.match_local(q_match, 1, 0.4, strict=True), 10))
    match_flops = 2.0 * q_match.shape[0] * len(gal) * 512
    match_stage = {"ms": match_ms, "tflops": match_flops / match_ms / 1e9, "queries": int(q_match.shape[0]),
                   "gallery_rows": len(gal), "flops": match_flops,
                   "frac_of_bf16_peak": match_flops / match_ms / 1e9 / pk["tensor"],
                   "what": "l2norm + tcgen05 cosine top-k GEMM (fp16 operands) + exact fp32 re-score, per rank "
                           "(queries of all ranks against this rank's gallery shard, as in the step)"}

    out = {
        "metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "fp16" if rec._engine.dtype == 0 else "bf16", "data": "synthetic",
        "config": workload_config(a, world), "roofline": roofline, "memory_kernels": mem,
        "e2e": {"value": e2e_value, "unit": unit, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": ms_e2e, "h2d_alone_ms": h2d_alone_ms, "h2d_alone_gbs": h2d / h2d_alone_ms / 1e6,
                "note": "batch i+1 is uploaded under batch i's compute; e2e approaches `value` when h2d_alone_ms < ms_per_step"},
        "gpu_launches": int(kernels_per_step * a.steps + eager_launches) if use_graph else int(eager_launches),
        "kernels_per_step": int(kernels_per_step + (eager_launches // max(a.steps, 1) if use_graph else 0)), "clocks": clocks,
        "faces_per_step_per_gpu": faces_per_step, "matched_faces": matched, "top1_correct": top1_correct,
        "decode_overflow_frames": overflow,
        "match_tflops": match_stage["tflops"] * world, "match": match_stage,
        "cuda_graph": use_graph, "sustained": sustained,
    }
    if a.config == 5:
        fps = B * world / (ms_step / 1e3)
        out["video"] = {"frames_per_s": fps, "realtime_24fps_streams": fps / 24.0,
                        "e2e_frames_per_s": B * world / (ms_e2e / 1e3), "faces_per_frame": faces_per_step / B}
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline(a)
    if rank == 0:
        print(json.dumps(out))
    job.finish()


# ---------------------------------------------------------------------------------------------
# config 3: embed 100k aligned crops + top-1 against the 1M gallery
# ---------------------------------------------------------------------------------------------
def run_embed_match(a):
    job = Job(a)
    torch, dist, dev, world, rank = job.torch, job.dist, job.dev, job.world, job.rank
    from models import ArcFace
    from scrfd_arcface_facerecognition_b200 import _lib
    from scrfd_arcface_facerecognition_b200.gallery import Gallery, shard_range
    pk = peaks()
    metric, unit = METRICS[3]
    per = -(-a.crops // world)                       # every rank embeds the same number of crops (the tail is padding-free
    c0, c1 = rank * per, min((rank + 1) * per, a.crops)  # when crops % world == 0, as at 100k / 8)
    assert c1 - c0 == per, "--crops must be a multiple of the number of GPUs"
    chunk = 1024
    rec = ArcFace(a.rec)
    gal = Gallery(rank=rank, world_size=world)
    g0, g1 = shard_range(a.gallery, rank, world)
    gen = torch.Generator(device=dev).manual_seed(2)
    rows = torch.randn((g1 - g0, 512), generator=gen, device=dev)
    gal.set_shard(rows, g0)
    del rows
    gen = torch.Generator(device=dev).manual_seed(100 + rank)
    crops = torch.randint(0, 256, (per, 112, 112, 3), generator=gen, device=dev, dtype=torch.uint8)   # 37.6 KB each
    emb = torch.empty((per, 512), dtype=torch.float32, device=dev)

    def embed_all(src):
        for lo in range(0, per, chunk):
            e = rec.embed_crops(src[lo:lo + chunk], copy=False)
            emb[lo:lo + e.shape[0]].copy_(e, non_blocking=True)

    def step(i):
        embed_all(crops)
        return gal.match_sharded_queries(emb, 0.4, strict=True)

    # plant every crop's embedding at a known row of this rank's shard: ground truth for top-1
    embed_all(crops)
    torch.cuda.synchronize()
    assert g1 - g0 >= per
    plant_rows = torch.from_numpy(np.random.default_rng(7 + rank).permutation(g1 - g0)[:per]).to(dev)
    gal.replace_rows(plant_rows, emb.clone())
    s, idx = step(0)
    torch.cuda.synchronize()
    # identical crops cannot occur (independent uniform bytes), so the planted row is the unique exact match
    top1_correct = int((idx.reshape(-1) == plant_rows + g0).sum().item())
    matched = int((idx >= 0).sum().item())

    for i in range(max(a.warmup - 1, 0)):
        step(i)
    sampler = ClockSampler(job.local)
    if rank == 0:
        sampler.start()
    launches0 = _lib.launch_count()
    ms_step = job.timed(step, a.steps) / a.steps
    launches = _lib.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    value = a.crops / (ms_step / 1e3)

    # stage split of one more step (device events, max over ranks)
    embed_ms = job.max_over_ranks(event_ms(torch, lambda: embed_all(crops)))
    match_ms = job.max_over_ranks(event_ms(torch, lambda: gal.match_sharded_queries(emb, 0.4, strict=True)))
    allq = emb.repeat(world, 1) if world > 1 else emb
    gemm_ms = job.max_over_ranks(event_ms(torch, lambda: gal.match_local(allq, 1, 0.4, strict=True), 3))
    match_flops = 2.0 * a.crops * a.gallery * 512                    # whole job: every query against every gallery row
    match = {"ms": match_ms, "tflops": match_flops / match_ms / 1e9, "queries": a.crops, "gallery_rows": a.gallery,
             "flops": match_flops, "per_rank_gemm_ms": gemm_ms,
             "per_rank_tflops": 2.0 * allq.shape[0] * len(gal) * 512 / gemm_ms / 1e9,
             "frac_of_bf16_peak": 2.0 * allq.shape[0] * len(gal) * 512 / gemm_ms / 1e9 / pk["tensor"],
             "what": "whole-job match (all_gather of queries + l2norm + tcgen05 cosine top-1 GEMM + exact fp32 re-score + key "
                     "all_reduce); per_rank_* is the local GEMM + re-score alone"}
    del allq

    # e2e: crops from pinned host memory in 1024-crop chunks on a copy stream, two staging buffers; results to the host
    host = torch.empty((per, 112, 112, 3), dtype=torch.uint8).pin_memory()
    host.copy_(crops)
    copy_stream = torch.cuda.Stream(device=dev)
    stage = [torch.empty((chunk, 112, 112, 3), dtype=torch.uint8, device=dev) for _ in range(2)]
    ready = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]
    res_s, res_i = torch.empty(per, dtype=torch.float32).pin_memory(), torch.empty(per, dtype=torch.int64).pin_memory()
    n_chunks = -(-per // chunk)

    def upload(c):
        lo = c * chunk
        n = min(chunk, per - lo)
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[c & 1])
            stage[c & 1][:n].copy_(host[lo:lo + n], non_blocking=True)
            ready[c & 1].record(copy_stream)

    def step_e2e(i):
        cur = torch.cuda.current_stream()
        upload(0)
        for c in range(n_chunks):
            if c + 1 < n_chunks:
                upload(c + 1)
            lo = c * chunk
            n = min(chunk, per - lo)
            cur.wait_event(ready[c & 1])
            e = rec.embed_crops(stage[c & 1][:n], copy=False)
            consumed[c & 1].record(cur)
            emb[lo:lo + n].copy_(e, non_blocking=True)
        sc, ix = gal.match_sharded_queries(emb, 0.4, strict=True)
        res_s.copy_(sc.reshape(-1), non_blocking=True)
        res_i.copy_(ix.reshape(-1), non_blocking=True)

    for e in consumed:
        e.record(torch.cuda.current_stream())
    step_e2e(0)
    torch.cuda.synchronize()
    ms_e2e = job.timed(step_e2e, a.steps) / a.steps
    e2e_ok = int((res_i == (plant_rows + g0).cpu()).sum().item())

    roofline, mem = None, None
    if rank == 0:
        st8 = rec._engine.stem8(chunk) if getattr(rec, "stem8", False) else None
        roofline, _ = conv_roofline(a, [("R50", rec._engine, chunk, st8[1] if st8 is not None else None, 2 if st8 is not None else 0)], pk)
        conv_ms = roofline.pop("conv_ms_per_step") * (per / chunk)
        roofline["launches_per_step"] = int(roofline["launches_per_step"] * n_chunks)
        roofline["flops_per_step"] = roofline["flops_per_step"] * per / chunk
        roofline["share_of_step"] = conv_ms / ms_step
        _lib.profile_begin()
        rec.embed_crops(crops[:chunk], copy=False)
        mem = memory_rooflines(_lib.profile_end(), pk)
    out = {
        "metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "fp16" if rec._engine.dtype == 0 else "bf16", "data": "synthetic", "config": workload_config(a, world),
        "roofline": roofline, "memory_kernels": mem,
        "e2e": {"value": a.crops / (ms_e2e / 1e3), "unit": unit, "h2d_bytes_per_step": per * 37632,
                "d2h_bytes_per_step": per * 12, "ms_per_step": ms_e2e, "top1_correct_per_gpu": e2e_ok},
        "gpu_launches": int(launches), "kernels_per_step": int(launches // max(a.steps, 1)), "clocks": clocks,
        "crops_per_gpu": per, "matched_faces_per_gpu": matched, "top1_correct_per_gpu": top1_correct,
        "stages": {"embed_ms": embed_ms, "match_ms": match_ms, "embed_faces_per_s": a.crops / (embed_ms / 1e3)},
        "match_tflops": match["tflops"], "match": match, "cuda_graph": False,
    }
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline(a)
    if rank == 0:
        print(json.dumps(out))
    job.finish()


# ---------------------------------------------------------------------------------------------
# config 4: all-pairs clustering of 200k embeddings
# ---------------------------------------------------------------------------------------------
def run_cluster(a):
    job = Job(a)
    torch, dist, dev, world, rank = job.torch, job.dist, job.dev, job.world, job.rank
    from scrfd_arcface_facerecognition_b200 import _lib
    from scrfd_arcface_facerecognition_b200.gallery import Gallery
    pk = peaks()
    metric, unit = METRICS[4]
    n = a.rows
    centres = n // 4
    # N/4 centres x 4 noisy members (cos ~ 0.89 inside an identity, < 0.3 across), shuffled; same on every rank
    gen = torch.Generator(device=dev).manual_seed(3)
    c = torch.nn.functional.normalize(torch.randn((centres, 512), generator=gen, device=dev), dim=1)
    x = c.repeat_interleave(4, dim=0)
    x = x + 0.35 * torch.nn.functional.normalize(torch.randn((centres * 4, 512), generator=gen, device=dev), dim=1)
    perm = torch.randperm(centres * 4, generator=gen, device=dev)
    x = x[perm].contiguous()
    truth = (torch.arange(centres * 4, device=dev) // 4)[perm]           # planted identity of every row
    n = x.shape[0]
    pairs_total = n * (n - 1) // 2
    G = Gallery(rank=rank, world_size=world)
    G.set_shard(x, 0)                                                    # embeddings replicated (205 MB fp16 + 410 MB fp32)

    def step(i):
        return G.merge_duplicates(0.8)

    leader = step(0)
    # ground truth: rows of one planted identity share one leader, the lowest row index among them
    lt = torch.from_numpy(leader.astype(np.int64)).to(dev)
    first_row = torch.full((centres,), n, dtype=torch.int64, device=dev).scatter_reduce(0, truth, torch.arange(n, device=dev), "amin")
    clusters_ok = bool((lt == first_row[truth]).all().item())
    n_clusters = int((lt == torch.arange(n, device=dev)).sum().item())
    if world > 1:                                                        # every rank resolved the same labels
        ref = lt.clone()
        dist.broadcast(ref, 0)
        clusters_ok = clusters_ok and bool((ref == lt).all().item())

    for i in range(max(a.warmup - 1, 0)):
        step(i)
    sampler = ClockSampler(job.local)
    if rank == 0:
        sampler.start()
    launches0 = _lib.launch_count()
    ms_step = job.timed(step, a.steps) / a.steps
    launches = _lib.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    value = pairs_total / (ms_step / 1e3)

    # the pair GEMM alone (this rank's row blocks), device events, max over ranks
    from scrfd_arcface_facerecognition_b200.gallery import triangle_range
    mine = [triangle_range(n, rank, world)] if world > 1 else [(0, n)]

    def pairs_only():
        for b, e in mine:
            G.duplicate_pairs(0.8, b, e)
    pairs_only()
    gemm_ms = job.max_over_ranks(event_ms(torch, pairs_only))
    pair_flops = 2.0 * pairs_total * 512
    # e2e: embeddings from pinned host memory every step (H2D + normalise + cluster), labels back on the host
    host = x.cpu().pin_memory()
    stage = torch.empty_like(x)

    def step_e2e(i):
        stage.copy_(host, non_blocking=True)
        G.set_shard(stage, 0)
        return G.merge_duplicates(0.8)                                   # returns host numpy labels (D2H inside)
    step_e2e(0)
    ms_e2e = job.timed(step_e2e, a.steps) / a.steps

    roofline = {"bound": "tensor", "kernel": "pair-threshold GEMM (tcgen05, upper-triangle tiles only) + exact fp32 re-check",
                "achieved": pair_flops / (gemm_ms / 1e3) / 1e12, "peak": pk["tensor"], "unit": "TFLOP/s",
                "frac": pair_flops / (gemm_ms / 1e3) / 1e12 / pk["tensor"], "traffic": None,
                "flops_per_launch": pair_flops / max(len(mine) * world, 1), "ms_per_launch": gemm_ms / max(len(mine), 1),
                "peak_source": pk["tensor_src"], "launches_per_step": len(mine), "flops_per_step": pair_flops,
                "share_of_step": gemm_ms / ms_step,
                "note": "algorithmic FLOPs = 1024 per unordered pair (SURVEY 8d); whole job, all ranks; time = slowest rank's blocks incl. "
                        "the pair-list sort"}
    out = {
        "metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "fp16", "data": "synthetic", "config": workload_config(a, world), "roofline": roofline,
        "e2e": {"value": pairs_total / (ms_e2e / 1e3), "unit": unit, "h2d_bytes_per_step": n * 512 * 4, "d2h_bytes_per_step": n * 4,
                "ms_per_step": ms_e2e},
        "gpu_launches": int(launches), "kernels_per_step": int(launches // max(a.steps, 1)), "clocks": clocks,
        "rows": n, "clusters_found": n_clusters, "clusters_planted": centres, "labels_correct": clusters_ok,
        "pairs_tflops": pair_flops / (ms_step / 1e3) / 1e12, "rows_per_s": n / (ms_step / 1e3), "cuda_graph": False,
    }
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline(a)
    if rank == 0:
        print(json.dumps(out))
    job.finish()


def main():
    a = parse_args()
    if a.impl == "reference":
        run_reference_arm(a)
    elif a.config in (2, 5):
        run_pipeline(a)
    elif a.config == 3:
        run_embed_match(a)
    else:
        run_cluster(a)


if __name__ == "__main__":
    main()
