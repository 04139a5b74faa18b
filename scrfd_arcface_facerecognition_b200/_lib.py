"""Build and bind libb2f.so (the C-ABI of include/b2f.h) through ctypes.

There is no CPU fallback: `lib()` raises if the shared library is missing or cannot be loaded,
and every compute wrapper raises on a non-zero status with the library's own error text.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import threading
from typing import List

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
SO_PATH = os.environ.get("B2F_SO") or os.path.join(CSRC, "libb2f.so")     # B2F_SO: an experimental build to A/B against
SOURCES = ["core.cu", "postproc.cu", "aux_ops.cu", "umma_conv.cu", "conv_tile.cu", "match_pair.cu", "overlay.cu"]
HEADERS = ["b2f_common.cuh", "umma_shared.cuh", os.path.join("..", "..", "include", "b2f.h")]
NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC", "-shared"]

_lock = threading.Lock()
_lib = None


class B2FError(RuntimeError):
    pass


def _stale() -> bool:
    if not os.path.exists(SO_PATH):
        return True
    so_m = os.path.getmtime(SO_PATH)
    for f in SOURCES + HEADERS:
        p = os.path.join(CSRC, f)
        if os.path.exists(p) and os.path.getmtime(p) > so_m:
            return True
    return False


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/*.cu for sm_100a into csrc/libb2f.so (in-tree, travels with the repo snapshot).
    Ranks of one torchrun job may all find the library stale: an exclusive file lock serialises them (the ones that
    waited see a fresh library and return), and each compile goes to its own temp file before the atomic rename."""
    if not force and not _stale():
        return SO_PATH
    import fcntl
    with open(os.path.join(CSRC, ".build.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not _stale():
                return SO_PATH
            nvcc = os.environ.get("NVCC", "nvcc")
            # one object per source, compiled in parallel and only when the source (or a header) is newer than it
            from concurrent.futures import ThreadPoolExecutor
            objdir = os.path.join(CSRC, "build")
            os.makedirs(objdir, exist_ok=True)
            hdr_m = max(os.path.getmtime(os.path.join(CSRC, h)) for h in HEADERS if os.path.exists(os.path.join(CSRC, h)))

            def compile_one(src):
                obj = os.path.join(objdir, src.replace(".cu", ".o"))
                src_m = max(os.path.getmtime(os.path.join(CSRC, src)), hdr_m)
                if not force and os.path.exists(obj) and os.path.getmtime(obj) >= src_m:
                    return obj, None
                cmd = [nvcc] + [f for f in NVCC_FLAGS if f != "-shared"] + ["-c", "-o", obj, src]
                res = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
                return obj, (None if res.returncode == 0 else f"({' '.join(cmd)}):\n{res.stdout}\n{res.stderr}")
            with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
                results = list(ex.map(compile_one, SOURCES))
            errors = [e for _, e in results if e]
            if errors:
                raise B2FError("nvcc failed " + "\n".join(errors))
            tmp = f"{SO_PATH}.{os.getpid()}.tmp"
            cmd = [nvcc] + NVCC_FLAGS + ["-o", tmp] + [o for o, _ in results]
            res = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
            if res.returncode != 0:
                if os.path.exists(tmp):
                    os.remove(tmp)
                raise B2FError(f"nvcc link failed ({' '.join(cmd)}):\n{res.stdout}\n{res.stderr}")
            os.replace(tmp, SO_PATH)
            if verbose:
                print(res.stdout, res.stderr)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return SO_PATH


class DetLevels(C.Structure):
    _fields_ = [("score", C.c_void_p * 3), ("bbox", C.c_void_p * 3), ("kps", C.c_void_p * 3),
                ("score_ps", C.c_int * 3), ("bbox_ps", C.c_int * 3), ("kps_ps", C.c_int * 3)]


class DrawCmd(C.Structure):
    _fields_ = [("kind", C.c_int), ("x0", C.c_int), ("y0", C.c_int), ("x1", C.c_int), ("y1", C.c_int),
                ("bgr", C.c_uint), ("mask_off", C.c_int), ("reserved", C.c_int)]


class ConvDesc(C.Structure):
    _fields_ = [("n", C.c_int), ("h", C.c_int), ("w", C.c_int), ("cin_p", C.c_int),
                ("ho", C.c_int), ("wo", C.c_int), ("cout_p", C.c_int),
                ("kh", C.c_int), ("kw", C.c_int), ("stride", C.c_int), ("pad", C.c_int),
                ("dtype", C.c_int), ("out_dtype", C.c_int), ("act", C.c_int), ("bias_classes", C.c_int),
                ("res_mode", C.c_int), ("res_h", C.c_int), ("res_w", C.c_int), ("force_kchunk", C.c_int),
                ("sig_hi", C.c_int),
                ("in_", C.c_void_p), ("weight", C.c_void_p), ("bias", C.c_void_p), ("slope", C.c_void_p),
                ("residual", C.c_void_p), ("out", C.c_void_p),
                ("sc_in", C.c_void_p), ("sc_weight", C.c_void_p),
                ("sc_cin_p", C.c_int), ("sc_stride", C.c_int), ("sc_h", C.c_int), ("sc_w", C.c_int), ("pool", C.c_int),
                ("splitk_ws", C.c_void_p), ("splitk_ws_bytes", C.c_longlong)]


_vp, _i, _f, _ll = C.c_void_p, C.c_int, C.c_float, C.c_longlong

# name -> argtypes; every entry returns int status unless listed in _RESTYPES
SIGNATURES = {
    "b2f_version": [],
    "b2f_last_error": [],
    "b2f_launch_count": [],
    "b2f_set_tuning": [_i, _i],
    "b2f_letterbox_u8": [_vp, _i, _i, _i, _i, _i, _i, _i, _vp, _vp],
    "b2f_preprocess": [_vp, _i, _i, _i, _i, _i, _i, _i, _f, _f, _vp, _i, _i, _vp],
    "b2f_preprocess_patches": [_vp, _i, _i, _i, _i, _i, _i, _i, _i, _f, _f, _vp, _i, _vp],
    "b2f_preprocess_conv1": [_vp, _i, _i, _i, _i, _i, _i, _i, _f, _f, _vp, _vp, _i, _i, _vp, _i, _vp],
    "b2f_blob_nchw_f32": [_vp, _i, _i, _i, _f, _f, _vp, _vp],
    "b2f_decode_nms": [C.POINTER(DetLevels), _i, _i, _i, _vp, _vp, _f, _f, _i, _i, _i, _i, _vp, _vp, _vp, _vp,
                       _vp, _ll, _vp],
    "b2f_decode_nms_workspace": [_i, _i],
    "b2f_nms": [_vp, _i, _f, _vp, _vp, _vp, _ll, _vp],
    "b2f_distance2bbox": [_vp, _vp, _i, _vp, _vp],
    "b2f_distance2kps": [_vp, _vp, _i, _i, _vp, _vp],
    "b2f_estimate_norm": [_vp, _i, _i, _vp, _vp],
    "b2f_warp_affine_u8": [_vp, _i, _i, _vp, _vp, _i, _i, _vp, _vp],
    "b2f_norm_crop": [_vp, _i, _i, _vp, _vp, _i, _i, _f, _f, _vp, _i, _i, _vp, _vp, _vp],
    "b2f_norm_crop_patches": [_vp, _i, _i, _vp, _vp, _i, _i, _f, _f, _vp, _i, _vp],
    "b2f_norm_crop_image8": [_vp, _i, _i, _vp, _vp, _i, _i, _f, _f, _vp, _i, _vp],
    "b2f_conv2d": [C.POINTER(ConvDesc), _vp],
    "b2f_stem_conv3x3": [_vp, _i, _i, _i, _i, _i, _vp, _vp, _vp, _i, _i, _i, _vp, _vp],
    "b2f_im2col3x3": [_vp, _i, _i, _i, _i, _i, _i, _i, _vp, _vp],
    "b2f_dwconv": [_vp, _i, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _i, _i, _vp, _vp],
    "b2f_pool": [_vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _vp],
    "b2f_eltwise": [_vp, _vp, _ll, _i, _vp, _vp, _vp, _i, _i, _vp, _vp],
    "b2f_l2norm_rows": [_vp, _ll, _i, _vp, _vp, _i, _vp, _vp],
    "b2f_cosine_pairs": [_vp, _vp, _i, _i, _vp, _vp],
    "b2f_match_partial": [_vp, _i, _vp, _ll, _i, _i, _vp, _vp, _i, _i, _vp, _vp, _vp],
    "b2f_match_partial_keep": [_vp, _i, _vp, _ll, _i, _i, _vp, _vp, _i, _i, _i, _vp, _vp, _vp],
    "b2f_match_partial_causal": [_vp, _i, _vp, _ll, _i, _i, _i, _i, _i, _ll, _vp, _vp, _vp],
    "b2f_rows_dot": [_vp, _ll, _i, _vp, _vp, _vp],
    "b2f_match_splits": [_ll, _i],
    "b2f_match_plan": [_i, _ll],
    "b2f_match_merge": [_vp, _vp, _i, _i, _vp, _vp, _i, _i, _f, _i, _ll, _vp, _vp, _vp],
    "b2f_topk_pack_keys": [_vp, _vp, _ll, _vp, _vp],
    "b2f_topk_unpack_keys": [_vp, _ll, _vp, _vp, _vp],
    "b2f_pairs_threshold": [_vp, _i, _i, _i, _i, _i, _f, _vp, _vp, _ll, _vp, _vp],
    "b2f_cluster_resolve": [_vp, _ll, _i, _vp, _vp],
    "b2f_draw_overlay": [_vp, _i, _i, _i, _vp, _vp, _vp, _vp, _vp],
    "b2f_debug_tma_probe": [_vp, _vp, _vp, _vp, _i, _vp, _vp, _i, _vp],
}
_RESTYPES = {"b2f_last_error": C.c_char_p, "b2f_launch_count": _ll, "b2f_decode_nms_workspace": _ll}
_NO_STATUS = {"b2f_version", "b2f_last_error", "b2f_launch_count", "b2f_decode_nms_workspace", "b2f_match_splits",
              "b2f_match_plan"}


def declared_symbols() -> List[str]:
    """extern-C names declared in include/b2f.h (used by the CPU-side export test)."""
    import re
    with open(os.path.join(_HERE, "..", "include", "b2f.h")) as f:
        text = f.read()
    return sorted(set(re.findall(r"\b(b2f_[a-z0-9_]+)\s*\(", text)))


def lib() -> C.CDLL:
    """Load libb2f.so, building it first if sources are newer.  Raises -- never falls back."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if _stale():
            # a library older than its sources is never loaded silently: rebuild, and if that fails, raise.  The one
            # exception is a box without a compiler that received a prebuilt library next to freshly copied sources
            # (file times are not preserved by the snapshot): B2F_PREBUILT=1, or no nvcc on PATH at all.
            import shutil
            have_nvcc = shutil.which(os.environ.get("NVCC", "nvcc")) is not None
            if os.path.exists(SO_PATH) and (os.environ.get("B2F_PREBUILT") == "1" or not have_nvcc):
                pass
            else:
                try:
                    build()
                except Exception as e:
                    raise B2FError(f"libb2f.so is stale or missing and could not be rebuilt: {e}") from e
        try:
            handle = C.CDLL(SO_PATH)
        except OSError as e:
            raise B2FError(f"cannot load {SO_PATH}: {e}") from e
        for name, args in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.argtypes = args
            fn.restype = _RESTYPES.get(name, C.c_int)
        # B2F_TUNE="key=value,..." applies b2f_set_tuning knobs at load (A/B runs of bench.py and the tests)
        for item in filter(None, os.environ.get("B2F_TUNE", "").split(",")):
            key, _, val = item.partition("=")
            if handle.b2f_set_tuning(int(key), int(val)) != 0:
                raise B2FError(f"B2F_TUNE: bad knob {item!r}")
        _lib = handle
    return _lib


def check(status: int, what: str = "") -> None:
    if status != 0:
        msg = lib().b2f_last_error()
        raise B2FError(f"{what or 'b2f call'} failed with status {status}: {msg.decode() if msg else ''}")


def call(name: str, *args):
    """Invoke a status-returning entry point and raise B2FError on failure."""
    fn = getattr(lib(), name)
    rc = fn(*args)
    if name not in _NO_STATUS:
        check(rc, name)
    return rc


# ---- per-call timing of the memory-bound kernels (bench.py's HBM rooflines) ---------------------------------------
_spans = None


class span:
    """`with _lib.span(name, algorithmic_bytes):` around one C-ABI call.  Does nothing unless `profile_begin()` armed it;
    then it brackets the call with CUDA events on the current stream (never used under graph capture)."""

    def __init__(self, name: str, nbytes: float):
        self.name, self.nbytes = name, float(nbytes)

    def __enter__(self):
        if _spans is not None:
            import torch
            self.e0, self.e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            self.e0.record()
        return self

    def __exit__(self, *exc):
        if _spans is not None:
            self.e1.record()
            _spans.append((self.name, self.nbytes, self.e0, self.e1))
        return False


def profile_begin() -> None:
    global _spans
    _spans = []


def profile_end():
    """[(name, algorithmic bytes, milliseconds)] of the spans recorded since profile_begin()."""
    global _spans
    import torch
    torch.cuda.synchronize()
    out = [(n, b, e0.elapsed_time(e1)) for n, b, e0, e1 in (_spans or [])]
    _spans = None
    return out


def launch_count() -> int:
    return int(lib().b2f_launch_count())
