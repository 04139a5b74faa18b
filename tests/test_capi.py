"""CPU: the C-ABI library builds for sm_100a, loads, and exports every symbol include/b2f.h declares.
No compute entry is called here (there is no GPU on the build box and no CPU fallback to call)."""
import ctypes
import os
import subprocess

import pytest

from scrfd_arcface_facerecognition_b200 import _lib


def test_library_builds_and_loads():
    path = _lib.build()
    assert os.path.exists(path)
    lib = _lib.lib()
    assert lib.b2f_version() == 4
    assert lib.b2f_launch_count() == 0 or lib.b2f_launch_count() > 0


def test_every_declared_symbol_is_exported_and_bound():
    declared = _lib.declared_symbols()
    assert len(declared) >= 25
    handle = ctypes.CDLL(_lib.SO_PATH)
    missing = [s for s in declared if not hasattr(handle, s)]
    assert not missing, f"declared in include/b2f.h but not exported: {missing}"
    assert sorted(_lib.SIGNATURES) == declared, "ctypes signature table and header disagree"


def test_sass_contains_tcgen05_and_tma():
    out = subprocess.run(["cuobjdump", "-sass", _lib.SO_PATH], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    assert "UTCHMMA" in out.stdout and "UTMALDG" in out.stdout and "LDTM" in out.stdout
    assert "HMMA.16" not in out.stdout                       # no legacy mma.sync path


def test_product_modules_do_not_import_the_oracle():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, "scrfd_arcface_facerecognition_b200")
    offenders = []
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                text = open(os.path.join(dirpath, f)).read()
                if "import oracle" in text or "from oracle" in text:
                    offenders.append(f)
    for f in ("models/__init__.py", "models/scrfd.py", "models/arcface.py", "utils/helpers.py"):
        text = open(os.path.join(root, f)).read()
        if "oracle" in text:
            offenders.append(f)
    assert not offenders


def test_compute_without_cuda_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from scrfd_arcface_facerecognition_b200 import helpers
    import numpy as np
    with pytest.raises(_lib.B2FError):
        helpers.compute_similarity(np.ones(4, np.float32), np.ones(4, np.float32))
    from models import SCRFD
    with pytest.raises(_lib.B2FError):
        SCRFD("weights/det_500m.onnx")


def test_fastdiv_formula_is_exact():
    """The multiply-shift division the conv kernels use for tile coordinates (csrc/conv_tile.cu, `FastDiv`):
    q = (x * ceil(2^(31+s) / d)) >> (31+s) with s = ceil(log2 d) must equal x // d for every 0 <= x < 2^31, and the
    multiplier must fit 32 bits."""
    import numpy as np
    rng = np.random.default_rng(0)
    divisors = list(range(1, 70)) + [98, 100, 127, 128, 129, 392, 784, 1000, 3136, 4097, 65535, 65536, 1 << 20, (1 << 20) + 1]
    for d in divisors:
        s = 0
        while (1 << s) < d:
            s += 1
        shift = 31 + s
        mul = ((1 << shift) + d - 1) // d
        assert mul < (1 << 32), d
        xs = np.concatenate([rng.integers(0, 1 << 31, 4000, dtype=np.int64),
                             np.arange(0, 4 * d + 2, dtype=np.int64),
                             (np.arange(1, 2000, dtype=np.int64) * d - 1) % (1 << 31),
                             np.array([(1 << 31) - 1, (1 << 31) - d, ((1 << 31) - 1) // d * d], dtype=np.int64)])
        xs = xs[(xs >= 0) & (xs < (1 << 31))]
        q = np.array([(int(x) * mul) >> shift for x in xs], dtype=np.int64)
        assert (q == xs // d).all(), d
