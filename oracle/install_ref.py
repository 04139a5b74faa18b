"""TEST / BASELINE INFRASTRUCTURE -- not part of the product path.

Recipe that places the reference's own hot-path sources under baseline/_ref/ so that the GPU box (where
/root/reference does not exist) can time and check against the REFERENCE ITSELF rather than the restated port:
    python oracle/install_ref.py            (also run by __graft_entry__.build() when the reference tree is present)
The reference is an application, not a package (no setup.py / pyproject.toml): `pip install --target baseline/_ref
/root/reference` has nothing to build, so the four pure-Python files of the path are placed there directly.
baseline/_ref/ is git-ignored (reference sources never enter this repository's history) but not gpurun-ignored, so
it travels with the snapshot.  oracle/ref_loader.py imports from it unmodified, over oracle/shims.py for the two
wheels the image lacks (onnxruntime -> torch-CPU fp32 session, scikit-image -> Umeyama)."""
from __future__ import annotations

import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FILES = ["models/__init__.py", "models/scrfd.py", "models/arcface.py", "utils/helpers.py"]


def install(src_root: str = None, dst_root: str = None) -> str:
    src_root = src_root or os.environ.get("B2F_REFERENCE") or "/root/reference"
    dst_root = dst_root or os.path.join(ROOT, "baseline", "_ref")
    if not os.path.isfile(os.path.join(src_root, "models", "scrfd.py")):
        return ""
    for rel in FILES:
        dst = os.path.join(dst_root, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(src_root, rel), dst)
    return dst_root


if __name__ == "__main__":
    out = install(*sys.argv[1:3])
    print(f"reference hot-path sources installed under {out}" if out else "reference tree not found: nothing installed")
