"""Copies the reference's Python package (read-only /root/reference/egom2p, authoring container only) to baseline/_ref/
(git-ignored, NOT gpurun-ignored: it travels to the GPU box) so that `baseline/ref_gpu.py` can time the UNMODIFIED
reference module on the same B200. The reference ships no setup.py / pyproject.toml, so `pip install --target` cannot
build it (recorded in DESIGN.md); the package is pure Python and runs from a plain copy. Nothing under baseline/_ref is
imported by the product package, the tests or the C-ABI library.   python tools/install_reference.py"""
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference/egom2p"
DST = os.path.join(ROOT, "baseline", "_ref", "egom2p")

def install(force: bool = True) -> int:
    """Copies the package; returns the number of files under baseline/_ref/egom2p (0 if the reference tree is absent)."""
    if not os.path.isdir(SRC):
        return 0
    if os.path.isdir(DST):
        if not force:
            return sum(len(f) for _, _, f in os.walk(DST))
        shutil.rmtree(DST)
    shutil.copytree(SRC, DST, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    return sum(len(f) for _, _, f in os.walk(DST))


if __name__ == "__main__":
    if not os.path.isdir(SRC):
        sys.exit(f"{SRC} not found (the reference tree exists in the authoring container only)")
    print(f"copied {install()} files to {DST}")
