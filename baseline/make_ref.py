"""Populate baseline/_ref/ with the UNMODIFIED reference Python the local-energy path runs through
(utils/, vmc/ of /root/reference), so that it can be executed on the GPU box against this repo's
`libs/C_extension.py` shim.

    python baseline/make_ref.py        # in the build container, where /root/reference is mounted

baseline/_ref/ is git-ignored (no reference source enters the history) but NOT gpurun-ignored, so it
travels to the box like a built .so.  Nothing under baseline/_ref is imported by the product; only
tests/test_reference_python.py and bench.py's reference arms put it on sys.path.
"""
from __future__ import annotations

import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
REFERENCE_ROOT = os.environ.get("PYNQS_REFERENCE_ROOT", "/root/reference")
PACKAGES = ("utils", "vmc")


def available() -> bool:
    return all(os.path.isdir(os.path.join(DEST, p)) for p in PACKAGES)


def populate(force: bool = False) -> str:
    if available() and not force:
        return DEST
    if not os.path.isdir(os.path.join(REFERENCE_ROOT, "vmc")):
        raise FileNotFoundError(f"{REFERENCE_ROOT} not mounted: baseline/_ref can only be made in the build container")
    os.makedirs(DEST, exist_ok=True)
    for p in PACKAGES:
        dst = os.path.join(DEST, p)
        if os.path.isdir(dst):
            shutil.rmtree(dst)
        shutil.copytree(os.path.join(REFERENCE_ROOT, p), dst, ignore=shutil.ignore_patterns("__pycache__", "*.pyc", "*.so"))
    return DEST


def on_path(libs_root: str) -> None:
    """sys.path so that `libs.C_extension` resolves to <libs_root>/libs and `utils`, `vmc` to the reference copy."""
    for p in (DEST, libs_root):
        if p in sys.path:
            sys.path.remove(p)
        sys.path.insert(0, p)


if __name__ == "__main__":
    print(populate(force="--force" in sys.argv))
