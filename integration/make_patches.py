"""Writes integration/patches/*.patch: the two optional edits to PyNQS that switch its own code to the additive fast paths
(everything else runs UNMODIFIED on the `libs/C_extension.py` shim -- tests/test_reference_python.py).

    python integration/make_patches.py            # where /root/reference is mounted

  eloc_sample_space.patch   vmc/energy/eloc.py: `_only_sample_space` calls the one-pass operator (no [n, M] arrays) when the
                            extension provides it; spin-raising / multi-psi runs keep the three-call body below it.
  gather_scatter_sample.patch  vmc/sample.py: `Sampler.gather_scatter_sample` delegates to pynqs_b200.compat.sampler (one size
                            exchange + one all-gather per column + identical merge on every rank) when it is importable.

Apply with `patch -p1 < integration/patches/<name>.patch` in the PyNQS checkout.  tests/test_patches.py applies them to a
scratch copy of the reference and runs the patched functions.
"""
from __future__ import annotations

import difflib
import os

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE_ROOT = os.environ.get("PYNQS_REFERENCE_ROOT", "/root/reference")

ELOC_ANCHOR = "    check_para(x)\n\n    device = x.device\n    dim: int = x.dim()\n    assert dim == 2\n    t0 = time.time_ns()\n\n    batch = x.size(0)\n    nSD = get_Num_SinglesDoubles(sorb, noa, nob) + 1\n"
ELOC_INSERT = '''    check_para(x)

    # pynqs_b200: the whole body in one pass on the GPU (no comb / Hmat / psi arrays of size [batch, nSD])
    if _ELOC_ONE_PASS is not None and x.is_cuda and not use_spin_raising and not use_multi_psi:
        t0 = time.time_ns()
        eloc, psi_x = _ELOC_ONE_PASS(
            x, h1e, h2e, sorb, nele, noa, nob, WF_LUT.bra_key, WF_LUT.wf_value, getattr(WF_LUT, "group_index", None)
        )
        delta = (time.time_ns() - t0) / 1.0e06
        return eloc.to(dtype), torch.zeros_like(eloc).to(dtype), psi_x.to(dtype), (0.0, delta, 0.0)

'''
ELOC_IMPORT_ANCHOR = "FUSED_HIJ = True\ntry:\n    from libs.C_extension import get_comb_hij_fused\nexcept ImportError:\n    FUSED_HIJ = False\n"
ELOC_IMPORT_INSERT = ELOC_IMPORT_ANCHOR + "\ntry:  # additive operator of pynqs_b200 (absent from the stock extension)\n    from libs.C_extension import eloc_sample_space as _ELOC_ONE_PASS\nexcept ImportError:\n    _ELOC_ONE_PASS = None\n"

SAMPLE_ANCHOR = "        t0 = time.time_ns()\n        # Gather unique, counts, wf_value\n"
SAMPLE_INSERT = '''        try:  # pynqs_b200: one all-gather per column and an identical merge on every rank (same return values)
            from pynqs_b200.compat.sampler import gather_scatter_sample as _fast_exchange
        except ImportError:
            _fast_exchange = None
        if _fast_exchange is not None:
            return _fast_exchange(self, unique, counts, wf_value)

'''


def patched_eloc(src: str) -> str:
    assert src.count(ELOC_IMPORT_ANCHOR) == 1, "eloc.py: import anchor not found"
    src = src.replace(ELOC_IMPORT_ANCHOR, ELOC_IMPORT_INSERT)
    head, sep, tail = src.partition("def _only_sample_space(")
    assert sep and tail.count(ELOC_ANCHOR) == 1, "eloc.py: _only_sample_space anchor not found"
    tail = tail.replace(ELOC_ANCHOR, ELOC_INSERT + ELOC_ANCHOR[len("    check_para(x)\n\n"):], 1)
    return head + sep + tail


def patched_sample(src: str) -> str:
    head, sep, tail = src.partition("    def gather_scatter_sample(")
    assert sep and tail.count(SAMPLE_ANCHOR) >= 1, "sample.py: gather_scatter_sample anchor not found"
    return head + sep + tail.replace(SAMPLE_ANCHOR, SAMPLE_INSERT + SAMPLE_ANCHOR, 1)


PATCHES = {
    "eloc_sample_space.patch": ("vmc/energy/eloc.py", patched_eloc),
    "gather_scatter_sample.patch": ("vmc/sample.py", patched_sample),
}


def make(root: str = REFERENCE_ROOT, out_dir: str = os.path.join(HERE, "patches")) -> None:
    os.makedirs(out_dir, exist_ok=True)
    for name, (rel, fn) in PATCHES.items():
        old = open(os.path.join(root, rel)).read()
        new = fn(old)
        diff = difflib.unified_diff(old.splitlines(keepends=True), new.splitlines(keepends=True), "a/" + rel, "b/" + rel, n=3)
        open(os.path.join(out_dir, name), "w").write("".join(diff))
        print(name, "->", rel)


if __name__ == "__main__":
    make()
