"""The two optional PyNQS edits shipped as integration/patches/*.patch: they apply cleanly to the reference sources
(baseline/_ref copy), the patched files compile, and -- on a GPU -- the patched `_only_sample_space` (one-pass operator)
reproduces the goldens of the unpatched reference."""
import os
import py_compile
import shutil
import subprocess
import sys
import tempfile

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")
PATCHES = os.path.join(ROOT, "integration", "patches")

needs_ref = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "vmc")), reason="baseline/_ref missing (python baseline/make_ref.py)")


def _patched_copy(tmp):
    dst = os.path.join(tmp, "pynqs")
    shutil.copytree(REF, dst)
    for name in sorted(os.listdir(PATCHES)):
        subprocess.run(["patch", "-p1", "--no-backup-if-mismatch", "-i", os.path.join(PATCHES, name)], cwd=dst, check=True, capture_output=True)
    return dst


@needs_ref
def test_patches_apply_cleanly_and_compile():
    assert sorted(os.listdir(PATCHES)) == ["eloc_sample_space.patch", "gather_scatter_sample.patch"]
    with tempfile.TemporaryDirectory() as tmp:
        dst = _patched_copy(tmp)
        for rel in ("vmc/energy/eloc.py", "vmc/sample.py"):
            py_compile.compile(os.path.join(dst, rel), doraise=True)
        assert "_ELOC_ONE_PASS" in open(os.path.join(dst, "vmc/energy/eloc.py")).read()
        assert "pynqs_b200.compat.sampler" in open(os.path.join(dst, "vmc/sample.py")).read()


@needs_ref
@pytest.mark.gpu
@pytest.mark.parametrize("tag,cplx", [("real", False), ("complex", True)])
def test_patched_only_sample_space_matches_the_reference_golden(tag, cplx):
    import torch

    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from util import fe2s2, load

    from pynqs_b200 import synthetic as S

    code = r'''
import sys, numpy as np, torch
sys.path[:0] = [sys.argv[1], sys.argv[2]]          # patched PyNQS copy, this repo (libs/ shim)
sys.path.insert(0, sys.argv[2] + "/tests")
from util import fe2s2, load
from pynqs_b200 import synthetic as S
import vmc.energy.eloc as E
from utils.public_function import WavefunctionLUT
assert E._ELOC_ONE_PASS is not None
cplx = sys.argv[3] == "1"
f, g = fe2s2(), load("eloc_fe2s2_" + ("complex" if cplx else "real"))
dev = "cuda:0"
d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
dtype = torch.complex128 if cplx else torch.double
psi = S.random_psi(f["ci"].shape[0], seed=int(g["psi_seed"]), complex_=cplx)
lut = WavefunctionLUT(d(f["ci"]), d(psi).to(dtype), f["sorb"], dev)
first, n = int(g["first"]), int(g["n"])
eloc, sloc, psi_x, _ = E.local_energy(d(f["ci"][first:first + n].copy()), d(f["h1e"]), d(f["h2e"]), None, None, f["sorb"], f["nele"],
                                      f["noA"], f["noB"], dtype=dtype, WF_LUT=lut, use_sample_space=True)
np.testing.assert_allclose(eloc.cpu().numpy(), g["eloc"], rtol=1e-12, atol=0)
np.testing.assert_array_equal(psi_x.cpu().numpy(), g["psi_x"])
print("patched ok")
'''
    with tempfile.TemporaryDirectory() as tmp:
        dst = _patched_copy(tmp)
        r = subprocess.run([sys.executable, "-c", code, dst, ROOT, "1" if cplx else "0"], capture_output=True, text=True, timeout=600)
        assert r.returncode == 0 and "patched ok" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
