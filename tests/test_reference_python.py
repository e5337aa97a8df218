"""The UNMODIFIED reference Python (vmc/energy/eloc.py, vmc/energy/flip.py, utils/public_function.py of PyNQS,
copied verbatim into the git-ignored baseline/_ref by baseline/make_ref.py) executed on a B200 with
`libs.C_extension` = this repository's shim (libs/C_extension.py -> pynqs_b200.C_extension -> C ABI).

What runs here is the reference's own code: `local_energy` with ElocMethod SAMPLE_SPACE, REDUCE (deterministic
and semi-stochastic) and SIMPLE, `Func` (LUT lookup + torch.unique of the misses), and the reference's own
`WavefunctionLUT` / `torch_sort_onv` on CUDA tensors.  Expected values are the goldens the same code produced
on the reference's CPU extension (tests/golden/make_golden.py).  Bars: E_loc 1e-12 relative, psi(x) exact."""
import os
import sys

import numpy as np
import pytest
import torch

from util import fe2s2, load, toy_amplitude

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


@pytest.fixture(scope="module")
def ref_py():
    """Import the reference packages against the shim; skip when baseline/_ref was not made (no reference mounted)."""
    sys.path.insert(0, os.path.join(ROOT, "baseline"))
    import make_ref  # baseline/make_ref.py

    sys.path.pop(0)
    if not make_ref.available():
        pytest.skip("baseline/_ref missing: run `python baseline/make_ref.py` where /root/reference is mounted")
    from pynqs_b200 import _lib

    _lib.load()
    assert torch.cuda.is_available()
    make_ref.on_path(ROOT)
    import libs.C_extension as shim

    assert os.path.dirname(os.path.abspath(shim.__file__)) == os.path.join(ROOT, "libs"), "libs.C_extension is not this repo's shim"
    import utils.public_function as pf
    import vmc.energy.eloc as eloc_mod
    import vmc.energy.flip as flip_mod

    assert os.path.abspath(eloc_mod.__file__).startswith(os.path.join(ROOT, "baseline", "_ref"))
    assert eloc_mod.FUSED_HIJ, "the reference did not find get_comb_hij_fused in the shim"
    return dict(pf=pf, eloc=eloc_mod, flip=flip_mod, shim=shim)


def _ansatz(sorb, cplx, shim, dtype):
    def ansatz(states):
        return toy_amplitude(states, sorb, cplx)

    def batcher(x, func):
        return func(shim.onv_to_tensor(x, sorb)).to(dtype)

    return ansatz, batcher


@pytest.mark.parametrize("tag,cplx", [("real", False), ("complex", True)])
def test_reference_local_energy_sample_space(ref_py, tag, cplx):
    """vmc/energy/eloc.py:23-132 -> _only_sample_space (:326-508) with the reference's WavefunctionLUT on CUDA."""
    f = fe2s2()
    g = load(f"eloc_fe2s2_{tag}")
    from pynqs_b200 import synthetic as S

    dtype = torch.complex128 if cplx else torch.double
    psi = S.random_psi(f["ci"].shape[0], seed=int(g["psi_seed"]), complex_=cplx)
    lut = ref_py["pf"].WavefunctionLUT(dev(f["ci"]), dev(psi).to(dtype), f["sorb"], DEV)
    first, n = int(g["first"]), int(g["n"])
    x = dev(f["ci"][first : first + n].copy())
    eloc, sloc, psi_x, _ = ref_py["eloc"].local_energy(
        x, dev(f["h1e"]), dev(f["h2e"]), None, None, f["sorb"], f["nele"], f["noA"], f["noB"], dtype=dtype, WF_LUT=lut,
        use_sample_space=True)
    np.testing.assert_allclose(eloc.cpu().numpy(), g["eloc"], rtol=1e-12, atol=0)
    np.testing.assert_array_equal(psi_x.cpu().numpy(), g["psi_x"])
    assert float(sloc.abs().max()) == 0.0


@pytest.mark.parametrize("with_lut", [False, True])
@pytest.mark.parametrize("tag,cplx", [("real", False), ("complex", True)])
def test_reference_local_energy_reduce_deterministic(ref_py, tag, cplx, with_lut):
    """_reduce_psi (eloc.py:204-323) with eps > 0, eps_sample = 0 and Func (flip.py:29-63) on the shim."""
    f = fe2s2()
    g = load("reduce_fe2s2")
    from pynqs_b200 import synthetic as S

    dtype = torch.complex128 if cplx else torch.double
    ansatz, batcher = _ansatz(f["sorb"], cplx, ref_py["shim"], dtype)
    first, n, eps = int(g["first"]), int(g["n"]), float(g["eps"])
    x = dev(f["ci"][first : first + n].copy())
    lut = None
    if with_lut:
        psi_tab = S.random_psi(f["ci"].shape[0] // 2, seed=43, complex_=cplx)
        lut = ref_py["pf"].WavefunctionLUT(dev(f["ci"][::2].copy()), dev(psi_tab).to(dtype), f["sorb"], DEV)
    eloc, _, psi_x, _ = ref_py["eloc"].local_energy(
        x, dev(f["h1e"]), dev(f["h2e"]), ansatz, batcher, f["sorb"], f["nele"], f["noA"], f["noB"], dtype=dtype, WF_LUT=lut,
        use_unique=True, reduce_psi=True, eps=eps, eps_sample=0)
    key = "_lut_" if with_lut else "_"
    np.testing.assert_allclose(eloc.cpu().numpy(), g[f"eloc{key}{tag}"], rtol=1e-12, atol=0)
    np.testing.assert_allclose(psi_x.cpu().numpy(), g[f"psi_x{key}{tag}"], rtol=1e-14, atol=0)


@pytest.mark.parametrize("with_lut", [False, True])
@pytest.mark.parametrize("tag,cplx", [("real", False), ("complex", True)])
def test_reference_local_energy_simple(ref_py, tag, cplx, with_lut):
    """_simple (eloc.py:134-202): get_comb_tensor + get_hij_torch + Func over all M determinants."""
    f = fe2s2()
    g = load("simple_fe2s2")
    from pynqs_b200 import synthetic as S

    dtype = torch.complex128 if cplx else torch.double
    ansatz, batcher = _ansatz(f["sorb"], cplx, ref_py["shim"], dtype)
    first, n = int(g["first"]), int(g["n"])
    x = dev(f["ci"][first : first + n].copy())
    lut = None
    if with_lut:
        psi_tab = S.random_psi(f["ci"].shape[0] // 2, seed=43, complex_=cplx)
        lut = ref_py["pf"].WavefunctionLUT(dev(f["ci"][::2].copy()), dev(psi_tab).to(dtype), f["sorb"], DEV)
    eloc, _, psi_x, _ = ref_py["eloc"].local_energy(
        x, dev(f["h1e"]), dev(f["h2e"]), ansatz, batcher, f["sorb"], f["nele"], f["noA"], f["noB"], dtype=dtype, WF_LUT=lut,
        use_unique=True)
    key = "_lut_" if with_lut else "_"
    np.testing.assert_allclose(eloc.cpu().numpy(), g[f"eloc{key}{tag}"], rtol=1e-12, atol=0)
    np.testing.assert_allclose(psi_x.cpu().numpy(), g[f"psi_x{key}{tag}"], rtol=1e-14, atol=0)


def test_reference_semi_stochastic_reduce_runs_and_is_unbiased(ref_py):
    """_reduce_psi with eps > 0 and eps_sample > 0 (eloc.py:257-283, the setting of every shipped input): torch's
    multinomial stream differs between devices, so the check is statistical -- the mean over repeated draws agrees
    with the exact (SIMPLE) local energy within 5 standard errors."""
    f = fe2s2()
    g = load("simple_fe2s2")
    dtype = torch.double
    ansatz, batcher = _ansatz(f["sorb"], False, ref_py["shim"], dtype)
    first, n = int(g["first"]), int(g["n"])
    x = dev(f["ci"][first : first + n].copy())
    torch.manual_seed(1234)
    draws = []
    for _ in range(48):
        eloc, _, _, _ = ref_py["eloc"].local_energy(
            x, dev(f["h1e"]), dev(f["h2e"]), ansatz, batcher, f["sorb"], f["nele"], f["noA"], f["noB"], dtype=dtype, WF_LUT=None,
            use_unique=True, reduce_psi=True, eps=1e-2, eps_sample=1000)
        draws.append(eloc.cpu().numpy())
    draws = np.stack(draws)
    mean, se = draws.mean(0), draws.std(0, ddof=1) / np.sqrt(draws.shape[0])
    assert np.all(np.abs(mean - g["eloc_real"]) <= 5 * se + 1e-9), (mean - g["eloc_real"], se)


def test_reference_wavefunction_lut_class_on_cuda(ref_py):
    """utils/public_function.py:749-868 unmodified: torch_sort_onv on CUDA tensors + the shim's wavefunction_lut."""
    from pynqs_b200 import synthetic as S

    g = load("lut_lookup_l1")
    rng = np.random.default_rng(31)
    keys = S.random_onvs(5000, 40, 15, 15, seed=32)
    psi = S.random_psi(5000, seed=33)
    lut = ref_py["pf"].WavefunctionLUT(dev(keys), dev(psi), 40, DEV)
    q = np.concatenate([keys[rng.permutation(5000)[:700]], S.random_onvs(700, 40, 15, 15, seed=34)])
    hit, miss, val = lut.lookup(dev(q))
    np.testing.assert_array_equal(hit.cpu().numpy(), g["hit"])
    np.testing.assert_array_equal(miss.cpu().numpy(), g["miss"])
    np.testing.assert_array_equal(val.cpu().numpy(), g["val"])
    idx, mask = ref_py["shim"].wavefunction_lut(lut.bra_key, dev(q), 40)
    np.testing.assert_array_equal(idx.cpu().numpy(), g["idx"])
    np.testing.assert_array_equal(mask.cpu().numpy(), g["mask"])
