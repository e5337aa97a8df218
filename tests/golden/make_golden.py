"""Generate tests/golden/*.npz from the UNMODIFIED reference (run in the build container only).

    python oracle/build_ref.py          # reference CPU extension -> oracle/_ref/
    python tests/golden/make_golden.py  # this script

Sources of truth used here:
  * the reference extension built from /root/reference/cpp_src (oracle/_ref/C_extension_L{1,2,3}.so):
    get_comb_tensor, get_comb_hij_fused, get_hij_torch, wavefunction_lut, onv_to_tensor, tensor_to_onv;
  * the reference Python imported from /root/reference: utils.public_function.WavefunctionLUT /
    torch_sort_onv and vmc.energy.eloc._only_sample_space (driven on the reference extension);
  * the Fe2S2 CAS(30e,20o) integrals and determinants of /root/reference/example/Fe2S2/fe2s2-OO.pth
    (copied into fe2s2_integrals.npz as data: the config-2 Hamiltonian).
Inputs are regenerated at test time from seeds (pynqs_b200/synthetic.py), so fixtures only hold
outputs: full arrays where small, otherwise SHA-256 digests plus strided samples.
"""
from __future__ import annotations

import hashlib
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle.build_ref import load_ref  # noqa: E402
from pynqs_b200 import synthetic as S  # noqa: E402

REFERENCE = "/root/reference"


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def save(name, **kw):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **kw)
    print(f"{name}.npz  {os.path.getsize(path) / 1024:.0f} KiB")


def t(a):
    return torch.from_numpy(np.ascontiguousarray(a))


def ops_case(name, sorb, noA, noB, n, seed, symmetric, keep_rows, dtype=np.float64, stride=1):
    L = (sorb - 1) // 64 + 1
    ref = load_ref(L)
    nele = noA + noB
    bra = S.random_onvs(n, sorb, noA, noB, seed=seed)
    h1e, h2e = S.random_packed_integrals(sorb, seed=seed + 1, symmetric=symmetric, dtype=dtype)
    comb, hmat = ref.get_comb_hij_fused(t(bra), t(h1e), t(h2e), sorb, nele, noA, noB)
    comb2, _ = ref.get_comb_tensor(t(bra), sorb, nele, noA, noB, False)
    assert torch.equal(comb, comb2)
    hmat3 = ref.get_hij_torch(t(bra), comb, t(h1e), t(h2e), sorb, nele)
    assert torch.equal(hmat3, hmat), "reference fused != unfused"
    comb, hmat = comb.numpy(), hmat.numpy()
    out = dict(
        sorb=sorb, noA=noA, noB=noB, n=bra.shape[0], seed=seed, symmetric=int(symmetric), stride=stride,
        dtype=np.dtype(dtype).name, bra_sha=sha(bra), h2e_sha=sha(h2e),
        comb_sha=sha(comb), hmat_sha=sha(hmat),
        comb_rows=comb[:keep_rows, ::stride], hmat_rows=hmat[:keep_rows, ::stride],
    )
    # dense matrix mode (2-D ket) on a small block
    m2 = min(bra.shape[0], 24)
    out["hij2d"] = ref.get_hij_torch(t(bra[:m2]), t(bra[:m2]), t(h1e), t(h2e), sorb, nele).numpy()
    save(name, **out)


def main():
    torch.set_num_threads(os.cpu_count())
    # ---- operator goldens --------------------------------------------------------------------
    ops_case("ops_c1_h6_12sorb", 12, 3, 3, 64, 11, False, keep_rows=64)
    ops_case("ops_c1_h6_12sorb_f32", 12, 3, 3, 32, 12, False, keep_rows=32, dtype=np.float32)
    ops_case("ops_odd_14sorb_4a2b", 14, 4, 2, 33, 13, False, keep_rows=8)          # M even/odd mix, noA != noB
    ops_case("ops_c3_n2_52sorb", 52, 5, 5, 6, 14, True, keep_rows=2)
    ops_case("ops_c4_h50_100sorb", 100, 25, 25, 1, 15, True, keep_rows=1, stride=997)
    ops_case("ops_l3_132sorb_3a2b", 132, 3, 2, 3, 16, False, keep_rows=1, stride=13)

    # ---- Fe2S2 (config 2): real integrals ------------------------------------------------------
    d = torch.load(os.path.join(REFERENCE, "example/Fe2S2/fe2s2-OO.pth"), weights_only=False)
    h1e, h2e = d["h1e"].numpy(), d["h2e"].numpy()
    ci = d["ci_space"].numpy()
    sorb, noA, noB, nele = int(d["sorb"]), int(d["noa"]), int(d["nob"]), int(d["nele"])
    save("fe2s2_integrals", h1e=h1e, h2e=h2e, ci_space=ci, sorb=sorb, noA=noA, noB=noB, nele=nele,
         e_ref=np.asarray(d["e_lst"], dtype=np.float64))
    ref = load_ref(1)
    nf = 16
    comb, hmat = ref.get_comb_hij_fused(t(ci[:nf]), t(h1e), t(h2e), sorb, nele, noA, noB)
    save("ops_c2_fe2s2", n=nf, comb_sha=sha(comb.numpy()), hmat_sha=sha(hmat.numpy()),
         comb_rows=comb.numpy()[:2], hmat_rows=hmat.numpy()[:4])

    # ---- reference Python: LUT sort order, lookup, sample-space E_loc ---------------------------
    libs = types.ModuleType("libs")
    libs.__path__ = []
    sys.modules["libs"] = libs
    sys.modules["libs.C_extension"] = ref
    libs.C_extension = ref
    sys.path.insert(0, REFERENCE)
    from utils.public_function import WavefunctionLUT, torch_sort_onv  # reference code
    from vmc.energy.eloc import _only_sample_space  # reference code

    # sort order incl. duplicate rows (stability) for L = 1, 2, 3
    sort_out = {}
    for L, sorb_l, na in ((1, 40, 15), (2, 100, 25), (3, 132, 3)):
        k = S.random_onvs(300, sorb_l, na, na, seed=20 + L)
        k = np.concatenate([k, k[:40]])[np.random.default_rng(5).permutation(340)]
        sort_out[f"idx_L{L}"] = torch_sort_onv(t(k)).numpy()
    save("lut_sort_order", **sort_out)

    # lookup golden through the reference class (hits and misses)
    rng = np.random.default_rng(31)
    keys = S.random_onvs(5000, 40, 15, 15, seed=32)
    psi = S.random_psi(5000, seed=33)
    lut = WavefunctionLUT(t(keys), t(psi), 40, "cpu")
    q = np.concatenate([keys[rng.permutation(5000)[:700]], S.random_onvs(700, 40, 15, 15, seed=34)])
    idx, mask = load_ref(1).wavefunction_lut(lut.bra_key, t(q), 40)
    hit, miss, val = lut.lookup(t(q))
    save("lut_lookup_l1", idx=idx.numpy(), mask=mask.numpy(), hit=hit.numpy(), miss=miss.numpy(), val=val.numpy())

    # E_loc (sample-space) on Fe2S2: table = ci_space (18496 dets), psi random, real and complex
    for tag, cplx in (("real", False), ("complex", True)):
        psi = S.random_psi(ci.shape[0], seed=41, complex_=cplx)
        dtype = torch.complex128 if cplx else torch.double
        lut = WavefunctionLUT(t(ci), t(psi).to(dtype), sorb, "cpu")
        x = t(ci[1000:1064].copy())
        eloc, _, psi_x, _ = _only_sample_space(x, t(h1e), t(h2e), None, None, sorb, nele, noA, noB, dtype=dtype, WF_LUT=lut)
        save(f"eloc_fe2s2_{tag}", eloc=eloc.numpy(), psi_x=psi_x.numpy(), first=1000, n=64, psi_seed=41)

    # E_loc on random 12-sorb space: every sample in the table, full space (all 400 dets)
    keys = S.random_onvs(400, 12, 3, 3, seed=51)
    assert keys.shape[0] == 400
    h1s, h2s = S.random_packed_integrals(12, seed=52, symmetric=True)
    psi = S.random_psi(400, seed=53)
    lut = WavefunctionLUT(t(keys), t(psi), 12, "cpu")
    eloc, _, psi_x, _ = _only_sample_space(t(keys), t(h1s), t(h2s), None, None, 12, 6, 3, 3, dtype=torch.double, WF_LUT=lut)
    save("eloc_c1_fullspace", eloc=eloc.numpy(), psi_x=psi_x.numpy())


def toy_amplitude(states: torch.Tensor, sorb: int, cplx: bool) -> torch.Tensor:
    """A deterministic stand-in for the ansatz of the REDUCE goldens: psi(x) from the +-1 occupation tensor.
    (tests/util.py holds the same function for the test side.)"""
    g = torch.Generator().manual_seed(77)
    w = torch.randn(sorb, 2, generator=g, dtype=torch.float64)
    z = states.to(torch.float64) @ w
    amp = 0.3 + torch.tanh(0.11 * z[:, 0]) ** 2
    return amp * torch.exp(1j * 0.7 * z[:, 1]) if cplx else amp * torch.sign(torch.cos(0.9 * z[:, 1]))


def main_reduce():
    """REDUCE-method goldens (vmc.energy.eloc._reduce_psi of the reference, eps > 0, eps_sample = 0) on the Fe2S2
    integrals: ansatz only, and ansatz + WavefunctionLUT; plus the kept set (torch.where(|H| >= eps)) itself."""
    torch.set_num_threads(os.cpu_count())
    ref = load_ref(1)
    libs = types.ModuleType("libs")
    libs.__path__ = []
    sys.modules["libs"] = libs
    sys.modules["libs.C_extension"] = ref
    libs.C_extension = ref
    sys.path.insert(0, REFERENCE)
    from utils.public_function import WavefunctionLUT  # reference code
    from vmc.energy.eloc import _reduce_psi  # reference code

    d = torch.load(os.path.join(REFERENCE, "example/Fe2S2/fe2s2-OO.pth"), weights_only=False)
    h1e, h2e = d["h1e"], d["h2e"]
    ci = d["ci_space"].numpy()
    sorb, noA, noB, nele = int(d["sorb"]), int(d["noa"]), int(d["nob"]), int(d["nele"])
    first, n, eps = 2000, 24, 1.0e-3
    x = t(ci[first : first + n].copy())
    comb = ref.get_comb_tensor(x, sorb, nele, noA, noB, False)[0]
    hij = ref.get_hij_torch(x, comb, h1e, h2e, sorb, nele)
    idx = torch.where(hij.reshape(-1).abs() >= eps)[0]
    out = dict(first=first, n=n, eps=eps, K=int(idx.numel()), idx_sha=sha(idx.numpy()),
               hij_sha=sha(hij.reshape(-1)[idx].numpy()), x_sha=sha(comb.reshape(-1, comb.size(2))[idx].numpy()),
               idx_head=idx.numpy()[:64], counts=(hij.abs() >= eps).sum(1).numpy())
    for tag, cplx in (("real", False), ("complex", True)):
        dtype = torch.complex128 if cplx else torch.double

        def ansatz(states):
            return toy_amplitude(states, sorb, cplx)

        def batcher(x, func):
            return func(ref.onv_to_tensor(x, sorb)).to(dtype)

        eloc, _, psi_x, _ = _reduce_psi(x, h1e, h2e, ansatz, batcher, sorb, nele, noA, noB, dtype=dtype, WF_LUT=None,
                                        use_unique=True, eps=eps, eps_sample=0)
        out[f"eloc_{tag}"] = eloc.numpy()
        out[f"psi_x_{tag}"] = psi_x.numpy()
        # with a LUT over half of the CI space (values differ from the ansatz on purpose)
        psi_tab = S.random_psi(ci.shape[0] // 2, seed=43, complex_=cplx)
        lut = WavefunctionLUT(t(ci[::2].copy()), t(psi_tab).to(dtype), sorb, "cpu")
        eloc, _, psi_x, _ = _reduce_psi(x, h1e, h2e, ansatz, batcher, sorb, nele, noA, noB, dtype=dtype, WF_LUT=lut,
                                        use_unique=True, eps=eps, eps_sample=0)
        out[f"eloc_lut_{tag}"] = eloc.numpy()
        out[f"psi_x_lut_{tag}"] = psi_x.numpy()
    save("reduce_fe2s2", **out)


def main_simple():
    """SIMPLE-method golden (vmc.energy.eloc._simple of the reference: every connected determinant goes through
    the ansatz, via Func with and without a LUT) on the Fe2S2 integrals, toy ansatz."""
    torch.set_num_threads(os.cpu_count())
    ref = load_ref(1)
    libs = types.ModuleType("libs")
    libs.__path__ = []
    sys.modules["libs"] = libs
    sys.modules["libs.C_extension"] = ref
    libs.C_extension = ref
    sys.path.insert(0, REFERENCE)
    from utils.public_function import WavefunctionLUT  # reference code
    from vmc.energy.eloc import _simple  # reference code

    d = torch.load(os.path.join(REFERENCE, "example/Fe2S2/fe2s2-OO.pth"), weights_only=False)
    h1e, h2e = d["h1e"], d["h2e"]
    ci = d["ci_space"].numpy()
    sorb, noA, noB, nele = int(d["sorb"]), int(d["noa"]), int(d["nob"]), int(d["nele"])
    first, n = 3000, 8
    x = t(ci[first : first + n].copy())
    out = dict(first=first, n=n)
    for tag, cplx in (("real", False), ("complex", True)):
        dtype = torch.complex128 if cplx else torch.double

        def ansatz(states):
            return toy_amplitude(states, sorb, cplx)

        def batcher(x, func):
            return func(ref.onv_to_tensor(x, sorb)).to(dtype)

        eloc, _, psi_x, _ = _simple(x, h1e, h2e, ansatz, batcher, sorb, nele, noA, noB, dtype=dtype, WF_LUT=None, use_unique=True)
        out[f"eloc_{tag}"] = eloc.numpy()
        out[f"psi_x_{tag}"] = psi_x.numpy()
        psi_tab = S.random_psi(ci.shape[0] // 2, seed=43, complex_=cplx)
        lut = WavefunctionLUT(t(ci[::2].copy()), t(psi_tab).to(dtype), sorb, "cpu")
        eloc, _, psi_x, _ = _simple(x, h1e, h2e, ansatz, batcher, sorb, nele, noA, noB, dtype=dtype, WF_LUT=lut, use_unique=True)
        out[f"eloc_lut_{tag}"] = eloc.numpy()
        out[f"psi_x_lut_{tag}"] = psi_x.numpy()
    save("simple_fe2s2", **out)


def main_reduce_sample():
    """Stochastic / semi-stochastic REDUCE goldens (vmc.energy.eloc._reduce_psi of the reference, eps_sample > 0) on the Fe2S2
    integrals, toy ansatz.  torch.multinomial is wrapped so that the draws it returned are stored with the golden: the
    reference result is a deterministic function of (inputs, draws), which is what the GPU test reproduces.
    Default dtype = double, as every shipped input sets it (main.py:30, example/Fe2S2/Fe2S2-OO-dcut-20.py:29): with
    float32 the reference's `_count / eps_sample` (eloc.py:276) is rounded to single precision."""
    torch.set_num_threads(os.cpu_count())
    torch.set_default_dtype(torch.double)
    ref = load_ref(1)
    libs = types.ModuleType("libs")
    libs.__path__ = []
    sys.modules["libs"] = libs
    sys.modules["libs.C_extension"] = ref
    libs.C_extension = ref
    sys.path.insert(0, REFERENCE)
    from vmc.energy.eloc import _reduce_psi  # reference code

    d = torch.load(os.path.join(REFERENCE, "example/Fe2S2/fe2s2-OO.pth"), weights_only=False)
    h1e, h2e = d["h1e"], d["h2e"]
    ci = d["ci_space"].numpy()
    sorb, noA, noB, nele = int(d["sorb"]), int(d["noa"]), int(d["nob"]), int(d["nele"])
    first, n = 2000, 24
    x = t(ci[first : first + n].copy())
    out = dict(first=first, n=n)
    real_multinomial = torch.multinomial
    for mode, eps, ns in (("semi", 1.0e-2, 1000), ("pure", 0.0, 500)):
        torch.manual_seed(2024)
        kept = {}

        def recording(prob, num, replacement=False, **kw):
            r = real_multinomial(prob, num, replacement=replacement, **kw)
            kept["draws"] = r.clone()
            return r

        for tag, cplx in (("real", False), ("complex", True)):
            dtype = torch.complex128 if cplx else torch.double

            def ansatz(states):
                return toy_amplitude(states, sorb, cplx)

            def batcher(x, func):
                return func(ref.onv_to_tensor(x, sorb)).to(dtype)

            if "draws" in kept:  # same draws for the complex run
                fixed = kept["draws"]
                torch.multinomial = lambda prob, num, replacement=False, **kw: fixed.clone()
            else:
                torch.multinomial = recording
            try:
                eloc, _, psi_x, _ = _reduce_psi(x, h1e, h2e, ansatz, batcher, sorb, nele, noA, noB, dtype=dtype, WF_LUT=None,
                                                use_unique=True, eps=eps, eps_sample=ns)
            finally:
                torch.multinomial = real_multinomial
            out[f"{mode}_eloc_{tag}"] = eloc.numpy()
            out[f"{mode}_psi_x_{tag}"] = psi_x.numpy()
        out[f"{mode}_draws"] = kept["draws"].numpy().astype(np.int16)
        out[f"{mode}_eps"] = eps
        out[f"{mode}_eps_sample"] = ns
    save("reduce_sample_fe2s2", **out)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "reduce_sample":
        main_reduce_sample()
    elif len(sys.argv) > 1 and sys.argv[1] == "reduce":
        main_reduce()
    elif len(sys.argv) > 1 and sys.argv[1] == "simple":
        main_simple()
    else:
        main()
        main_reduce()
        main_simple()
        main_reduce_sample()
