"""GPU-vs-GPU: this library against the UNMODIFIED reference CUDA extension (cpp_src/compile.sh -s GPU recipe,
cross-compiled for sm_100a by oracle/build_ref.py into the git-ignored oracle/_ref/C_extension_cuda_L{1,2}.so)
on the same device tensors -- SURVEY.md section 8(c) calls it the literal parity target.  Bit-exact for
determinant lists, lookup indices, masks and conversions; H_ij bit-exact (same order of additions)."""
import numpy as np
import pytest
import torch

from pynqs_b200 import C_extension as ops
from pynqs_b200 import synthetic as S

from util import fe2s2

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def ref_cuda(L):
    from oracle import build_ref

    if not build_ref.cuda_available(L):
        pytest.skip(f"oracle/_ref/C_extension_cuda_L{L}.so missing: run `python oracle/build_ref.py` where /root/reference is mounted")
    from pynqs_b200 import _lib

    _lib.load()
    assert torch.cuda.is_available()
    return build_ref.load_ref(L, cuda=True)


def test_fe2s2_fused_bit_identical_to_reference_cuda():
    """get_comb_hij_fused on 2048 determinants of the reference's Fe2S2 ci_space, real integrals (config 2)."""
    ref = ref_cuda(1)
    f = fe2s2()
    x, h1e, h2e = dev(f["ci"][:2048].copy()), dev(f["h1e"]), dev(f["h2e"])
    want_c, want_h = ref.get_comb_hij_fused(x, h1e, h2e, f["sorb"], f["nele"], f["noA"], f["noB"])
    for prepared in (None, False):
        comb, hmat = ops.get_comb_hij_fused(x, h1e, h2e, f["sorb"], f["nele"], f["noA"], f["noB"], prepared=prepared)
        assert torch.equal(comb, want_c)
        assert torch.equal(hmat.view(torch.int64), want_h.view(torch.int64))  # bit-exact, NaN-proof


@pytest.mark.parametrize("L,sorb,noA,noB,n", [(1, 40, 15, 15, 512), (1, 52, 5, 5, 256), (1, 12, 3, 3, 300), (2, 100, 25, 25, 3), (2, 100, 4, 3, 200)])
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_operators_bit_identical_to_reference_cuda(L, sorb, noA, noB, n, dtype):
    ref = ref_cuda(L)
    nele = noA + noB
    bra = S.random_onvs(n, sorb, noA, noB, seed=91)
    h1e, h2e = S.random_packed_integrals(sorb, seed=92, symmetric=False, dtype=dtype)
    x, d1, d2 = dev(bra), dev(h1e), dev(h2e)
    try:
        ref.check_sorb(sorb, nele)
    except Exception as e:  # the reference's own limits (MAX_NV)
        pytest.skip(f"reference rejects this geometry: {e}")
    want_c, want_h = ref.get_comb_hij_fused(x, d1, d2, sorb, nele, noA, noB)
    comb, hmat = ops.get_comb_hij_fused(x, d1, d2, sorb, nele, noA, noB)
    assert torch.equal(comb, want_c)
    ity = torch.int64 if dtype == np.float64 else torch.int32
    assert torch.equal(hmat.view(ity), want_h.view(ity))
    # get_comb_tensor / get_hij_torch (3-D and 2-D)
    assert torch.equal(ops.get_comb_tensor(x, sorb, nele, noA, noB, False)[0], ref.get_comb_tensor(x, sorb, nele, noA, noB, False)[0])
    assert torch.equal(ops.get_hij_torch(x, comb, d1, d2, sorb, nele).view(ity), ref.get_hij_torch(x, want_c, d1, d2, sorb, nele).view(ity))
    m = min(n, 128)
    assert torch.equal(ops.get_hij_torch(x[:m], x[:m].contiguous(), d1, d2, sorb, nele).view(ity),
                       ref.get_hij_torch(x[:m], x[:m].contiguous(), d1, d2, sorb, nele).view(ity))
    # conversions
    st = ref.onv_to_tensor(x, sorb)
    assert torch.equal(ops.onv_to_tensor(x, sorb), st)
    occ = ((st + 1) / 2).to(torch.uint8)
    assert torch.equal(ops.tensor_to_onv(occ, sorb), ref.tensor_to_onv(occ, sorb))
    # lookup of every connected determinant of the first samples in a table made of the samples themselves
    order = ops.sort_table(x, None, sorb)[0]
    q = comb[: min(n, 64)].reshape(-1, comb.size(2)).contiguous()
    want_i, want_m = ref.wavefunction_lut(order, q, sorb)
    got_i, got_m = ops.wavefunction_lut(order, q, sorb)
    assert torch.equal(got_m, want_m)
    assert torch.equal(got_i[got_m], want_i[want_m])  # the reference leaves idx of misses unspecified (-1 here too)
    assert torch.equal(got_i, want_i)
