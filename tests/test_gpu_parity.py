"""Parity of the CUDA path (through the C ABI, via pynqs_b200.C_extension) against
  * the reference goldens in tests/golden/ (outputs of the unmodified reference), and
  * the CPU oracle on the same seeded inputs.
Bars: determinant lists / lookup indices bit-exact; H_ij bit-exact where the summation order is
the reference's (asserted via SHA-256 of the raw bytes), and never worse than 1e-12 relative;
E_loc 1e-12 relative; mean energy 1e-10 Ha."""
import numpy as np
import pytest
import torch

from oracle import oracle as O
from pynqs_b200 import C_extension as ops
from pynqs_b200 import synthetic as S
from pynqs_b200.energy import local_energy_sample_space, local_energy_three_call
from pynqs_b200.lut import WavefunctionLUT, sort_onv

from util import OPS_CASES, fe2s2, load, ops_inputs, sha

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


@pytest.fixture(scope="module", autouse=True)
def _library_loaded():
    from pynqs_b200 import _lib

    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device (run with -m gpu on the B200 box)")
    _lib.load()  # fails loudly if libpynqs_b200.so is missing: there is no fallback


# ---- operators against the reference goldens -----------------------------------------------------
@pytest.mark.parametrize("mode", ["prepared", "plain"])
@pytest.mark.parametrize("name", OPS_CASES)
def test_comb_hij_fused_matches_reference(name, mode):
    """Both kernels behind get_comb_hij_fused: the table-driven one on the prepared integrals
    (production) and the plain one on the packed arrays."""
    c = ops_inputs(name)
    g = c["g"]
    h2e = dev(c["h2e"])
    prepared = ops.PreparedIntegrals(h2e, c["sorb"]) if mode == "prepared" else False
    comb, hmat = ops.get_comb_hij_fused(dev(c["bra"]), dev(c["h1e"]), h2e, c["sorb"], c["nele"], c["noA"], c["noB"], prepared=prepared)
    comb, hmat = comb.cpu().numpy(), hmat.cpu().numpy()
    k = g["comb_rows"].shape[0]
    np.testing.assert_array_equal(comb[:k, :: c["stride"]], g["comb_rows"])
    np.testing.assert_allclose(hmat[:k, :: c["stride"]], g["hmat_rows"], rtol=1e-12 if hmat.dtype == np.float64 else 1e-6, atol=0)
    assert sha(comb) == str(g["comb_sha"])      # connected determinants bit-exact
    assert sha(hmat) == str(g["hmat_sha"])      # H_ij bit-exact


@pytest.mark.parametrize("name", OPS_CASES)
def test_get_comb_tensor_matches_reference(name):
    c = ops_inputs(name)
    comb, second = ops.get_comb_tensor(dev(c["bra"]), c["sorb"], c["nele"], c["noA"], c["noB"], False)
    assert sha(comb.cpu().numpy()) == str(c["g"]["comb_sha"])
    assert second.device.type == "cpu" and second.dtype == torch.float64 and second.tolist() == [1.0]


def test_get_comb_tensor_flag_bit_states():
    c = ops_inputs("ops_odd_14sorb_4a2b")
    comb, states = ops.get_comb_tensor(dev(c["bra"][:7]), c["sorb"], c["nele"], c["noA"], c["noB"], True)
    want = O.onv_to_tensor(comb.cpu().numpy().reshape(-1, 8), c["sorb"]).reshape(7, -1, c["sorb"])
    np.testing.assert_array_equal(states.cpu().numpy(), want)


@pytest.mark.parametrize("name", OPS_CASES)
def test_get_hij_torch_3d_equals_fused_and_2d_matches_reference(name):
    c = ops_inputs(name)
    g = c["g"]
    bra, h1e, h2e = dev(c["bra"]), dev(c["h1e"]), dev(c["h2e"])
    comb, hmat = ops.get_comb_hij_fused(bra, h1e, h2e, c["sorb"], c["nele"], c["noA"], c["noB"])
    h3 = ops.get_hij_torch(bra, comb, h1e, h2e, c["sorb"], c["nele"])
    assert torch.equal(h3, hmat)
    m2 = g["hij2d"].shape[0]
    h2 = ops.get_hij_torch(bra[:m2], bra[:m2].clone(), h1e, h2e, c["sorb"], c["nele"])
    np.testing.assert_array_equal(h2.cpu().numpy(), g["hij2d"])


def test_fe2s2_real_integrals_match_reference():
    f = fe2s2()
    g = load("ops_c2_fe2s2")
    n = int(g["n"])
    comb, hmat = ops.get_comb_hij_fused(dev(f["ci"][:n]), dev(f["h1e"]), dev(f["h2e"]), f["sorb"], f["nele"], f["noA"], f["noB"])
    assert tuple(comb.shape) == (n, 7876, 8)
    np.testing.assert_array_equal(hmat[:4].cpu().numpy(), g["hmat_rows"])
    assert sha(comb.cpu().numpy()) == str(g["comb_sha"]) and sha(hmat.cpu().numpy()) == str(g["hmat_sha"])


def test_pyi_known_answers_on_device():
    st = torch.tensor([1, 1, 1, 1, 0, 0, 0, 0], dtype=torch.uint8, device=DEV)
    assert ops.tensor_to_onv(st, 8).cpu().tolist() == [[0b1111, 0, 0, 0, 0, 0, 0, 0]]
    onv = torch.tensor([0b1111, 0, 0, 0, 0, 0, 0, 0], dtype=torch.uint8, device=DEV)
    old = torch.get_default_dtype()
    try:
        torch.set_default_dtype(torch.float64)
        out = ops.onv_to_tensor(onv, 8)
        assert out.dtype == torch.float64 and out.cpu().tolist() == [[1, 1, 1, 1, -1, -1, -1, -1]]
        torch.set_default_dtype(torch.float32)
        assert ops.onv_to_tensor(onv, 8).dtype == torch.float32  # follows the default dtype (cpu_tensor.cpp:55)
    finally:
        torch.set_default_dtype(old)
    bra = torch.tensor([[0b1100, 0, 0, 0, 0, 0, 0, 0]], dtype=torch.uint8, device=DEV)
    comb, states = ops.get_comb_tensor(bra, 4, 2, 1, 1, True)
    assert comb[0, :, 0].cpu().tolist() == [12, 9, 6, 3]
    assert states[0].cpu().tolist() == [[-1, -1, 1, 1], [1, -1, -1, 1], [-1, 1, 1, -1], [1, 1, -1, -1]]
    key = np.zeros((6, 8), dtype=np.uint8)
    key[:, :2] = [[3, 0], [6, 0], [12, 0], [9, 1], [9, 2], [1, 3]]
    q = np.zeros((6, 8), dtype=np.uint8)
    q[:, :2] = [[12, 0], [9, 2], [6, 0], [3, 0], [14, 0], [1, 3]]
    idx, mask = ops.wavefunction_lut(dev(key), dev(q), 4)
    assert idx.cpu().tolist() == [2, 4, 1, 0, -1, 5] and mask.cpu().tolist() == [True, True, True, True, False, True]
    assert idx.dtype == torch.int64 and mask.dtype == torch.bool


def test_conversions_roundtrip_multiword():
    for sorb in (12, 64, 100, 132, 192):
        rng = np.random.default_rng(sorb)
        st = (rng.random((257, sorb)) < 0.4).astype(np.uint8)
        onv = ops.tensor_to_onv(dev(st), sorb)
        np.testing.assert_array_equal(onv.cpu().numpy(), O.tensor_to_onv(st, sorb))
        old = torch.get_default_dtype()
        torch.set_default_dtype(torch.float64)
        try:
            z = ops.onv_to_tensor(onv, sorb).cpu().numpy()
        finally:
            torch.set_default_dtype(old)
        np.testing.assert_array_equal(z, 2.0 * st - 1.0)


# ---- lookup ----------------------------------------------------------------------------------------
@pytest.mark.parametrize("L,sorb,na", [(1, 40, 15), (2, 100, 25), (3, 132, 3)])
def test_lookup_classic_and_hashed_equal_oracle(L, sorb, na):
    keys = S.random_onvs(20000 if L < 3 else 3000, sorb, na, na, seed=60 + L)
    order = O.sort_onv(keys)
    skeys = keys[order]
    rng = np.random.default_rng(61)
    q = np.concatenate([keys[rng.permutation(len(keys))[:5000]], S.random_onvs(5000, sorb, na, na, seed=62 + L)])
    want_idx, want_mask = O.lut(skeys, q)
    dk, dq = dev(skeys), dev(q)
    idx, mask = ops.wavefunction_lut(dk, dq, sorb)                                  # classic search
    np.testing.assert_array_equal(idx.cpu().numpy(), want_idx)
    np.testing.assert_array_equal(mask.cpu().numpy(), want_mask)
    idx2, mask2 = ops.wavefunction_lut(dk, dq, sorb, hash_index=ops.HashIndex(dk))  # hash index
    np.testing.assert_array_equal(idx2.cpu().numpy(), want_idx)
    np.testing.assert_array_equal(mask2.cpu().numpy(), want_mask)
    # device sort == reference order
    np.testing.assert_array_equal(sort_onv(dev(keys)).cpu().numpy(), order)


@pytest.mark.parametrize("n_q,shift", [(10003, 0), (10002, 0), (9001, 1), (3, 0)])
def test_lookup_hashed_tails_and_unaligned_queries(n_q, shift):
    """The four-queries-per-thread kernel takes n - n % 4 queries of a one-word lookup, the one-query kernel the rest, and all
    of them when the query tensor is not 16-byte aligned (a view that starts one row into its storage).  Queries in runs
    that share a beta or an alpha string (rows of comb) and random ones, against the oracle."""
    sorb, na = 40, 15
    keys = S.random_onvs(30000, sorb, na, na, seed=90)
    skeys = keys[O.sort_onv(keys)]
    comb, _ = ops.get_comb_tensor(dev(keys[:2]), sorb, 2 * na, na, na, False)     # runs of shared strings; row 0 of each is a hit
    rng = np.random.default_rng(91)
    pool = np.concatenate([comb.cpu().numpy().reshape(-1, 8), keys[rng.permutation(len(keys))[:3000]], S.random_onvs(3000, sorb, na, na, seed=92)])
    q = np.ascontiguousarray(pool[shift : shift + n_q])
    assert q.shape[0] == n_q
    want_idx, want_mask = O.lut(skeys, q)
    dk = dev(skeys)
    dq = dev(pool)[shift : shift + n_q]           # shift = 1: data pointer 8 bytes into the storage
    assert dq.is_contiguous() and (dq.data_ptr() % 16 == 8) == (shift == 1)
    idx, mask = ops.wavefunction_lut(dk, dq, sorb, hash_index=ops.HashIndex(dk))
    np.testing.assert_array_equal(idx.cpu().numpy(), want_idx)
    np.testing.assert_array_equal(mask.cpu().numpy(), want_mask)
    assert int(want_mask.sum()) >= 1


def test_lookup_golden_through_lut_mirror():
    g = load("lut_lookup_l1")
    rng = np.random.default_rng(31)
    keys = S.random_onvs(5000, 40, 15, 15, seed=32)
    psi = S.random_psi(5000, seed=33)
    lut = WavefunctionLUT(dev(keys), dev(psi), 40, DEV, rank=0, world_size=1)
    q = np.concatenate([keys[rng.permutation(5000)[:700]], S.random_onvs(700, 40, 15, 15, seed=34)])
    hit, miss, val = lut.lookup(dev(q))
    np.testing.assert_array_equal(hit.cpu().numpy(), g["hit"])
    np.testing.assert_array_equal(miss.cpu().numpy(), g["miss"])
    np.testing.assert_array_equal(val.cpu().numpy(), g["val"])


def test_lookup_with_duplicate_keys_falls_back_to_classic_probe_sequence():
    keys = S.random_onvs(1000, 40, 15, 15, seed=70)
    keys = np.concatenate([keys, keys[:300], keys[:100]])
    skeys = keys[O.sort_onv(keys)]
    q = np.concatenate([keys[:500], S.random_onvs(200, 40, 15, 15, seed=71)])
    want_idx, _ = O.lut(skeys, q)
    dk = dev(skeys)
    idx, _ = ops.wavefunction_lut(dk, dev(q), 40, hash_index=ops.HashIndex(dk))
    np.testing.assert_array_equal(idx.cpu().numpy(), want_idx)   # same element of each duplicate run as the reference


def test_empty_and_tiny_inputs():
    e = torch.empty((0, 8), dtype=torch.uint8, device=DEV)
    h1e, h2e = (dev(a) for a in S.random_packed_integrals(12, seed=1, symmetric=False))
    comb, hmat = ops.get_comb_hij_fused(e, h1e, h2e, 12, 6, 3, 3)
    assert tuple(comb.shape) == (0, 118, 8) and tuple(hmat.shape) == (0, 118)
    assert tuple(ops.get_comb_tensor(e, 12, 6, 3, 3)[0].shape) == (0, 118, 8)
    keys = dev(S.random_onvs(10, 12, 3, 3, seed=2))
    idx, mask = ops.wavefunction_lut(keys, e, 12)
    assert idx.numel() == 0 and mask.numel() == 0
    assert tuple(ops.get_hij_torch(e, keys, h1e, h2e, 12, 6).shape) == (0, 10)
    assert tuple(ops.tensor_to_onv(torch.empty((0, 12), dtype=torch.uint8, device=DEV), 12).shape) == (0, 8)
    # a table with no keys: everything is a miss
    idx, mask = ops.wavefunction_lut(e, keys, 12)
    assert (idx == -1).all() and not mask.any()
    idx, mask = ops.wavefunction_lut(e, keys, 12, hash_index=ops.HashIndex(e))
    assert (idx == -1).all() and not mask.any()
    with pytest.raises(ValueError):
        ops.get_comb_tensor(keys, 100, 6, 3, 3)          # width does not match sorb
    with pytest.raises(RuntimeError):
        ops.get_comb_tensor(keys[:, :4].T, 12, 6, 3, 3)  # non-contiguous


# ---- E_loc ------------------------------------------------------------------------------------------
@pytest.fixture(autouse=True)
def _default_tuning():
    from pynqs_b200 import _lib

    _lib.set_tuning()
    yield
    _lib.set_tuning()


@pytest.fixture(params=["folded", "full_keys", "block", "block_whole_tiles", "block_4_parts", "full_keys_tiles"])
def scan_route(request):
    """One-word ONVs are scanned through folded 32-bit strings; the knob full_keys forces the full-key
    route that multi-word ONVs (and tables of 2^30 keys or more) take; "block" sends every call, however small, through
    the grouping pass and the block kernel (samples sharing a beta string walk their groups together), with tiles from
    4 samples on.  All must give the same numbers."""
    from pynqs_b200 import _lib

    if request.param == "full_keys":
        _lib.set_tuning("full_keys", 1)
    elif request.param == "full_keys_tiles":
        _lib.set_tuning("full_keys", 1)
        _lib.set_tuning("eval_tiles", 2)
    elif request.param.startswith("block"):
        _lib.set_tuning("block_min_samples", 1)
        _lib.set_tuning("block_min_group", 4)
        _lib.set_tuning("eval_tiles", 2)  # ... and the evaluation kernel that takes 32 samples per warp
        # the alpha-beta groups of a tile walked by one warp, or split over 4 ("block": the kernel's own choice, 2 for small calls)
        if request.param != "block":
            _lib.set_tuning("block_parts", 1 if request.param == "block_whole_tiles" else 4)
    return request.param


@pytest.mark.parametrize("tag,cplx", [("real", False), ("complex", True)])
def test_eloc_matches_reference_python(tag, cplx, scan_route):
    f = fe2s2()
    g = load(f"eloc_fe2s2_{tag}")
    psi = S.random_psi(f["ci"].shape[0], seed=int(g["psi_seed"]), complex_=cplx)
    lut = WavefunctionLUT(dev(f["ci"]), dev(psi), f["sorb"], DEV, rank=0, world_size=1)
    first, n = int(g["first"]), int(g["n"])
    x = dev(f["ci"][first : first + n])
    h1e, h2e = dev(f["h1e"]), dev(f["h2e"])
    dtype = torch.complex128 if cplx else torch.double
    eloc, sloc, psi_x = local_energy_sample_space(x, h1e, h2e, lut, f["sorb"], f["nele"], f["noA"], f["noB"], dtype)
    np.testing.assert_allclose(eloc.cpu().numpy(), g["eloc"], rtol=1e-12, atol=0)
    np.testing.assert_array_equal(psi_x.cpu().numpy(), g["psi_x"])
    assert not sloc.any()
    eloc3, _, psi3 = local_energy_three_call(x, h1e, h2e, lut, f["sorb"], f["nele"], f["noA"], f["noB"], dtype, batch=24)
    np.testing.assert_allclose(eloc3.cpu().numpy(), g["eloc"], rtol=1e-12, atol=0)
    np.testing.assert_array_equal(psi3.cpu().numpy(), g["psi_x"])


def test_eloc_full_space_and_rayleigh_quotient(scan_route):
    g = load("eloc_c1_fullspace")
    keys = S.random_onvs(400, 12, 3, 3, seed=51)
    h1e, h2e = S.random_packed_integrals(12, seed=52, symmetric=True)
    psi = S.random_psi(400, seed=53)
    lut = WavefunctionLUT(dev(keys), dev(psi), 12, DEV, rank=0, world_size=1)
    eloc, _, psi_x = local_energy_sample_space(dev(keys), dev(h1e), dev(h2e), lut, 12, 6, 3, 3)
    np.testing.assert_allclose(eloc.cpu().numpy(), g["eloc"], rtol=1e-12, atol=1e-13)
    np.testing.assert_array_equal(psi_x.cpu().numpy(), psi)
    H = ops.get_hij_torch(dev(keys), dev(keys), dev(h1e), dev(h2e), 12, 6).cpu().numpy()
    np.testing.assert_array_equal(H, H.T)
    e_dense = psi @ H @ psi / (psi @ psi)
    e_vmc = float(np.sum(psi * psi * eloc.cpu().numpy()) / (psi @ psi))
    assert abs(e_dense - e_vmc) < 1e-10


@pytest.mark.parametrize("route", ["default", "block"])
def test_eloc_sample_missing_from_table_gives_nan_like_reference(route):
    if route == "block":
        from pynqs_b200 import _lib

        _lib.set_tuning("block_min_samples", 1)
        _lib.set_tuning("block_min_group", 4)
        _lib.set_tuning("eval_tiles", 2)
    keys = S.random_onvs(300, 12, 3, 3, seed=80)
    h1e, h2e = S.random_packed_integrals(12, seed=81, symmetric=True)
    lut = WavefunctionLUT(dev(keys[:200]), dev(S.random_psi(200, seed=82)), 12, DEV, rank=0, world_size=1)
    eloc, _, psi_x = local_energy_sample_space(dev(keys[190:210]), dev(h1e), dev(h2e), lut, 12, 6, 3, 3)
    e = eloc.cpu().numpy()
    assert np.isfinite(e[:10]).all() and np.isnan(e[10:]).all()       # psi(x) = 0 -> 0/0
    assert (psi_x[10:] == 0).all()


@pytest.mark.parametrize("search_factor", ["64", "8"])
def test_eloc_dense_table_full_space(scan_route, search_factor):
    """Full 24-spin-orbital space (6a6b, 853 776 keys): EVERY connected determinant is in the table
    (1819 hits per sample) and the groups are large enough (924 keys) for the alpha-beta groups to be
    searched instead of walked when the threshold is lowered (factor 8) -- the result must equal the oracle and
    the three-call path either way."""
    from pynqs_b200 import _lib

    _lib.set_tuning("search_factor", int(search_factor))
    sorb, noA, noB, nele = 24, 6, 6, 12
    import itertools

    strings = [sum(1 << (2 * k) for k in c) for c in itertools.combinations(range(12), 6)]
    a = np.array(strings, dtype=np.uint64)
    keys64 = (a[:, None] | (a[None, :] << np.uint64(1))).reshape(-1)          # alpha on even, beta on odd bits
    keys = keys64.view(np.uint8).reshape(-1, 8)
    assert keys.shape[0] == 853776
    rng = np.random.default_rng(5)
    keys = keys[rng.permutation(keys.shape[0])]
    psi = S.random_psi(keys.shape[0], seed=6)
    h1e, h2e = S.random_packed_integrals(sorb, seed=8, symmetric=True)
    lut = WavefunctionLUT(dev(keys), dev(psi), sorb, DEV, rank=0, world_size=1)
    x = dev(keys[:48])
    e1, _, p1 = local_energy_sample_space(x, dev(h1e), dev(h2e), lut, sorb, nele, noA, noB)
    e3, _, p3 = local_energy_three_call(x, dev(h1e), dev(h2e), lut, sorb, nele, noA, noB, batch=16)
    np.testing.assert_allclose(e1.cpu().numpy(), e3.cpu().numpy(), rtol=1e-12, atol=0)
    assert torch.equal(p1, p3)
    order = O.sort_onv(keys)
    want = O.eloc_sample_space(keys[:6], h1e, h2e, keys[order], psi[order], sorb, nele, noA, noB)
    np.testing.assert_allclose(e1[:6].cpu().numpy(), want, rtol=1e-12, atol=0)
    # every one of the 1819 rows must have been found
    comb, _ = ops.get_comb_tensor(x[:4], sorb, nele, noA, noB)
    _, mask = ops.wavefunction_lut(lut.bra_key, comb.view(-1, 8), sorb, hash_index=lut.hash_index)
    assert bool(mask.all())


def test_eloc_hit_queue_overflow_takes_the_full_route(scan_route):
    """Fe2S2 shape with ALL 15504 alpha strings on three beta strings: the own-beta group of a sample
    yields 75 + 1050 hits in one warp (> 512 queue slots), so the sample is redone by full enumeration
    + classic search -- same numbers as the oracle."""
    import itertools

    sorb, noA, noB, nele = 40, 15, 15, 30
    a = np.array([sum(1 << (2 * k) for k in c) for c in itertools.combinations(range(20), 15)], dtype=np.uint64)
    b = np.array([sum(1 << (2 * k + 1) for k in c) for c in (range(15), list(range(14)) + [17], list(range(13)) + [15, 19])], dtype=np.uint64)
    keys = (a[:, None] | b[None, :]).reshape(-1).view(np.uint8).reshape(-1, 8)
    keys = keys[np.random.default_rng(3).permutation(keys.shape[0])]
    psi = S.random_psi(keys.shape[0], seed=4)
    h1e, h2e = S.random_packed_integrals(sorb, seed=7, symmetric=True)
    lut = WavefunctionLUT(dev(keys), dev(psi), sorb, DEV, rank=0, world_size=1)
    x = keys[:40]
    e1, _, p1 = local_energy_sample_space(dev(x), dev(h1e), dev(h2e), lut, sorb, nele, noA, noB)
    order = O.sort_onv(keys)
    want = O.eloc_sample_space(x[:5], h1e, h2e, keys[order], psi[order], sorb, nele, noA, noB)
    np.testing.assert_allclose(e1[:5].cpu().numpy(), want, rtol=1e-12, atol=0)
    e3, _, p3 = local_energy_three_call(dev(x), dev(h1e), dev(h2e), lut, sorb, nele, noA, noB, batch=8)
    np.testing.assert_allclose(e1.cpu().numpy(), e3.cpu().numpy(), rtol=1e-12, atol=0)
    assert torch.equal(p1, p3)


@pytest.mark.parametrize("alpha_only", [True, False])
def test_eloc_one_huge_group_is_searched(alpha_only, scan_route):
    """36 spin orbitals, 9 electrons of ONE spin: the whole table (all 48 620 strings) is a single group,
    35x larger than the 1378 determinants connected to a sample -> with the search threshold lowered to 16x the
    own-string bucket is searched (binary search inside the bucket) instead of walked."""
    import itertools

    from pynqs_b200 import _lib

    _lib.set_tuning("search_factor", 16)

    sorb = 36
    noA, noB = (9, 0) if alpha_only else (0, 9)
    shift = 0 if alpha_only else 1
    keys64 = np.array([sum(1 << (2 * k + shift) for k in c) for c in itertools.combinations(range(18), 9)], dtype=np.uint64)
    keys = keys64.view(np.uint8).reshape(-1, 8)
    keys = keys[np.random.default_rng(9).permutation(keys.shape[0])]
    psi = S.random_psi(keys.shape[0], seed=10, complex_=alpha_only)
    h1e, h2e = S.random_packed_integrals(sorb, seed=11, symmetric=True)
    lut = WavefunctionLUT(dev(keys), dev(psi), sorb, DEV, rank=0, world_size=1)
    x = keys[:64]
    dtype = torch.complex128 if alpha_only else torch.double
    e1, _, p1 = local_energy_sample_space(dev(x), dev(h1e), dev(h2e), lut, sorb, 9, noA, noB, dtype)
    order = O.sort_onv(keys)
    want = O.eloc_sample_space(x[:8], h1e, h2e, keys[order], psi[order], sorb, 9, noA, noB)
    np.testing.assert_allclose(e1[:8].cpu().numpy(), want, rtol=1e-12, atol=0)
    e3, _, p3 = local_energy_three_call(dev(x), dev(h1e), dev(h2e), lut, sorb, 9, noA, noB, dtype, batch=32)
    np.testing.assert_allclose(e1.cpu().numpy(), e3.cpu().numpy(), rtol=1e-12, atol=0)
    assert torch.equal(p1, p3)


@pytest.mark.parametrize("nkeys", [40, 300, 3000])
def test_eloc_tiny_tables_many_groups_per_bucket(nkeys, scan_route):
    """Fe2S2 shape on tables of a few dozen to a few thousand keys: far fewer buckets than the 77 groups
    of a sample need, so groups of one sample share buckets (every bucket must still be walked once
    per sample) and unrelated strings share buckets with them (must be rejected on the full key)."""
    sorb, noA, noB = 40, 15, 15
    seeds = S.random_onvs(4, sorb, noA, noB, seed=91)
    comb = O.comb(seeds, sorb, noA, noB).reshape(-1, 8)
    rng = np.random.default_rng(92)
    keys = np.unique(np.concatenate([seeds, comb[rng.permutation(comb.shape[0])[: nkeys - 4]]]), axis=0)
    psi = S.random_psi(keys.shape[0], seed=93)
    h1e, h2e = S.random_packed_integrals(sorb, seed=7, symmetric=True)
    lut = WavefunctionLUT(dev(keys), dev(psi), sorb, DEV, rank=0, world_size=1)
    x = np.concatenate([seeds, keys[:12]])
    e1, _, p1 = local_energy_sample_space(dev(x), dev(h1e), dev(h2e), lut, sorb, noA + noB, noA, noB)
    order = O.sort_onv(keys)
    want = O.eloc_sample_space(x, h1e, h2e, keys[order], psi[order], sorb, noA + noB, noA, noB)
    np.testing.assert_allclose(e1.cpu().numpy(), want, rtol=1e-12, atol=0)


@pytest.mark.parametrize("route", ["default", "block"])
@pytest.mark.parametrize("sorb,noA,noB", [(40, 15, 15), (100, 3, 3)])
def test_eloc_table_with_duplicate_keys_takes_the_reference_route(sorb, noA, noB, route):
    """A table that holds some keys twice (the reference tolerates it: its binary search lands on one of the
    copies): the build notices, and every sample is evaluated by full enumeration + the reference's probe
    sequence, so even the choice among copies with DIFFERENT psi values is the reference's."""
    if route == "block":  # grouping pass + block kernel + tile evaluation see the duplicate flag too
        from pynqs_b200 import _lib

        _lib.set_tuning("block_min_samples", 1)
        _lib.set_tuning("block_min_group", 4)
        _lib.set_tuning("eval_tiles", 2)
    seeds = S.random_onvs(3, sorb, noA, noB, seed=61)
    comb = O.comb(seeds, sorb, noA, noB).reshape(-1, seeds.shape[1])
    rng = np.random.default_rng(62)
    keys = np.unique(np.concatenate([seeds, comb[rng.permutation(comb.shape[0])[:1500]]]), axis=0)
    keys = np.concatenate([keys, keys[::7], keys[::31]])          # duplicates, some three times
    psi = S.random_psi(keys.shape[0], seed=63)                     # different values on the copies
    h1e, h2e = S.random_packed_integrals(sorb, seed=7, symmetric=True)
    lut = WavefunctionLUT(dev(keys), dev(psi), sorb, DEV, rank=0, world_size=1)
    x = np.concatenate([seeds, keys[:9]])
    e1, _, p1 = local_energy_sample_space(dev(x), dev(h1e), dev(h2e), lut, sorb, noA + noB, noA, noB)
    order = O.sort_onv(keys)
    want = O.eloc_sample_space(x, h1e, h2e, keys[order], psi[order], sorb, noA + noB, noA, noB)
    np.testing.assert_allclose(e1.cpu().numpy(), want, rtol=1e-12, atol=0)
    e3, _, p3 = local_energy_three_call(dev(x), dev(h1e), dev(h2e), lut, sorb, noA + noB, noA, noB, batch=4)
    np.testing.assert_allclose(e1.cpu().numpy(), e3.cpu().numpy(), rtol=1e-12, atol=0)
    assert torch.equal(p1, p3)


def test_eloc_multiword_onvs_against_oracle():
    """L = 2 (100 spin orbitals) and L = 3 (132): one-pass E_loc vs the oracle on a small table."""
    for sorb, noA, noB, nkeys in ((100, 3, 3, 4000), (132, 3, 2, 3000)):
        # a table that contains real connections: a few seeds plus all their single/double excitations
        seeds = S.random_onvs(3, sorb, noA, noB, seed=sorb)
        comb = O.comb(seeds, sorb, noA, noB).reshape(-1, seeds.shape[1])
        rng = np.random.default_rng(1)
        pick = comb[rng.permutation(comb.shape[0])[: nkeys - 3]]
        keys = np.unique(np.concatenate([seeds, pick]), axis=0)
        psi = S.random_psi(keys.shape[0], seed=2, complex_=(sorb == 132))
        h1e, h2e = S.random_packed_integrals(sorb, seed=3, symmetric=False)
        lut = WavefunctionLUT(dev(keys), dev(psi), sorb, DEV, rank=0, world_size=1)
        x = np.concatenate([seeds, keys[:5]])
        dtype = torch.complex128 if sorb == 132 else torch.double
        e1, _, _ = local_energy_sample_space(dev(x), dev(h1e), dev(h2e), lut, sorb, noA + noB, noA, noB, dtype)
        order = O.sort_onv(keys)
        want = O.eloc_sample_space(x, h1e, h2e, keys[order], psi[order], sorb, noA + noB, noA, noB)
        np.testing.assert_allclose(e1.cpu().numpy(), want, rtol=1e-12, atol=0)


def test_eloc_h50_shape_many_splits():
    """Config-4 geometry: 100 spin orbitals, 25a25b, M = 571 876, 627 groups per sample split over many CTAs (L = 2)."""
    sorb, noA, noB = 100, 25, 25
    seeds = S.random_onvs(2, sorb, noA, noB, seed=44)
    comb = O.comb(seeds, sorb, noA, noB).reshape(-1, 16)
    rng = np.random.default_rng(45)
    keys = np.unique(np.concatenate([seeds, comb[rng.permutation(comb.shape[0])[:30000]]]), axis=0)
    psi = S.random_psi(keys.shape[0], seed=46)
    h1e, h2e = S.random_packed_integrals(sorb, seed=47, symmetric=False)
    lut = WavefunctionLUT(dev(keys), dev(psi), sorb, DEV, rank=0, world_size=1)
    e1, _, p1 = local_energy_sample_space(dev(seeds), dev(h1e), dev(h2e), lut, sorb, noA + noB, noA, noB)
    order = O.sort_onv(keys)
    want = O.eloc_sample_space(seeds, h1e, h2e, keys[order], psi[order], sorb, noA + noB, noA, noB)
    np.testing.assert_allclose(e1.cpu().numpy(), want, rtol=1e-12, atol=0)


# ---- larger, size-independent properties ---------------------------------------------------------------
def test_fe2s2_shape_at_scale_properties(scan_route):
    """Config-2 geometry (40 sorb, 15a15b, M = 7876) at 2e5 table keys / 8192 evaluated samples:
    fused == rederived, hash lookup == classic search, one-pass E_loc == three-call path,
    mean energy within 1e-10 Ha; oracle spot check."""
    sorb, noA, noB, nele = 40, 15, 15, 30
    N, n = 200_000, 8192
    keys = S.random_onvs(N, sorb, noA, noB, seed=1234)
    psi = S.random_psi(N, seed=1235)
    h1e_np, h2e_np = S.random_packed_integrals(sorb, seed=7, symmetric=True)
    h1e, h2e = dev(h1e_np), dev(h2e_np)
    lut = WavefunctionLUT(dev(keys), dev(psi), sorb, DEV, rank=0, world_size=1)
    x = dev(keys[:n])
    comb, hmat = ops.get_comb_hij_fused(x[:2048], h1e, h2e, sorb, nele, noA, noB)
    assert torch.equal(ops.get_hij_torch(x[:2048], comb, h1e, h2e, sorb, nele), hmat)
    flat = comb.view(-1, 8)
    i1, m1 = ops.wavefunction_lut(lut.bra_key, flat[: 4_000_000], sorb)                       # >= hash threshold
    from pynqs_b200 import _lib

    i2 = torch.empty_like(i1)
    m2 = torch.empty_like(m1)
    _lib.check(_lib.load().pynqs_lut(_lib.vp(lut.bra_key.data_ptr()), _lib.i64(N), _lib.vp(flat.data_ptr()), _lib.i64(4_000_000), 1,
                                     _lib.vp(i2.data_ptr()), _lib.vp(m2.data_ptr()), _lib.vp(torch.cuda.current_stream().cuda_stream)))
    assert torch.equal(i1, i2) and torch.equal(m1, m2)
    assert int(m1.sum()) >= 508  # at least the bra rows
    e1, _, p1 = local_energy_sample_space(x, h1e, h2e, lut, sorb, nele, noA, noB)
    e3, _, p3 = local_energy_three_call(x, h1e, h2e, lut, sorb, nele, noA, noB, batch=2048)
    np.testing.assert_allclose(e1.cpu().numpy(), e3.cpu().numpy(), rtol=1e-12, atol=0)
    assert torch.equal(p1, p3) and torch.equal(p1, dev(psi[:n]))
    prob = (p1 * p1) / (p1 * p1).sum()
    assert abs(float((prob * e1).sum() - (prob * e3).sum())) < 1e-10
    # oracle spot check of the one-pass kernel
    order = O.sort_onv(keys)
    want = O.eloc_sample_space(keys[:8], h1e_np, h2e_np, keys[order], psi[order], sorb, nele, noA, noB)
    np.testing.assert_allclose(e1[:8].cpu().numpy(), want, rtol=1e-12, atol=0)


@pytest.mark.parametrize("table", ["uniform", "zipf0.8"])
def test_headline_size_table_against_the_oracle(table):
    """The benchmarked configuration itself -- 10^6 unique Fe2S2 samples that are their own table, the reference's Fe2S2
    integrals -- and its Zipf-skewed variant (beta strings with thousands of samples next to strings with a handful): E_loc
    of ALL samples on the GPU, an oracle subsample spread over the table plus, for the skewed table, samples of the heaviest
    beta strings (the long groups, the long hit lists, the warp-per-sample evaluation route)."""
    import bench

    f = fe2s2()
    sorb, noA, noB, nele = int(f["sorb"]), int(f["noA"]), int(f["noB"]), int(f["nele"])
    keys = bench.make_table(table, 1_000_000)
    n = keys.shape[0]
    psi = S.random_psi(n, seed=1235)
    h1e_np, h2e_np = np.ascontiguousarray(f["h1e"]), np.ascontiguousarray(f["h2e"])
    lut = WavefunctionLUT(dev(keys), dev(psi), sorb, DEV, rank=0, world_size=1)
    eloc, _, psi_x = local_energy_sample_space(lut.bra_key, dev(h1e_np), dev(h2e_np), lut, sorb, nele, noA, noB)
    e = eloc.cpu().numpy()
    assert np.isfinite(e).all()
    assert torch.equal(psi_x, lut.wf_value)
    skeys, spsi = lut.bra_key.cpu().numpy(), lut.wf_value.cpu().numpy()
    words = skeys.view(np.uint64).reshape(-1)
    assert (words[1:] > words[:-1]).all()  # the table the oracle searches is the sorted unique set
    pick = list(range(0, n, n // 16))[:16]
    if table != "uniform":
        beta = words & np.uint64(0xAAAAAAAAAAAAAAAA)
        vals, counts = np.unique(beta, return_counts=True)
        for b in vals[np.argsort(counts)[-3:]]:  # the three heaviest beta strings: two samples of each
            pick += list(np.nonzero(beta == b)[0][[0, -1]])
        assert counts.max() > 2000
    pick = np.array(sorted(set(int(i) for i in pick)))
    want = O.eloc_sample_space(skeys[pick], h1e_np, h2e_np, skeys, spsi, sorb, nele, noA, noB)
    np.testing.assert_allclose(e[pick], want, rtol=1e-12, atol=0)


def test_row_index_beyond_int32():
    """n * M > 2^31 rows in one call (the reference's int indexing overflows here, kernel.cu:190,220)."""
    sorb, noA, noB, nele = 40, 15, 15, 30
    n = 273_000                      # 273000 * 7876 = 2.15e9 rows, 17.2 GB comb + 17.2 GB Hmat
    free, _ = torch.cuda.mem_get_info()
    if free < 48 * 2**30:
        pytest.skip("needs ~36 GB of free device memory")
    bra = S.random_onvs(4096, sorb, noA, noB, seed=90)
    big = np.tile(bra, (n // 4096 + 1, 1))[:n]
    h1e_np, h2e_np = S.random_packed_integrals(sorb, seed=7, symmetric=True)
    comb, hmat = ops.get_comb_hij_fused(dev(big), dev(h1e_np), dev(h2e_np), sorb, nele, noA, noB)
    want_c, want_h = O.comb_hij_fused(big[-2:], h1e_np, h2e_np, sorb, nele, noA, noB)
    np.testing.assert_array_equal(comb[-2:].cpu().numpy(), want_c)
    np.testing.assert_array_equal(hmat[-2:].cpu().numpy(), want_h)
    np.testing.assert_array_equal(comb[0].cpu().numpy(), O.comb(big[:1], sorb, noA, noB)[0])
    del comb, hmat
    torch.cuda.empty_cache()


def test_192_sorb_three_words_against_oracle():
    """Config-5 geometry: 192 spin orbitals (L = 3), 4a4b, M = 186393, packed h2e = 1.345 GB."""
    sorb, noA, noB = 192, 4, 4
    h1e_np, h2e_np = S.random_packed_integrals(sorb, seed=5, symmetric=False)
    assert h2e_np.size == 168_113_616
    bra = S.random_onvs(2, sorb, noA, noB, seed=6)
    want_c, want_h = O.comb_hij_fused(bra, h1e_np, h2e_np, sorb, 8, noA, noB)
    h2e = dev(h2e_np)
    for prepared in (False, ops.PreparedIntegrals(h2e, sorb)):   # prepared copy: ~1.07 GB at 192 spin orbitals
        comb, hmat = ops.get_comb_hij_fused(dev(bra), dev(h1e_np), h2e, sorb, 8, noA, noB, prepared=prepared)
        np.testing.assert_array_equal(comb.cpu().numpy(), want_c)
        np.testing.assert_array_equal(hmat.cpu().numpy(), want_h)
    # one-pass E_loc with 30a30b: 1982 groups per sample, M = 5.8 million -- the scan kernel's shared memory only
    # fits with a shortened chunk list (one chunk per group), the rest of the groups are walked one warp each
    noA = noB = 30
    seed = S.random_onvs(1, sorb, noA, noB, seed=8)
    comb = O.comb(seed, sorb, noA, noB).reshape(-1, 24)
    rng = np.random.default_rng(9)
    keys = np.unique(np.concatenate([seed, comb[rng.permutation(comb.shape[0])[:20000]]]), axis=0)
    del comb
    psi = S.random_psi(keys.shape[0], seed=10)
    lut = WavefunctionLUT(dev(keys), dev(psi), sorb, DEV, rank=0, world_size=1)
    x = np.concatenate([seed, keys[:3]])
    e1, _, _ = local_energy_sample_space(dev(x), dev(h1e_np), h2e, lut, sorb, 60, noA, noB)
    order = O.sort_onv(keys)
    want = O.eloc_sample_space(x[:2], h1e_np, h2e_np, keys[order], psi[order], sorb, 60, noA, noB)
    np.testing.assert_allclose(e1[:2].cpu().numpy(), want, rtol=1e-12, atol=0)


# ---- the steps either side of the kernels: table sort, energy moments -------------------------------
@pytest.mark.parametrize("L,sorb,na,cplx", [(1, 40, 15, False), (1, 64, 20, True), (2, 100, 25, True), (3, 132, 3, False), (3, 192, 40, True)])
def test_sort_table_matches_oracle_order(L, sorb, na, cplx):
    """sort_table on the significant bits == the reference's lexsort order (stable: duplicates stay
    in input order), keys / values / permutation all consistent."""
    k = S.random_onvs(3000, sorb, na, na, seed=70 + L)
    k = np.concatenate([k, k[:500]])[np.random.default_rng(6).permutation(3500)]
    psi = S.random_psi(3500, seed=71, complex_=cplx)
    order = O.sort_onv(k)
    for bits in (sorb, 0):
        sk, sp, perm = ops.sort_table(dev(k), dev(psi), bits)
        np.testing.assert_array_equal(perm.cpu().numpy(), order)
        np.testing.assert_array_equal(sk.cpu().numpy(), k[order])
        np.testing.assert_array_equal(sp.cpu().numpy(), psi[order])
    lut = WavefunctionLUT(dev(k[:3000]), dev(psi[:3000]), sorb, DEV, rank=0, world_size=1)
    o3 = O.sort_onv(k[:3000])
    np.testing.assert_array_equal(lut.bra_key.cpu().numpy(), k[:3000][o3])
    np.testing.assert_array_equal(lut.wf_value.cpu().numpy(), psi[:3000][o3])
    inv = np.empty(3000, dtype=np.int64)
    inv[o3] = np.arange(3000)
    np.testing.assert_array_equal(lut.idx_sorted.cpu().numpy(), inv)


def test_sort_table_one_million_keys_sorted_and_a_permutation():
    k = S.random_onvs(1_000_000, 40, 15, 15, seed=72)
    psi = S.random_psi(1_000_000, seed=73)
    sk, sp, perm = ops.sort_table(dev(k), dev(psi), 40)
    w = sk.view(torch.int64).view(-1)
    assert bool((w[1:] > w[:-1]).all())  # strictly ascending: sorted and unique
    assert torch.equal(torch.sort(perm).values, torch.arange(1_000_000, device=DEV))
    assert torch.equal(dev(k)[perm], sk) and torch.equal(dev(psi)[perm], sp)
    e, _, e_perm = ops.sort_table(dev(k[:0]), None, 40)
    assert e.shape == (0, 8) and e_perm.numel() == 0


@pytest.mark.parametrize("n", [1, 31, 1000, 300_001])
@pytest.mark.parametrize("cplx", [False, True])
def test_weighted_moments_and_statistics(n, cplx):
    from pynqs_b200.distributed import energy_statistics, energy_statistics_amplitudes

    rng = np.random.default_rng(80 + n)
    e = rng.standard_normal(n) - 116.6
    if cplx:
        e = e + 1e-3j * rng.standard_normal(n)
    amp = S.random_psi(n, seed=81, complex_=cplx)
    w = np.abs(amp) ** 2
    for weight, is_amp in ((dev(w), False), (dev(amp), True)):
        m = ops.weighted_moments(dev(e), weight, is_amp).cpu().numpy()
        d = e - e[0]
        want = [w.sum(), (w * d.real).sum(), (w * d.imag).sum(), (w * np.abs(d) ** 2).sum(), e[0].real, e[0].imag, n]
        np.testing.assert_allclose(m, want, rtol=1e-12, atol=1e-12 * w.sum())
        m2 = ops.weighted_moments(dev(e), weight, is_amp).cpu().numpy()
        np.testing.assert_array_equal(m, m2)  # deterministic, and the ticket was reset
    p = w / w.sum()
    mean = (p * e).sum()
    var = (p * np.abs(e - mean) ** 2).sum()
    for st in (energy_statistics(dev(e), dev(p)), energy_statistics_amplitudes(dev(e), dev(amp))):
        assert abs(st["mean"] - mean) < 1e-11 and abs(st["var"] - var) < 1e-11 * max(1.0, var) and st["n"] == n


@pytest.mark.parametrize("sorb,noA,noB", [(12, 3, 3), (40, 15, 15), (100, 2, 1), (132, 3, 2)])
def test_diagonal_from_the_shared_memory_table_is_bit_identical(sorb, noA, noB):
    """Large batches compute H_ii from a per-CTA table of <pq||pq> in shared memory, small ones gather from
    the packed array: same values, same order of additions -> identical bits; oracle spot check."""
    n = 6000 if sorb != 40 else 4500
    x = S.random_onvs(n, sorb, noA, noB, seed=50 + sorb) if sorb > 12 else np.repeat(S.random_onvs(300, sorb, noA, noB, seed=62), 20, axis=0)
    h1e, h2e = S.random_packed_integrals(sorb, seed=51, symmetric=False)
    dx, dh1, dh2 = dev(x), dev(h1e), dev(h2e)
    _, big = ops.get_comb_hij_fused(dx, dh1, dh2, sorb, noA + noB, noA, noB)
    diag_big = big[:, 0].clone()
    del big
    parts = [ops.get_comb_hij_fused(dx[i : i + 1000], dh1, dh2, sorb, noA + noB, noA, noB)[1][:, 0].clone() for i in range(0, n, 1000)]
    assert torch.equal(diag_big, torch.cat(parts))
    _, want = O.comb_hij_fused(x[:4], h1e, h2e, sorb, noA + noB, noA, noB)
    np.testing.assert_array_equal(diag_big[:4].cpu().numpy(), want[:, 0])


# ---- REDUCE method: compacted connected determinants --------------------------------------------------------
@pytest.mark.parametrize("tag,cplx", [("real", False), ("complex", True)])
def test_reduce_method_matches_reference_python(tag, cplx):
    """get_comb_hij_reduced == the reference's torch.where(|H| >= eps) set (digests: bit-exact indices, values,
    determinants) and local_energy_reduced == vmc.energy.eloc._reduce_psi with the toy ansatz, with and without LUT."""
    from pynqs_b200.energy import local_energy_reduced
    from util import toy_amplitude

    f = fe2s2()
    g = load("reduce_fe2s2")
    first, n, eps = int(g["first"]), int(g["n"]), float(g["eps"])
    sorb, nele, noA, noB = f["sorb"], f["nele"], f["noA"], f["noB"]
    x, h1e, h2e = dev(f["ci"][first : first + n]), dev(f["h1e"]), dev(f["h2e"])
    xk, hk, idx, offsets = ops.get_comb_hij_reduced(x, h1e, h2e, sorb, nele, noA, noB, eps)
    assert idx.numel() == int(g["K"]) and int(offsets[-1]) == int(g["K"])
    assert sha(idx.cpu().numpy()) == str(g["idx_sha"]) and sha(hk.cpu().numpy()) == str(g["hij_sha"]) and sha(xk.cpu().numpy()) == str(g["x_sha"])
    np.testing.assert_array_equal(torch.diff(offsets).cpu().numpy(), g["counts"])
    dtype = torch.complex128 if cplx else torch.double

    def ansatz(xq):
        return toy_amplitude(ops.onv_to_tensor(xq, sorb), sorb, cplx)

    eloc, _, psi_x = local_energy_reduced(x, h1e, h2e, ansatz, sorb, nele, noA, noB, dtype, eps=eps, batch=10)
    np.testing.assert_allclose(eloc.cpu().numpy(), g[f"eloc_{tag}"], rtol=1e-12, atol=0)
    np.testing.assert_allclose(psi_x.cpu().numpy(), g[f"psi_x_{tag}"], rtol=1e-14, atol=0)
    keys = f["ci"][::2]
    lut = WavefunctionLUT(dev(keys), dev(S.random_psi(keys.shape[0], seed=43, complex_=cplx)), sorb, DEV, rank=0, world_size=1)

    def with_lut(xq):  # the reference's Func(ansatz, x, WF_LUT): table values where found, ansatz elsewhere
        found, missing, value = lut.lookup(xq)
        out = torch.empty(xq.size(0), dtype=dtype, device=xq.device)
        out[found] = value
        out[missing] = ansatz(xq[missing]).to(dtype)
        return out

    eloc, _, psi_x = local_energy_reduced(x, h1e, h2e, with_lut, sorb, nele, noA, noB, dtype, eps=eps)
    np.testing.assert_allclose(eloc.cpu().numpy(), g[f"eloc_lut_{tag}"], rtol=1e-12, atol=0)
    np.testing.assert_allclose(psi_x.cpu().numpy(), g[f"psi_x_lut_{tag}"], rtol=1e-14, atol=0)


@pytest.mark.parametrize("name", OPS_CASES)
def test_reduced_rows_against_oracle_and_fused(name):
    """Every operator golden shape (L = 1, 2, 3; float32 and float64): the kept rows equal the oracle's at several
    thresholds, eps = 0 returns the whole fused output in order, a huge eps returns nothing."""
    c = ops_inputs(name)
    sorb, nele, noA, noB = c["sorb"], c["nele"], c["noA"], c["noB"]
    x, h1e, h2e = dev(c["bra"]), dev(c["h1e"]), dev(c["h2e"])
    comb, hmat = ops.get_comb_hij_fused(x, h1e, h2e, sorb, nele, noA, noB)
    xk, hk, idx, off = ops.get_comb_hij_reduced(x, h1e, h2e, sorb, nele, noA, noB, 0.0)
    assert torch.equal(xk, comb.view(-1, comb.size(2))) and torch.equal(hk, hmat.view(-1))
    assert torch.equal(idx, torch.arange(hmat.numel(), device=DEV))
    absh = np.abs(hmat.cpu().numpy().ravel())
    for q in (0.5, 0.97):
        eps = float(np.quantile(absh[absh > 0], q))
        xk, hk, idx, off = ops.get_comb_hij_reduced(x, h1e, h2e, sorb, nele, noA, noB, eps)
        wx, wh, wi, wo = O.reduced(c["bra"], c["h1e"], c["h2e"], sorb, nele, noA, noB, eps)
        np.testing.assert_array_equal(idx.cpu().numpy(), wi)
        np.testing.assert_array_equal(hk.cpu().numpy(), wh)
        np.testing.assert_array_equal(xk.cpu().numpy(), wx)
        np.testing.assert_array_equal(off.cpu().numpy(), wo)
    xk, hk, idx, off = ops.get_comb_hij_reduced(x, h1e, h2e, sorb, nele, noA, noB, 1e30)
    assert xk.shape == (0, comb.size(2)) and hk.numel() == 0 and not off.any()
    e0 = ops.get_comb_hij_reduced(x[:0], h1e, h2e, sorb, nele, noA, noB, 1e-3)
    assert e0[0].shape == (0, comb.size(2)) and e0[3].numel() == 1


def test_merge_rank_sample_matches_the_reference_loop():
    """merge_counts[idx[i]] += counts[i] (cpu_tensor.cpp:537-556) incl. repeated indices, empty input, length 0."""
    rng = np.random.default_rng(17)
    idx = rng.integers(0, 5000, size=200_000)
    cnt = rng.integers(1, 1000, size=200_000)
    want = np.zeros(5000, dtype=np.int64)
    np.add.at(want, idx, cnt)
    got = ops.merge_rank_sample(dev(idx), dev(cnt), torch.empty(0, device=DEV), 5000)
    assert got.dtype == torch.int64
    np.testing.assert_array_equal(got.cpu().numpy(), want)
    assert not ops.merge_rank_sample(dev(idx[:0]), dev(cnt[:0]), torch.empty(0, device=DEV), 7).any()
    assert ops.merge_rank_sample(dev(idx[:0]), dev(cnt[:0]), torch.empty(0, device=DEV), 0).numel() == 0


# ---- REDUCE method, stochastic / semi-stochastic branch (eloc.py:257-283) -------------------------------------------
@pytest.mark.parametrize("mode", ["semi", "pure"])
@pytest.mark.parametrize("tag,cplx", [("real", False), ("complex", True)])
def test_stochastic_reduce_with_the_references_draws_is_exact(mode, tag, cplx):
    """The reference's _reduce_psi with eps_sample > 0 is a deterministic function of (inputs, multinomial draws); the golden
    stores the draws torch.multinomial returned.  Same draws here -> same E_loc to 1e-12, without any [n, M] array."""
    from pynqs_b200.energy import local_energy_reduced
    from util import toy_amplitude

    f = fe2s2()
    g = load("reduce_sample_fe2s2")
    first, n = int(g["first"]), int(g["n"])
    eps, ns = float(g[f"{mode}_eps"]), int(g[f"{mode}_eps_sample"])
    x = dev(f["ci"][first : first + n].copy())
    draws = dev(g[f"{mode}_draws"].astype(np.int64))
    dtype = torch.complex128 if cplx else torch.double

    def psi_of(xk):
        return toy_amplitude(ops.onv_to_tensor(xk, f["sorb"]), f["sorb"], cplx).to(dtype)

    eloc, _, psi_x = local_energy_reduced(x, dev(f["h1e"]), dev(f["h2e"]), psi_of, f["sorb"], f["nele"], f["noA"], f["noB"], dtype=dtype,
                                          eps=eps, eps_sample=ns, draws=draws, batch=10)
    np.testing.assert_allclose(eloc.cpu().numpy(), g[f"{mode}_eloc_{tag}"], rtol=1e-12, atol=0)
    np.testing.assert_allclose(psi_x.cpu().numpy(), g[f"{mode}_psi_x_{tag}"], rtol=1e-14, atol=0)


def test_stochastic_reduce_rows_against_a_torch_restatement():
    """Kept / drawn rows, their order, flat indices and re-weighted elements against the reference's formulas evaluated with
    torch on the materialised Hmat of a few samples (counts from the same draws)."""
    c = ops_inputs("ops_odd_14sorb_4a2b")
    bra, h1e, h2e = dev(c["bra"][:9]), dev(c["h1e"]), dev(c["h2e"])
    comb, hmat = ops.get_comb_hij_fused(bra, h1e, h2e, c["sorb"], c["nele"], c["noA"], c["noB"])
    n, M = hmat.shape
    eps, ns = float(hmat.abs().median()), 64
    habs = hmat.abs()
    sub = torch.where(habs >= eps, torch.zeros_like(habs), habs)
    prob = sub / sub.sum(1, keepdim=True)
    torch.manual_seed(7)
    draws = torch.multinomial(prob, ns, replacement=True)
    x, hij, idx, off = ops.get_comb_hij_sampled(bra, h1e, h2e, c["sorb"], c["nele"], c["noA"], c["noB"], eps, ns, draws=draws)
    want_idx, want_h = [], []
    for s in range(n):
        kept = torch.where(habs[s] >= eps)[0]
        drawn, cnt = draws[s].unique(sorted=True, return_counts=True)
        want_idx.append(torch.cat([kept, drawn]) + s * M)
        want_h.append(torch.cat([hmat[s, kept], (cnt / ns) * hmat[s, drawn] / prob[s, drawn]]))
        assert int(off[s + 1] - off[s]) == kept.numel() + drawn.numel()
    want_idx, want_h = torch.cat(want_idx), torch.cat(want_h)
    assert torch.equal(idx, want_idx)
    np.testing.assert_allclose(hij.cpu().numpy(), want_h.cpu().numpy(), rtol=1e-13, atol=0)
    assert torch.equal(x, comb.reshape(-1, comb.size(2))[want_idx])


def test_stochastic_reduce_seeded_generator_is_unbiased_and_reproducible():
    """With the library's own generator (Philox keyed by the seed): same seed -> identical output; the estimator's mean over
    many seeds agrees with the exact local energy (all rows) within 5 standard errors."""
    from pynqs_b200.energy import local_energy_reduced
    from util import toy_amplitude

    f = fe2s2()
    g = load("simple_fe2s2")
    first, n = int(g["first"]), int(g["n"])
    x = dev(f["ci"][first : first + n].copy())
    h1e, h2e = dev(f["h1e"]), dev(f["h2e"])

    def psi_of(xk):
        return toy_amplitude(ops.onv_to_tensor(xk, f["sorb"]), f["sorb"], False)

    args = (x, h1e, h2e, psi_of, f["sorb"], f["nele"], f["noA"], f["noB"])
    a = local_energy_reduced(*args, eps=1e-2, eps_sample=1000, seed=11)[0]
    b = local_energy_reduced(*args, eps=1e-2, eps_sample=1000, seed=11)[0]
    assert torch.equal(a, b)
    draws = np.stack([local_energy_reduced(*args, eps=1e-2, eps_sample=1000, seed=100 + k)[0].cpu().numpy() for k in range(64)])
    mean, se = draws.mean(0), draws.std(0, ddof=1) / np.sqrt(draws.shape[0])
    assert np.all(se > 0)
    assert np.all(np.abs(mean - g["eloc_real"]) <= 5 * se + 1e-9), (mean - g["eloc_real"], se)


# ---- index glue of the lookup: compaction and unique-of-misses ------------------------------
@pytest.mark.parametrize("n,p_hit", [(0, 0.5), (1, 1.0), (2047, 0.3), (2048, 0.0), (2049, 1.0), (1_000_003, 0.004), (3_000_001, 0.6)])
@pytest.mark.parametrize("cplx", [False, True])
def test_lookup_compact_equals_the_references_index_ops(n, p_hit, cplx):
    g = torch.Generator(device=DEV).manual_seed(n + 17)
    N = 5000
    mask = torch.rand(n, generator=g, device=DEV) < p_hit
    idx = torch.where(mask, torch.randint(0, N, (n,), generator=g, device=DEV), torch.full((n,), -1, device=DEV))
    val = dev(S.random_psi(N, seed=3, complex_=cplx))
    hit, miss, value = ops.lookup_compact(idx, mask, val)
    baseline = torch.arange(n, device=DEV, dtype=torch.int64)  # utils/public_function.py:825-838
    assert torch.equal(hit, baseline[mask])
    assert torch.equal(miss, baseline[torch.logical_not(mask)])
    assert torch.equal(torch.view_as_real(value) if cplx else value, torch.view_as_real(val[idx.masked_select(mask)]) if cplx else val[idx.masked_select(mask)])


@pytest.mark.parametrize("L,sorb,na,n", [(1, 40, 15, 50_000), (1, 12, 3, 5000), (2, 100, 25, 3000), (3, 132, 3, 4097), (1, 40, 15, 1), (1, 40, 15, 0)])
def test_unique_onv_inverse_and_order(L, sorb, na, n):
    base = S.random_onvs(max(n // 3, 1), sorb, na, na, seed=60 + L)
    rng = np.random.default_rng(61)
    x = base[rng.integers(0, base.shape[0], size=n)] if n else base[:0]
    uniq, inv = ops.unique_onv(dev(x))
    assert torch.equal(uniq[inv], dev(x))
    want = np.unique(x, axis=0).shape[0] if n else 0
    assert uniq.size(0) == want
    if want > 1:  # strictly ascending as little-endian multi-word integers
        order = O.sort_onv(uniq.cpu().numpy())
        assert np.array_equal(order, np.arange(want))


def test_func_mirror_matches_the_references_func_semantics():
    """energy.Func (lookup + unique of the misses + ansatz) against a plain evaluation of every row."""
    from pynqs_b200.energy import Func
    from util import toy_amplitude

    f = fe2s2()
    sorb = f["sorb"]
    x = dev(f["ci"][:600].copy())
    comb = ops.get_comb_tensor(x[:3], sorb, f["nele"], f["noA"], f["noB"], False)[0].reshape(-1, 8)
    rows = torch.cat([comb, comb[:5000], x])  # duplicates among hits and among misses
    psi_tab = dev(S.random_psi(300, seed=44))
    lut = WavefunctionLUT(dev(f["ci"][:600:2].copy()), psi_tab, sorb, DEV, rank=0, world_size=1)
    calls = []

    def ansatz(xx):
        calls.append(xx.size(0))
        return toy_amplitude(ops.onv_to_tensor(xx, sorb), sorb, False)

    want = ansatz(rows)
    found, _, value = lut.lookup(rows)
    want[found] = value
    calls.clear()
    got = Func(ansatz, rows, lut, use_unique=True)
    # (the toy ansatz is a matmul: its last bit depends on the batch it is evaluated in)
    np.testing.assert_allclose(got.cpu().numpy(), want.cpu().numpy(), rtol=1e-13, atol=0)
    assert torch.equal(got[found], value)
    n_distinct_misses = np.unique(rows.cpu().numpy()[np.setdiff1d(np.arange(rows.size(0)), found.cpu().numpy())], axis=0).shape[0]
    assert calls == [n_distinct_misses]  # the ansatz saw every distinct missing determinant exactly once
    np.testing.assert_allclose(Func(ansatz, rows, None, use_unique=True).cpu().numpy(), ansatz(rows).cpu().numpy(), rtol=1e-13, atol=0)


@pytest.mark.parametrize("L,sorb,na,n,m", [(1, 40, 15, 300, 4097), (1, 12, 3, 400, 400), (2, 100, 3, 65, 1000), (3, 132, 3, 10, 513)])
def test_hij_2d_tiled_matrix_equals_untiled_rows_and_is_symmetric(L, sorb, na, n, m):
    """2-D get_hij_torch (bra tile in shared memory, one ket per thread) against the 3-D mode, which evaluates the same pairs
    one by one; for a symmetric Hamiltonian the square matrix is symmetric."""
    keys = S.random_onvs(max(n, m), sorb, na, na, seed=70 + L)
    n, m = min(n, keys.shape[0]), min(m, keys.shape[0])
    h1e, h2e = S.random_packed_integrals(sorb, seed=71, symmetric=True)
    bra, ket = dev(keys[:n]), dev(keys[:m])
    H = ops.get_hij_torch(bra, ket, dev(h1e), dev(h2e), sorb, 2 * na)
    rows = ops.get_hij_torch(bra, ket.unsqueeze(0).expand(n, m, ket.size(1)).contiguous(), dev(h1e), dev(h2e), sorb, 2 * na)
    assert torch.equal(H, rows)
    k = min(n, m)
    assert torch.equal(H[:k, :k], H[:k, :k].T.contiguous())
    want = O.hij(keys[:4], keys[:m], h1e, h2e, sorb, 2 * na) if hasattr(O, "hij") else None
    if want is not None:
        np.testing.assert_array_equal(H[:4].cpu().numpy(), want)


@pytest.mark.parametrize("cplx", [False, True])
def test_block_route_product_table_every_tile_shape(cplx):
    """Block kernel at its production settings on a table where the samples per beta string range from 1 to 150 (tiles of every
    size, groups below the threshold left to the per-sample kernel, hit lists long enough for the warp-per-sample evaluator):
    all samples against the three-call path, a subset against the oracle."""
    sorb, noA, noB, nele = 40, 15, 15, 30
    base = S.random_onvs(4000, sorb, noA, noB, seed=90).view(np.uint64).reshape(-1)
    a_str = np.unique(base & np.uint64(0x5555555555555555))[:150]
    b_str = np.unique(base & np.uint64(0xAAAAAAAAAAAAAAAA))[:120]
    rng = np.random.default_rng(91)
    rows = []
    for i, b in enumerate(b_str):  # 1, 2, 3, ... up to 150 alpha strings on the i-th beta string
        take = a_str[rng.permutation(a_str.size)[: 1 + (i * 5) % 150]]
        rows.append(take | b)
    keys = np.unique(np.concatenate(rows)).view(np.uint8).reshape(-1, 8)
    assert keys.shape[0] >= 4096  # above block_min_samples: the default route is the block route
    psi = S.random_psi(keys.shape[0], seed=92, complex_=cplx)
    h1e, h2e = S.random_packed_integrals(sorb, seed=7, symmetric=True)
    lut = WavefunctionLUT(dev(keys), dev(psi), sorb, DEV, rank=0, world_size=1)
    x = lut.bra_key
    e1, _, p1 = local_energy_sample_space(x, dev(h1e), dev(h2e), lut, sorb, nele, noA, noB, dtype=lut.dtype)
    e3, _, p3 = local_energy_three_call(x, dev(h1e), dev(h2e), lut, sorb, nele, noA, noB, dtype=lut.dtype, batch=2048)
    # (terms of order 1 that cancel to ~1e-5 for a few samples: the two routes add them in different orders)
    np.testing.assert_allclose(e1.cpu().numpy(), e3.cpu().numpy(), rtol=1e-12, atol=1e-14)
    assert torch.equal(torch.view_as_real(p1) if cplx else p1, torch.view_as_real(p3) if cplx else p3)
    pick = np.arange(0, keys.shape[0], keys.shape[0] // 24)[:24]
    skeys, spsi = lut.bra_key.cpu().numpy(), lut.wf_value.cpu().numpy()
    want = O.eloc_sample_space(skeys[pick], h1e, h2e, skeys, spsi, sorb, nele, noA, noB)
    np.testing.assert_allclose(e1.cpu().numpy()[pick], want, rtol=1e-12, atol=1e-14)


@pytest.mark.parametrize("L,sorb,na", [(1, 40, 15), (2, 100, 25), (3, 132, 3)])
def test_non_disjoint_merge_on_the_device_equals_torch_unique(L, sorb, na):
    """exchange_unique_samples(disjoint=False) on one rank: merged rows in torch.unique(dim=0) order (the reference's merge,
    vmc/sample.py:672-690), psi of the first occurrence, counts summed -- through the library's sort instead of torch.unique."""
    from pynqs_b200.distributed import exchange_unique_samples

    base = S.random_onvs(3000, sorb, na, na, seed=95 + L)
    rng = np.random.default_rng(96)
    pick = rng.integers(0, base.shape[0], size=7000)
    onv, psi = dev(base[pick]), dev(S.random_psi(7000, seed=97, complex_=True))
    counts = dev(rng.integers(1, 9, size=7000).astype(np.int64))
    uniq, wf, cnt = exchange_unique_samples(onv, psi, counts, disjoint=False)
    want_u, want_inv = torch.unique(onv, dim=0, return_inverse=True)
    assert torch.equal(uniq, want_u)
    first = torch.full((want_u.size(0),), 7000, dtype=torch.int64, device=DEV).scatter_reduce_(0, want_inv, torch.arange(7000, device=DEV), reduce="amin")
    assert torch.equal(torch.view_as_real(wf), torch.view_as_real(psi[first]))
    assert torch.equal(cnt, torch.zeros(want_u.size(0), dtype=torch.int64, device=DEV).index_add_(0, want_inv, counts))
