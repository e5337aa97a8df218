"""Host-side logic that runs without a GPU: key sort, LUT bookkeeping, integral packing,
synthetic generators."""
import numpy as np
import torch

from oracle import oracle as O
from pynqs_b200 import C_extension as ops
from pynqs_b200 import synthetic as S
from pynqs_b200.lut import WavefunctionLUT, sort_onv, split_length_idx

from util import load


def test_sort_onv_equals_reference_lexsort():
    g = load("lut_sort_order")
    for L, sorb_l, na in ((1, 40, 15), (2, 100, 25), (3, 132, 3)):
        k = S.random_onvs(300, sorb_l, na, na, seed=20 + L)
        k = np.concatenate([k, k[:40]])[np.random.default_rng(5).permutation(340)]
        idx = sort_onv(torch.from_numpy(k)).numpy()
        np.testing.assert_array_equal(idx, g[f"idx_L{L}"])


def test_sort_onv_high_bit_and_empty():
    k = np.zeros((4, 8), dtype=np.uint8)
    k[0, 7] = 0x80  # bit 63 set: must sort last (unsigned order)
    k[1, 0] = 1
    k[2, 7] = 0x7F
    idx = sort_onv(torch.from_numpy(k)).tolist()
    assert idx == [3, 1, 2, 0]
    assert sort_onv(torch.empty((0, 8), dtype=torch.uint8)).numel() == 0


def test_split_length_idx_matches_reference_examples():
    assert split_length_idx(11, 3) == [4, 8, 11]          # utils/public_function.py:736-740
    assert split_length_idx(1000, 3) == [334, 667, 1000]  # ragged 334/333/333


def test_lut_mirror_bookkeeping_on_cpu():
    keys = S.random_onvs(1000, 40, 15, 15, seed=3)
    psi = S.random_psi(1000, seed=4)
    lut = WavefunctionLUT(torch.from_numpy(keys), torch.from_numpy(psi), 40, rank=1, world_size=3)
    order = O.sort_onv(keys)
    np.testing.assert_array_equal(lut.bra_key.numpy(), keys[order])
    np.testing.assert_array_equal(lut.wf_value.numpy(), psi[order])
    assert (lut.rank_begin, lut.rank_end) == (334, 667)
    np.testing.assert_array_equal(lut.index_value(0, 10).numpy(), psi[334:344])  # unsorted order of this rank
    assert lut.hash_index is None  # CPU tensors: no device index (and lookups would raise)


def test_compress_decompress_roundtrip_and_layout():
    sorb = 8
    rng = np.random.default_rng(0)
    a = rng.standard_normal((sorb,) * 4)
    a = a - a.transpose(1, 0, 2, 3)
    a = a - a.transpose(0, 1, 3, 2)
    a = a + a.transpose(2, 3, 0, 1)  # <ij||kl> antisymmetric in (ij), (kl), symmetric under pair swap
    h1 = rng.standard_normal((sorb, sorb))
    p1, p2 = ops.compress_h1e_h2e(h1, a, sorb)
    pair = sorb * (sorb - 1) // 2
    assert p1.shape == (sorb * sorb,) and p2.shape == (pair * (pair + 1) // 2,)
    # index rule of cpp_src/tensor/integral.cpp:24-42
    i, j, k, l = 5, 2, 3, 1
    ij, kl = i * (i - 1) // 2 + j, k * (k - 1) // 2 + l
    assert p2[ij * (ij + 1) // 2 + kl] == a[i, j, k, l]
    d1, d2 = ops.decompress_h1e_h2e(p1, p2, sorb)
    np.testing.assert_array_equal(d1, h1)
    np.testing.assert_array_equal(d2, a)


def test_synthetic_integrals_are_a_hamiltonian():
    """8-fold symmetric spatial integrals, antisymmetrised: dense H over the full 12-sorb space is symmetric."""
    h1e, h2e = S.random_packed_integrals(12, seed=52, symmetric=True)
    keys = S.random_onvs(400, 12, 3, 3, seed=51)
    assert keys.shape[0] == 400 and len({bytes(r) for r in keys}) == 400
    H = O.hij(keys[:60], keys[:60], h1e, h2e, 12, 6)
    np.testing.assert_array_equal(H, H.T)
    _, dense = ops.decompress_h1e_h2e(h1e, h2e, 12)
    # <pq||rs> = [pr|qs] - [ps|qr]: opposite-spin exchange part vanishes
    assert dense[2, 1, 2, 1] != 0 and dense[2, 1, 1, 2] == -dense[2, 1, 2, 1]


def test_random_onvs_have_fixed_occupation():
    x = S.random_onvs(256, 100, 25, 25, seed=9)
    bits = np.unpackbits(x, axis=1, bitorder="little")[:, :100]
    assert (bits[:, 0::2].sum(1) == 25).all() and (bits[:, 1::2].sum(1) == 25).all()
    assert len({bytes(r) for r in x}) == 256


def test_bench_host_helpers():
    """bench.py's host side: the algorithmic byte count of SURVEY.md 8(d), the parity block (tolerances, weights), the
    seeded tables (unique rows, fixed occupation, the skew of the Zipf variant) -- no GPU involved."""
    import bench

    assert bench.algorithmic_bytes_per_sample(7876, 1) == 259924            # Fe2S2, SURVEY.md 8(d)
    assert bench.algorithmic_bytes_per_sample(571876, 2, 16) == 16 + 571876 * 49 + 16
    psi = np.array([1.0, 2.0, -1.0, 0.5])
    ref = np.array([-1.0, -2.0, -3.0, -4.0])
    ok = bench.parity_block(ref * (1 + 1e-15), ref, psi, "x")
    assert ok["ok"] and ok["n"] == 4 and ok["max_rel_err"] < 1e-14 and ok["mean_energy_abs_diff_ha"] < 1e-13
    w = psi ** 2 / np.sum(psi ** 2)
    assert abs(ok["mean_energy_ha"] - float(np.sum(w * ref))) < 1e-15
    bad = bench.parity_block(ref * np.array([1, 1, 1 + 1e-9, 1]), ref, psi, "x")
    assert not bad["ok"] and bad["max_rel_err"] > 1e-10
    assert not bench.parity_block(np.array([np.nan, -2, -3, -4.0]), ref, psi, "x")["ok"]
    for kind in ("uniform", "zipf0.8"):
        keys = bench.make_table(kind, 20000)
        words = keys.view(np.uint64).reshape(-1)
        assert keys.shape == (20000, 8) and np.unique(words).size == 20000
        bits = np.unpackbits(keys, axis=1, bitorder="little")
        assert (bits[:, 0:40:2].sum(1) == 15).all() and (bits[:, 1:40:2].sum(1) == 15).all() and not bits[:, 40:].any()
        beta_counts = np.unique(words & np.uint64(0xAAAAAAAAAAAAAAAA), return_counts=True)[1]
        if kind != "uniform":
            assert beta_counts.max() > 20 * np.median(beta_counts)           # a few heavy strings, many light ones
        np.testing.assert_array_equal(keys, bench.make_table(kind, 20000))   # seeded
