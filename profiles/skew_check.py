#!/usr/bin/env python
"""Sensitivity of the one-pass local energy to the shape of the table (run on a GPU box):
Fe2S2 geometry, 1e6 unique keys whose BETA strings come from only `nbeta` distinct strings with Zipf-like
weights, so the groups the scan kernel walks are large and uneven (uniform random samples: 15504 strings,
~64 keys each).  Prints samples/s and checks a few samples against the oracle."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from oracle import oracle as O  # noqa: E402
from pynqs_b200 import C_extension as ops  # noqa: E402
from pynqs_b200 import synthetic as S  # noqa: E402
from pynqs_b200.lut import WavefunctionLUT  # noqa: E402

SORB, NOA, NOB, NELE = 40, 15, 15, 30


def strings(n, rng):
    occ = np.argsort(rng.random((n, 20)), axis=1)[:, :15]
    out = np.zeros(n, dtype=np.uint64)
    for c in range(15):
        out |= np.uint64(1) << (2 * occ[:, c]).astype(np.uint64)
    return out


def main(nbeta=2000, N=1_000_000, n_eval=200_000):
    rng = np.random.default_rng(3)
    beta = np.unique(strings(8 * nbeta, rng))[:nbeta] << np.uint64(1)
    nbeta = beta.size
    w = 1.0 / np.arange(1, nbeta + 1) ** 0.8
    keys = np.unique(strings(3 * N, rng) | beta[rng.choice(nbeta, size=3 * N, p=w / w.sum())])
    keys = keys[rng.permutation(keys.size)[:N]]
    kb = keys & np.uint64(0xAAAAAAAAAAAAAAAA)
    _, cnt = np.unique(kb, return_counts=True)
    print(f"nbeta {nbeta}: {keys.size} keys, beta groups: mean {cnt.mean():.0f}, median {np.median(cnt):.0f}, max {cnt.max()}")
    k8 = keys.view(np.uint8).reshape(-1, 8)
    psi = S.random_psi(keys.size, seed=5)
    h1e, h2e = S.random_packed_integrals(SORB, seed=7, symmetric=True)
    dev = "cuda:0"
    d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)  # noqa: E731
    lut = WavefunctionLUT(d(k8), d(psi), SORB, dev, rank=0, world_size=1)
    gidx = lut.group_index
    x = lut.bra_key[torch.randperm(keys.size, generator=torch.Generator().manual_seed(1))[:n_eval].to(dev)].contiguous()  # like VMC: the samples ARE table rows
    dh1, dh2 = d(h1e), d(h2e)
    for _ in range(2):
        eloc, _ = ops.eloc_sample_space(x, dh1, dh2, SORB, NELE, NOA, NOB, lut.bra_key, lut.wf_value, gidx)
    torch.cuda.synchronize()
    t0 = time.time()
    eloc, _ = ops.eloc_sample_space(x, dh1, dh2, SORB, NELE, NOA, NOB, lut.bra_key, lut.wf_value, gidx)
    torch.cuda.synchronize()
    dt = time.time() - t0
    print(f"  {n_eval / dt / 1e6:.1f} M samples/s ({dt * 1e3:.2f} ms for {n_eval} samples)")
    xs = x[:4].cpu().numpy()
    want = O.eloc_sample_space(xs, h1e, h2e, lut.bra_key.cpu().numpy(), lut.wf_value.cpu().numpy(), SORB, NELE, NOA, NOB)
    np.testing.assert_allclose(eloc[:4].cpu().numpy(), want, rtol=1e-12, atol=0)
    print("  oracle check ok")


if __name__ == "__main__":
    for nb in (15504, 4000, 1000, 250):
        main(nb)
