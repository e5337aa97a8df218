"""Shared helpers of the test-suite: golden loading and seeded input regeneration."""
import hashlib
import os

import numpy as np

from pynqs_b200 import synthetic as S

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

OPS_CASES = [
    "ops_c1_h6_12sorb",
    "ops_c1_h6_12sorb_f32",
    "ops_odd_14sorb_4a2b",
    "ops_c3_n2_52sorb",
    "ops_c4_h50_100sorb",
    "ops_l3_132sorb_3a2b",
]


def sha(a) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def load(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)


_cache = {}


def ops_inputs(name):
    """Regenerate the seeded inputs of an ops_* golden and check them against the stored digests."""
    if name in _cache:
        return _cache[name]
    g = load(name)
    sorb, noA, noB, n, seed = (int(g[k]) for k in ("sorb", "noA", "noB", "n", "seed"))
    dtype = np.dtype(str(g["dtype"]))
    bra = S.random_onvs(n, sorb, noA, noB, seed=seed)
    h1e, h2e = S.random_packed_integrals(sorb, seed=seed + 1, symmetric=bool(int(g["symmetric"])), dtype=dtype)
    assert sha(bra) == str(g["bra_sha"]), "seeded ONVs differ from the ones the golden was made with"
    assert sha(h2e) == str(g["h2e_sha"]), "seeded integrals differ from the ones the golden was made with"
    out = dict(g=g, sorb=sorb, noA=noA, noB=noB, nele=noA + noB, bra=bra, h1e=h1e, h2e=h2e, stride=int(g["stride"]))
    _cache[name] = out
    return out


def fe2s2():
    g = load("fe2s2_integrals")
    return dict(h1e=g["h1e"], h2e=g["h2e"], ci=g["ci_space"], sorb=int(g["sorb"]), noA=int(g["noA"]), noB=int(g["noB"]),
                nele=int(g["nele"]))


def toy_amplitude(states, sorb: int, cplx: bool):
    """The stand-in ansatz of the REDUCE goldens (same as tests/golden/make_golden.py:toy_amplitude): psi(x) from
    the +-1 occupation tensor [K, sorb] (torch, any device)."""
    import torch

    g = torch.Generator().manual_seed(77)
    w = torch.randn(sorb, 2, generator=g, dtype=torch.float64).to(states.device)
    z = states.to(torch.float64) @ w
    amp = 0.3 + torch.tanh(0.11 * z[:, 0]) ** 2
    return amp * torch.exp(1j * 0.7 * z[:, 1]) if cplx else amp * torch.sign(torch.cos(0.9 * z[:, 1]))
