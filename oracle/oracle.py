"""numpy front-end of oracle/pynqs_oracle.c -- TEST INFRASTRUCTURE ONLY (see the C header).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product package pynqs_b200 never does.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "pynqs_oracle.c")
LIB = os.path.join(HERE, "liboracle.so")

_lib = None


def build(force: bool = False) -> str:
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(SRC):
        subprocess.check_call(["gcc", "-O2", "-std=c99", "-fPIC", "-shared", "-o", LIB, SRC, "-lm"])
    return LIB


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(LIB)
        _lib.orc_num_sd.restype = ctypes.c_int
    return _lib


def _p(a: np.ndarray):
    return a.ctypes.data_as(ctypes.c_void_p)


def _words(onv_u8: np.ndarray) -> np.ndarray:
    a = np.ascontiguousarray(onv_u8, dtype=np.uint8)
    assert a.shape[-1] % 8 == 0
    return a.view(np.uint64)


def num_sd(sorb: int, noA: int, noB: int) -> int:
    return int(lib().orc_num_sd(sorb, noA, noB))


def merged(bra_u8, sorb):
    x = _words(bra_u8)
    out = np.empty((x.shape[0], sorb), dtype=np.int32)
    lib().orc_merged(_p(x), ctypes.c_int64(x.shape[0]), sorb, _p(out))
    return out


def unpack(sorb, noA, noB, r):
    out = (ctypes.c_int * 5)()
    lib().orc_unpack(sorb, noA, noB, int(r), out)
    return list(out)


def comb(bra_u8, sorb, noA, noB):
    x = _words(bra_u8)
    n, L = x.shape
    M = num_sd(sorb, noA, noB) + 1
    out = np.empty((n, M, L), dtype=np.uint64)
    lib().orc_comb(_p(x), ctypes.c_int64(n), sorb, noA, noB, _p(out))
    return out.view(np.uint8).reshape(n, M, 8 * L)


def excitations(bra_u8, sorb, noA, noB):
    x = _words(bra_u8)
    n = x.shape[0]
    nsd = num_sd(sorb, noA, noB)
    orbs = np.empty((n, nsd, 4), dtype=np.int32)
    sgn = np.empty((n, nsd), dtype=np.int8)
    lib().orc_excitations(_p(x), ctypes.c_int64(n), sorb, noA, noB, _p(orbs), _p(sgn))
    return orbs, sgn


def _sfx(dtype):
    dtype = np.dtype(dtype)
    if dtype == np.float64:
        return "f64"
    if dtype == np.float32:
        return "f32"
    raise TypeError(dtype)


def comb_hij_fused(bra_u8, h1e, h2e, sorb, nele, noA, noB):
    x = _words(bra_u8)
    n, L = x.shape
    M = num_sd(sorb, noA, noB) + 1
    h1e = np.ascontiguousarray(h1e)
    h2e = np.ascontiguousarray(h2e, dtype=h1e.dtype)
    out = np.empty((n, M, L), dtype=np.uint64)
    hmat = np.empty((n, M), dtype=h1e.dtype)
    getattr(lib(), "orc_comb_hij_" + _sfx(h1e.dtype))(
        _p(x), _p(h1e), _p(h2e), ctypes.c_int64(n), sorb, nele, noA, noB, _p(out), _p(hmat)
    )
    return out.view(np.uint8).reshape(n, M, 8 * L), hmat


def hij(bra_u8, ket_u8, h1e, h2e, sorb, nele):
    x = _words(bra_u8)
    y = _words(ket_u8)
    n = x.shape[0]
    ket3d = y.ndim == 3
    m = y.shape[1] if ket3d else y.shape[0]
    h1e = np.ascontiguousarray(h1e)
    h2e = np.ascontiguousarray(h2e, dtype=h1e.dtype)
    out = np.empty((n, m), dtype=h1e.dtype)
    getattr(lib(), "orc_hij_" + _sfx(h1e.dtype))(
        _p(x), _p(y), _p(h1e), _p(h2e), ctypes.c_int64(n), ctypes.c_int64(m), int(ket3d), sorb, nele, _p(out)
    )
    return out


def onv_to_tensor(bra_u8, sorb, dtype=np.float64):
    x = _words(bra_u8)
    out = np.empty((x.shape[0], sorb), dtype=dtype)
    getattr(lib(), "orc_onv_to_tensor_" + _sfx(dtype))(_p(x), ctypes.c_int64(x.shape[0]), sorb, _p(out))
    return out


def tensor_to_onv(states_u8, sorb):
    s = np.ascontiguousarray(states_u8, dtype=np.uint8).reshape(-1, sorb)
    L = (sorb - 1) // 64 + 1
    out = np.empty((s.shape[0], L), dtype=np.uint64)
    lib().orc_tensor_to_onv(_p(s), ctypes.c_int64(s.shape[0]), sorb, _p(out))
    return out.view(np.uint8).reshape(-1, 8 * L)


def lut(key_u8, onv_u8):
    k = _words(key_u8)
    q = _words(onv_u8)
    N, L = k.shape
    n = q.shape[0]
    assert q.shape[1] == L
    idx = np.empty(n, dtype=np.int64)
    mask = np.empty(n, dtype=np.uint8)
    lib().orc_lut(_p(k), ctypes.c_int64(N), _p(q), ctypes.c_int64(n), L, _p(idx), _p(mask))
    return idx, mask.astype(bool)


def sort_onv(key_u8):
    """Row order of the sorted table: stable LSD sort over byte columns 0..8L-1
    (utils/public_function.py:626-689) == ascending little-endian multi-word integer."""
    k = np.ascontiguousarray(key_u8, dtype=np.uint8)
    return np.lexsort([k[:, c] for c in range(k.shape[1])])


def eloc_rows(idx, hmat, psi):
    idx = np.ascontiguousarray(idx, dtype=np.int64)
    hmat = np.ascontiguousarray(hmat, dtype=np.float64)
    n, M = hmat.shape
    cpsi = np.iscomplexobj(psi)
    psi = np.ascontiguousarray(psi, dtype=np.complex128 if cpsi else np.float64)
    out = np.empty(n, dtype=psi.dtype)
    lib().orc_eloc_rows(_p(idx), _p(hmat), _p(psi), int(cpsi), ctypes.c_int64(n), ctypes.c_int64(M), _p(out))
    return out


def eloc_sample_space(bra_u8, h1e, h2e, key_sorted_u8, psi_sorted, sorb, nele, noA, noB):
    """Reference three-call path (fused -> lut -> reduce), vmc/energy/eloc.py:326-397."""
    c, h = comb_hij_fused(bra_u8, np.asarray(h1e, dtype=np.float64), np.asarray(h2e, dtype=np.float64), sorb, nele, noA, noB)
    n, M, w = c.shape
    idx, _ = lut(key_sorted_u8, c.reshape(n * M, w))
    return eloc_rows(idx.reshape(n, M), h, psi_sorted)


def reduced(bra_u8, h1e, h2e, sorb, nele, noA, noB, eps):
    """REDUCE method, kept set (vmc/energy/eloc.py:257-259, 289): torch.where(|Hmat| >= eps) over the
    row-major [n, M] matrix of the fused operator -> (x [K, 8L], hij [K], flat idx [K], offsets [n + 1])."""
    c, h = comb_hij_fused(bra_u8, h1e, h2e, sorb, nele, noA, noB)
    n, M, w = c.shape
    eps_t = h.dtype.type(eps)  # torch compares in the tensor's dtype
    keep = np.abs(h) >= eps_t
    idx = np.nonzero(keep.reshape(-1))[0].astype(np.int64)
    offsets = np.concatenate([[0], np.cumsum(keep.sum(1))]).astype(np.int64)
    return c.reshape(n * M, w)[idx], h.reshape(-1)[idx], idx, offsets


def reduce_eloc(psi_kept, hij_kept, idx, offsets, M):
    """Last lines of _reduce_psi (eloc.py:294-307): psi scattered into zeros [n, M], eloc = sum (psi / psi[:, 0]) * Hmat."""
    n = len(offsets) - 1
    cplx = np.iscomplexobj(psi_kept)
    psi = np.zeros(n * M, dtype=np.complex128 if cplx else np.float64)
    hm = np.zeros(n * M, dtype=np.float64)
    psi[idx] = psi_kept
    hm[idx] = hij_kept
    psi = psi.reshape(n, M)
    with np.errstate(divide="ignore", invalid="ignore"):
        eloc = ((psi.T / psi[:, 0]).T * hm.reshape(n, M)).sum(-1)
    return eloc, psi[:, 0].copy()
