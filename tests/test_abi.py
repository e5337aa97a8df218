"""The C-ABI library loads without a GPU and exports exactly what include/pynqs_b200.h declares."""
import ctypes
import os
import re

import pytest

from pynqs_b200 import _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    build.build()
    return _lib.load()


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "pynqs_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(pynqs_[a-z0-9_]+)\s*\(", txt)))


def test_header_and_loader_agree():
    assert header_symbols() == sorted(_lib.SYMBOLS)


def test_every_declared_symbol_is_exported(lib):
    for s in header_symbols():
        assert hasattr(lib, s), s


def test_host_only_entry_points(lib):
    assert lib.pynqs_abi_version() == 1
    out = ctypes.c_int64()
    assert lib.pynqs_num_sd(40, 15, 15, ctypes.byref(out)) == 0 and out.value == 7875   # Fe2S2
    assert lib.pynqs_num_sd(12, 3, 3, ctypes.byref(out)) == 0 and out.value == 117
    assert lib.pynqs_num_sd(52, 5, 5, ctypes.byref(out)) == 0 and out.value == 15435
    assert lib.pynqs_num_sd(100, 25, 25, ctypes.byref(out)) == 0 and out.value == 571875
    assert lib.pynqs_check_sorb(40, 30) == 0
    assert lib.pynqs_check_sorb(52, 10) == 0            # relaxed w.r.t. the reference
    assert lib.pynqs_check_sorb(193, 10) == _lib.EVALUE
    assert lib.pynqs_check_sorb(64, 121) == _lib.EOVERFLOW
    assert b"nele" in lib.pynqs_last_error()
    nb = ctypes.c_int64()
    # header + two directories of 2^21 16-byte slots + pool of 4N + 2 32-byte buckets + 2N u32 scratch
    assert lib.pynqs_hash_bytes(ctypes.c_int64(1000000), 1, ctypes.byref(nb)) == 0
    assert nb.value == 256 + 2 * 16 * (1 << 21) + (16 + 16) * 4000002 + 16000016
    assert lib.pynqs_hash_bytes(ctypes.c_int64(10), 4, ctypes.byref(nb)) == _lib.EVALUE


def test_error_mapping():
    from pynqs_b200 import C_extension as ops

    with pytest.raises(ValueError):
        ops.check_sorb(200, 10)
    with pytest.raises(OverflowError):
        ops.check_sorb(64, 200)
    ops.check_sorb(40, 30)
    assert ops.get_Num_SinglesDoubles(40, 15, 15) == 7875
    assert (ops.MAX_SORB, ops.MAX_SORB_LEN, ops.MAX_NELE) == (192, 3, 120)


def test_cpu_tensors_are_refused_loudly():
    import torch

    from pynqs_b200 import C_extension as ops

    bra = torch.zeros((1, 8), dtype=torch.uint8)
    with pytest.raises(RuntimeError, match="CUDA-only"):
        ops.get_comb_tensor(bra, 4, 2, 1, 1)
    with pytest.raises(RuntimeError, match="CUDA-only"):
        ops.wavefunction_lut(bra, bra, 4)
    with pytest.raises(NotImplementedError):
        ops.spin_flip_rand(bra, 4, 2, 1, 1, 0)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "pynqs_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.replace("# oracle", ""), f"{f} mentions the oracle"


def test_libs_c_extension_exposes_the_reference_names():
    """Every function / attribute of the reference's libs/C_extension.pyi (minus the experimental CUDA
    hash-table API, whose import is try/except-guarded in the reference, utils/public_function.py:24-30)."""
    import libs.C_extension as ext

    names = ["tensor_to_onv", "onv_to_tensor", "get_comb_tensor", "get_hij_torch", "get_comb_hij_fused", "MCMC_sample",
             "spin_flip_rand", "mps_vbatch", "permute_sgn", "convert_sites", "merge_rank_sample", "constrain_make_charts",
             "wavefunction_lut", "check_sorb", "compress_h1e_h2e", "decompress_h1e_h2e", "MAX_SORB", "MAX_SORB_LEN", "MAX_NELE"]
    for n in names:
        assert hasattr(ext, n), n
    import inspect

    assert list(inspect.signature(ext.get_comb_hij_fused).parameters)[:7] == ["bra", "h1e", "h2e", "sorb", "nele", "noA", "noB"]
    assert list(inspect.signature(ext.get_comb_tensor).parameters) == ["bra", "sorb", "nele", "noA", "noB", "flag_bit"]
    assert list(inspect.signature(ext.get_hij_torch).parameters) == ["bra", "ket", "h1e", "h2e", "sorb", "nele"]
    assert list(inspect.signature(ext.wavefunction_lut).parameters)[:4] == ["bra_key", "onv", "sorb", "little_endian"]


def test_every_tuning_knob_is_documented_and_settable(lib):
    """The knob table of pynqs_set_tuning (csrc/abi.cu) and its description in the header list the same names; every knob
    accepts its default-range value without a GPU, an unknown name or an out-of-range value is EVALUE."""
    src = open(os.path.join(ROOT, "pynqs_b200", "csrc", "abi.cu")).read()
    table = src[src.index("} knobs[] = {"): src.index("if (name == nullptr)")]
    knobs = re.findall(r'\{"([a-z_]+)",\s*&t\.', table)
    assert len(knobs) >= 9 and len(set(knobs)) == len(knobs)
    header = open(os.path.join(ROOT, "include", "pynqs_b200.h")).read()
    doc = header[header.index("Test / experiment knobs"): header.index("int pynqs_set_tuning")]
    for k in knobs:
        assert f'"{k}"' in doc, f"knob {k} is missing from include/pynqs_b200.h"
    try:
        for k, v in (("block_parts", 2), ("lut_pipeline", 0), ("eval_tiles", 2), ("block_min_group", 4)):
            assert lib.pynqs_set_tuning(k.encode(), ctypes.c_int64(v)) == 0
        assert lib.pynqs_set_tuning(b"block_parts", ctypes.c_int64(5)) == _lib.EVALUE
        assert lib.pynqs_set_tuning(b"no_such_knob", ctypes.c_int64(1)) == _lib.EVALUE
    finally:
        assert lib.pynqs_set_tuning(None, ctypes.c_int64(0)) == 0
