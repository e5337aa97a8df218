"""Drop-in for PyNQS's compiled `libs.C_extension` (stubs: libs/C_extension.pyi of the reference):
put this repository's root on sys.path ahead of PyNQS's and its Python runs on libpynqs_b200.so.
See INTEGRATION.md."""
from pynqs_b200.C_extension import *  # noqa: F401,F403
from pynqs_b200.C_extension import (  # noqa: F401
    MAX_NELE, MAX_SORB, MAX_SORB_LEN, BKDR, MCMC_sample, check_sorb, compress_h1e_h2e, constrain_make_charts,
    convert_sites, decompress_h1e_h2e, get_comb_hij_fused, get_comb_tensor, get_hij_torch, merge_rank_sample,
    mps_vbatch, onv_to_tensor, permute_sgn, spin_flip_rand, tensor_to_onv, wavefunction_lut, wavefunction_lut_map,
)
