"""pynqs_b200 -- B200-native VMC local-energy path behind the PyNQS `libs.C_extension` operator API.

Layout (only what the hot path needs):
  csrc/            hand-written CUDA kernels for sm_100a + the extern "C" ABI (include/pynqs_b200.h)
  _lib.py          ctypes loader of csrc/libpynqs_b200.so (fails loudly when it is missing)
  C_extension.py   host mirror of the reference operator API (libs/C_extension.pyi)
  lut.py           WavefunctionLUT mirror (utils/public_function.py:749-868) with a device hash index
  energy.py        sample-space local energy (vmc/energy/eloc.py:326-397) on the fused kernel
  distributed.py   NCCL exchange of unique samples + fused energy statistics
  synthetic.py     seeded ONVs and 8-fold-symmetric packed integrals for tests / bench
"""
__version__ = "0.1.0"
