"""Drop-in mirrors of the reference modules either side of the local-energy kernels:

    pynqs_b200.compat.distributed   <->  utils/distributed/comm.py      (same names, signatures, return values)
    pynqs_b200.compat.stats         <->  utils/stats/{dist_stats,mc_stats}.py
    pynqs_b200.compat.sampler       <->  Sampler.gather_scatter_sample  (vmc/sample.py:627-772)

PyNQS keeps importing `utils.distributed` / `utils.stats`; INTEGRATION.md section 6 shows the two re-export lines.
"""
