import sys, json, statistics
sys.path.insert(0, '.')
import numpy as np, torch
import bench
from pynqs_b200 import C_extension as ops
from pynqs_b200.lut import WavefunctionLUT
dev = torch.device('cuda', 0)
keys = bench.make_table('uniform', 1_000_000); psi = bench.make_psi(keys.shape[0], False)
h1e_np, h2e_np, _ = bench.load_integrals()
h1e, h2e = torch.from_numpy(h1e_np).to(dev), torch.from_numpy(h2e_np).to(dev)
lut = WavefunctionLUT(torch.from_numpy(keys).to(dev), torch.from_numpy(psi).to(dev), 40, dev, rank=0, world_size=1)
x = lut.bra_key[:32768]
prep = ops.PreparedIntegrals(h2e, 40)
def t(fn, reps=9):
    for _ in range(3): out = fn()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); out = fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return statistics.median(ts), out
M = 7876
f_ms, (comb, hmat) = t(lambda: ops.get_comb_hij_fused(x, h1e, h2e, 40, 30, 15, 15, prepared=prep))
flat = comb.view(-1, 8)
l_ms, _ = t(lambda: ops.wavefunction_lut(lut.bra_key, flat, 40, hash_index=lut.hash_index))
fb, lb = (16 * M + 8) * 32768, 17 * M * 32768
print(json.dumps({"fused_ms": f_ms, "fused_frac": fb / f_ms / 1e6 / 6538.3, "lut_ms": l_ms, "lut_frac": lb / l_ms / 1e6 / 6538.3}))
