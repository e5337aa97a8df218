"""Stage-by-stage run of the two-rank test body with a log per rank (gpurun_out/mg_rank<r>.log) and hard timeouts:
    python profiles/micro/mg_debug.py [n] [complex]"""
import os
import sys
import time
import traceback

import numpy as np
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def worker(rank, world, port, n, cplx):
    log = open(os.path.join(ROOT, "gpurun_out", f"mg_rank{rank}.log"), "a")

    def say(msg):
        log.write(f"[{time.time() % 1000:8.2f}] n={n} cplx={cplx} {msg}\n")
        log.flush()

    try:
        import torch.distributed as dist

        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        torch.cuda.set_device(rank)
        dev = torch.device("cuda", rank)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
        say("process group up")
        from pynqs_b200 import C_extension as ops
        from pynqs_b200 import _lib, peer
        from pynqs_b200 import synthetic as S
        from pynqs_b200.distributed import energy_statistics_amplitudes, exchange_unique_samples, sample_space_energy_sharded
        from pynqs_b200.lut import WavefunctionLUT, split_length_idx
        from pynqs_b200.step import SampleSpaceStep

        _lib.set_tuning("block_min_samples", 1)
        for a in sys.argv[3:]:
            if "=" in a:
                k, v = a.split("=")
                _lib.set_tuning(k, int(v))
        sorb, noA, noB = 40, 15, 15
        keys = S.random_onvs(n, sorb, noA, noB, seed=5)
        psi = S.random_psi(n, seed=6, complex_=cplx)
        h1e, h2e = S.random_packed_integrals(sorb, seed=7, symmetric=True)
        d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)  # noqa: E731
        cuts = [0] + split_length_idx(n, world)
        lo, hi = cuts[rank], cuts[rank + 1]
        uniq, wf, _ = exchange_unique_samples(d(keys[lo:hi]), d(psi[lo:hi]), None, disjoint=True)
        torch.cuda.synchronize()
        say(f"exchange done ({peer.route()})")
        lut = WavefunctionLUT(uniq, wf, sorb, dev, rank=rank, world_size=world)
        eloc, psi0 = sample_space_energy_sharded(lut, d(h1e), d(h2e), sorb, 30, noA, noB)
        torch.cuda.synchronize()
        say("sharded energy done")
        b, e = lut.rank_begin, lut.rank_end
        want, want0 = ops.eloc_sample_space(lut.bra_key[b:e].contiguous(), d(h1e), d(h2e), sorb, 30, noA, noB, lut.bra_key, lut.wf_value,
                                            lut.group_index)
        torch.cuda.synchronize()
        r = lambda t: torch.view_as_real(t) if cplx else t  # noqa: E731
        say(f"reference rows done: equal {bool(torch.equal(r(eloc), r(want)))}, max diff {float((r(eloc) - r(want)).abs().max()):.3e}")
        st = energy_statistics_amplitudes(eloc, psi0)
        say(f"statistics {st['mean']}")
        step = SampleSpaceStep(hi - lo, keys.shape[1], d(psi[:1]).dtype, d(h1e), d(h2e), sorb, 30, noA, noB, device=dev, warmup=1)
        for it in range(4):
            e2, p2, st2 = step(d(keys[lo:hi]), d(psi[lo:hi]))
            torch.cuda.synchronize()
            say(f"step call {it}: graph {step.graph is not None} ({step.why_eager}), equal {bool(torch.equal(r(e2), r(eloc)))}, "
                f"max diff {float((r(e2) - r(eloc)).abs().max()):.3e}")
        say(f"step statistics equal {st2.result()['mean'] == st['mean']}")
        say("ALL DONE")
    except BaseException:  # noqa: BLE001
        say("EXCEPTION\n" + traceback.format_exc())
    finally:
        log.close()
        os._exit(0)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 20001
    cplx = len(sys.argv) > 2 and sys.argv[2] == "complex"
    import socket

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    procs = [ctx.Process(target=worker, args=(r, 2, port, n, cplx), daemon=True) for r in range(2)]
    for p in procs:
        p.start()
    t0 = time.time()
    while time.time() - t0 < 75 and any(p.is_alive() for p in procs):
        time.sleep(0.5)
    for p in procs:
        if p.is_alive():
            p.kill()
    print("elapsed", round(time.time() - t0, 1))


if __name__ == "__main__":
    main()
