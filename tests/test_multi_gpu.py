"""Two ranks, two GPUs, NCCL: the sharded step (exchange -> table -> local energy split by beta string -> statistics) gives
every rank exactly the local energies a single GPU computes for the same rows, and the same statistics.  Skipped with fewer
than two GPUs (run it with `gpurun --gpus 2`)."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, cplx, q):
    """One rank.  Whatever goes wrong is reported through the queue at once (the other rank would otherwise wait in a
    collective until the parent's timeout) and the process leaves without waiting for its peer."""
    import traceback

    code = 0
    try:
        _worker_body(rank, world, port, n, cplx, q)
    except BaseException:  # noqa: BLE001
        q.put((rank, "error", traceback.format_exc()))
        code = 1
    # leave without tearing the process group down: destroy_process_group() was seen to block for minutes after the
    # symmetric-memory rendezvous (both ranks had already delivered their results), which kept pytest from exiting
    q.close()
    q.join_thread()
    os._exit(code)


def _worker_body(rank, world, port, n, cplx, q):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    if True:
        from pynqs_b200 import C_extension as ops
        from pynqs_b200 import _lib
        from pynqs_b200 import synthetic as S
        from pynqs_b200.distributed import energy_statistics_amplitudes, exchange_unique_samples, sample_space_energy_sharded
        from pynqs_b200.lut import WavefunctionLUT, split_length_idx
        from pynqs_b200.step import SampleSpaceStep

        _lib.set_tuning("block_min_samples", 1)
        sorb, noA, noB = 40, 15, 15
        keys = S.random_onvs(n, sorb, noA, noB, seed=5)
        psi = S.random_psi(n, seed=6, complex_=cplx)
        h1e, h2e = S.random_packed_integrals(sorb, seed=7, symmetric=True)
        d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)  # noqa: E731
        cuts = [0] + split_length_idx(n, world)
        lo, hi = cuts[rank], cuts[rank + 1]
        uniq, wf, _ = exchange_unique_samples(d(keys[lo:hi]), d(psi[lo:hi]), None, disjoint=True)
        lut = WavefunctionLUT(uniq, wf, sorb, dev, rank=rank, world_size=world)
        eloc, psi0 = sample_space_energy_sharded(lut, d(h1e), d(h2e), sorb, 30, noA, noB)
        b, e = lut.rank_begin, lut.rank_end
        want, want0 = ops.eloc_sample_space(lut.bra_key[b:e].contiguous(), d(h1e), d(h2e), sorb, 30, noA, noB, lut.bra_key, lut.wf_value,
                                            lut.group_index)
        st = energy_statistics_amplitudes(eloc, psi0)
        # the same through the captured step (ragged pieces: 20001 samples over two ranks), eager calls first, then replays
        step = SampleSpaceStep(hi - lo, keys.shape[1], d(psi[:1]).dtype, d(h1e), d(h2e), sorb, 30, noA, noB, device=dev, warmup=1)
        for _ in range(4):
            e2, p2, st2 = step(d(keys[lo:hi]), d(psi[lo:hi]))
        assert step.graph is not None, step.why_eager
        assert torch.equal(torch.view_as_real(e2) if cplx else e2, torch.view_as_real(eloc) if cplx else eloc)
        assert st2.result()["mean"] == st["mean"]
        torch.cuda.synchronize()
        q.put((rank, bool(torch.equal(torch.view_as_real(eloc) if cplx else eloc, torch.view_as_real(want) if cplx else want)),
               bool(torch.equal(torch.view_as_real(psi0) if cplx else psi0, torch.view_as_real(want0) if cplx else want0)), st["mean"], st["var"], e - b))
    torch.cuda.synchronize()
    dist.barrier()


@pytest.mark.parametrize("n,cplx", [(20001, False), (6000, True)])
def test_sharded_step_equals_single_gpu(n, cplx):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, cplx, q), daemon=True) for r in range(2)]
    for p in procs:
        p.start()
    try:
        res = []
        for _ in range(2):
            item = q.get(timeout=180)
            assert item[1] != "error", f"rank {item[0]} failed:\n{item[2]}"
            res.append(item)
        res.sort(key=lambda t: t[0])
        for p in procs:
            p.join(timeout=60)
            assert p.exitcode == 0
    finally:  # never leave a rank behind that waits for its peer (it would keep pytest from exiting)
        for p in procs:
            if p.is_alive():
                p.kill()
    assert all(r[1] and r[2] for r in res), res
    assert res[0][3] == res[1][3] and res[0][4] == res[1][4]  # the same statistics on every rank
    assert res[0][5] + res[1][5] == n
