"""Rank-for-rank differential test of the multi-rank plumbing against the UNMODIFIED reference Python under gloo on CPU
(SURVEY.md section 8e): the reference's own utils/distributed/comm.py, utils/stats (operator_statistics) and
Sampler.gather_scatter_sample (vmc/sample.py:627-772) -- imported from baseline/_ref with the reference's CPU extension as
libs.C_extension -- run in the same processes as pynqs_b200.compat on the same inputs.

Needs baseline/_ref (python baseline/make_ref.py) and oracle/_ref (python oracle/build_ref.py); skipped where the reference
was never mounted."""
import os
import socket
import sys
import types

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _have_reference():
    sys.path.insert(0, ROOT)
    from oracle import build_ref

    return build_ref.available(1) and os.path.isdir(os.path.join(ROOT, "baseline", "_ref", "utils"))


pytestmark = pytest.mark.skipif(not _have_reference(), reason="baseline/_ref or oracle/_ref missing (reference not mounted)")


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _import_reference():
    """reference utils / vmc on the reference's own CPU extension (never this repo's shim: CPU tensors)"""
    from oracle.build_ref import load_ref

    ref = load_ref(1)
    libs = types.ModuleType("libs")
    libs.__path__ = []
    sys.modules["libs"] = libs
    sys.modules["libs.C_extension"] = ref
    libs.C_extension = ref
    sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))
    import utils.distributed.comm as rcomm
    import utils.stats as rstats
    import vmc.sample as rsample

    return rcomm, rstats, rsample


def _same(a, b):
    if a is None or b is None:
        return a is None and b is None
    if isinstance(a, (list, tuple)):
        return len(a) == len(b) and all(_same(x, y) for x, y in zip(a, b))
    return a.shape == b.shape and a.dtype == b.dtype and torch.equal(a, b)


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.set_default_dtype(torch.double)  # as every shipped input does (main.py:30)
        sys.path.insert(0, ROOT)
        rcomm, rstats, rsample = _import_reference()
        from pynqs_b200 import synthetic as S
        from pynqs_b200.compat import distributed as ccomm
        from pynqs_b200.compat import stats as cstats
        from pynqs_b200.compat.sampler import gather_scatter_sample

        dev = torch.device("cpu")
        g = torch.Generator().manual_seed(100 + rank)
        fails, n_checks = [], [0]

        def check(name, a, b):
            n_checks[0] += 1
            if not _same(a, b):
                fails.append(name)

        # ---- comm.py wrappers ---------------------------------------------------------------------------------------------
        n_r = 40 + 7 * rank  # ragged
        for dt in (torch.float64, torch.complex128, torch.int64):
            t = torch.randn(n_r, 3, generator=g).to(dt) if dt != torch.int64 else torch.randint(0, 99, (n_r,), generator=g)
            check(f"all_gather_tensor {dt}", ccomm.all_gather_tensor(t, dev, world), rcomm.all_gather_tensor(t, dev, world))
            check(f"gather_tensor {dt}", ccomm.gather_tensor(t, dev, world, 0), rcomm.gather_tensor(t, dev, world, 0))
        for rows, dt in ((1000, torch.float64), (10, torch.uint8), (7, torch.float64)):
            full = (torch.arange(rows * 8, dtype=torch.float64).reshape(rows, 8) % 251).to(dt) if rank == 0 else None
            check(f"scatter_tensor {rows}", ccomm.scatter_tensor(full, dev, dt, world), rcomm.scatter_tensor(full, dev, dt, world))
            one = torch.arange(rows, dtype=torch.float64).to(dt) if rank == 0 else None
            check(f"scatter_tensor 1-D {rows}", ccomm.scatter_tensor(one, dev, dt, world), rcomm.scatter_tensor(one, dev, dt, world))
        for dt in (torch.float64, torch.complex128, torch.uint8):
            src = (torch.randn(33, 2, generator=g) * 50).to(dt) if rank == 0 else None
            check(f"broadcast_tensor {dt}", ccomm.broadcast_tensor(src, dev, dt), rcomm.broadcast_tensor(src, dev, dt))
        a, b = torch.full((5,), float(rank + 1)), torch.full((5,), float(rank + 1))
        ccomm.all_reduce_tensor(a, world_size=world)
        rcomm.all_reduce_tensor(b, world_size=world)
        check("all_reduce_tensor", a, b)
        outs = ccomm.all_reduce_tensor([torch.ones(2) * rank, torch.ones(3)], world_size=world, in_place=False)
        refs = rcomm.all_reduce_tensor([torch.ones(2) * rank, torch.ones(3)], world_size=world, in_place=False)
        check("all_reduce_tensor list", outs, refs)
        assert ccomm.get_rank() == rcomm.get_rank() and ccomm.get_world_size() == rcomm.get_world_size()

        # ---- Sampler.gather_scatter_sample: the reference's method on a stand-in `self` --------------------------------
        sorb, noA = 40, 15
        keys = S.random_onvs(1000, sorb, noA, noA, seed=100)
        states = np.unpackbits(keys, axis=1, bitorder="little")[:, :sorb]
        psi_all = S.random_psi(1000, seed=101, complex_=True)
        cuts = np.linspace(0, 1000, world + 1).astype(int)
        stats_in = {}
        for same_tree in (True, False):
            sel = np.arange(cuts[rank], cuts[rank + 1])
            if not same_tree:  # overlapping pieces: ranks share samples 400..499
                sel = np.unique(np.concatenate([sel, np.arange(400, 500)]))
            uniq = torch.from_numpy(states[sel].astype(np.uint8))
            counts = torch.from_numpy(np.arange(1, len(sel) + 1, dtype=np.int64))
            wf = torch.from_numpy(psi_all[sel])

            def me():
                return types.SimpleNamespace(sorb=sorb, device=dev, world_size=world, rank=rank, use_LUT=True, use_same_tree=same_tree,
                                             dtype=torch.complex128, all_sample_counts=None)

            s_ref, s_new = me(), me()
            u0, p0, pr0, lut0 = rsample.Sampler.gather_scatter_sample(s_ref, uniq, counts, wf)
            u1, p1, pr1, lut1 = gather_scatter_sample(s_new, uniq, counts, wf)
            check(f"gather_scatter unique_rank same_tree={same_tree}", u1, u0)
            check(f"gather_scatter placeholder same_tree={same_tree}", p1, p0)
            if not (pr1.dtype == pr0.dtype and torch.allclose(pr1, pr0, rtol=1e-15, atol=0)):
                fails.append(f"gather_scatter prob same_tree={same_tree}")
            check(f"gather_scatter LUT keys same_tree={same_tree}", lut1.bra_key, lut0.bra_key)
            check(f"gather_scatter LUT values same_tree={same_tree}", lut1.wf_value, lut0.wf_value)
            if rank == 0:
                check(f"gather_scatter all_sample_counts same_tree={same_tree}", s_new.all_sample_counts, s_ref.all_sample_counts)
            stats_in[same_tree] = (u0.size(0), pr0)

        # ---- operator_statistics on this rank's slice, the prob * world convention of the sampler -----------------------
        for cplx in (False, True):
            n_loc, prob = stats_in[False]
            e = torch.randn(n_loc, generator=g, dtype=torch.float64) - 116.6
            if cplx:
                e = torch.complex(e, 1e-3 * torch.randn(n_loc, generator=g, dtype=torch.float64))
            total = torch.tensor([float(n_loc)])
            dist.all_reduce(total)
            r = rstats.operator_statistics(e, prob, int(total.item()), "E")
            c = cstats.operator_statistics(e, prob, int(total.item()), "E")
            for k in ("mean", "var", "sd", "se"):
                rv, cv = complex(r[k].item()), complex(c[k].item())
                if abs(rv - cv) > 1e-12 * max(1.0, abs(rv)):
                    fails.append(f"operator_statistics {k} complex={cplx}: {rv} vs {cv}")
            if repr(r)[:24] != repr(c)[:24]:
                fails.append(f"operator_statistics repr: {r!r} vs {c!r}")
        q.put((rank, fails, n_checks[0]))
    except Exception as ex:  # noqa: BLE001
        import traceback

        q.put((rank, [f"exception: {ex!r}\n{traceback.format_exc()}"], 0))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_compat_layer_equals_the_reference_python_rank_for_rank(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=300) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
    for rank, fails, n_checks in res:
        assert not fails, f"rank {rank}: " + "; ".join(fails)
        assert n_checks >= 25, f"rank {rank}: only {n_checks} comparisons ran"
