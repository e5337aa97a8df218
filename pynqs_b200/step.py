"""One VMC local-energy step -- sample exchange -> sorted table + string-grouped copies -> one-pass E_loc -> energy statistics --
as a single object, captured in a CUDA graph.

At 8 GPUs the step is ~1 ms of ~50 short kernels and two or three collectives: launched one by one from Python it is bound by
launch latency, not by the GPU.  `SampleSpaceStep` runs the step eagerly a few times (allocations, lazy set-up, the peer-memory
rendezvous), then captures it once (torch.cuda.CUDAGraph: the library's kernels launch on torch's current stream, so they are
captured like torch's own; NCCL collectives and the peer-memory barriers capture too) and replays it for every new sample set
of the same size.  Results are identical to the eager calls -- the graph contains exactly those launches.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist
from torch import Tensor

from . import C_extension as ops
from .distributed import PendingStatistics, energy_statistics_amplitudes, exchange_unique_samples, sample_space_energy_sharded
from .lut import WavefunctionLUT


class SampleSpaceStep:
    """Fixed-shape step: this rank contributes `n_local` unique ONVs (uint8 [n_local, 8L]) and their psi (the ranks' counts
    may differ; the pieces must be disjoint, as with the sampler's use_same_tree).

        step = SampleSpaceStep(n_local, 8 * L, psi_dtype, h1e, h2e, sorb, nele, noa, nob)
        eloc, psi0, stats = step(onv, psi)          # tensors are views of static buffers, valid until the next call
        stats.result()                              # host read of the 7 doubles per rank

    `eloc` / `psi0` belong to this rank's rows of the sorted table (WavefunctionLUT.rank_begin .. rank_end); `step.lut` is
    the table of the last call."""

    def __init__(self, n_local: int, width: int, psi_dtype: torch.dtype, h1e: Tensor, h2e: Tensor, sorb: int, nele: int, noa: int,
                 nob: int, device: Optional[torch.device] = None, use_graph: bool = True, warmup: int = 2):
        self.dev = device or h1e.device
        self.args = (sorb, nele, noa, nob)
        self.h1e, self.h2e = h1e, h2e
        on = dist.is_available() and dist.is_initialized()
        self.rank = dist.get_rank() if on else 0
        self.world = dist.get_world_size() if on else 1
        # every rank's sample count, exchanged once here so that the step itself needs no size handshake (and no host read)
        self.sizes = [int(n_local)]
        if self.world > 1:
            mine = torch.tensor([int(n_local)], dtype=torch.int64, device=self.dev)
            allv = torch.empty(self.world, dtype=torch.int64, device=self.dev)
            dist.all_gather_into_tensor(allv, mine)
            self.sizes = allv.tolist()
        self.onv = torch.zeros((n_local, width), dtype=torch.uint8, device=self.dev)
        self.psi = torch.zeros(n_local, dtype=psi_dtype, device=self.dev)
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self.why_eager = "" if use_graph else "disabled by the caller"
        self._warmup = warmup
        self._calls = 0
        self._use_graph = use_graph
        self.out: Optional[Tuple[Tensor, Tensor, PendingStatistics]] = None
        self.lut: Optional[WavefunctionLUT] = None

    def _body(self):
        sorb, nele, noa, nob = self.args
        uniq, wf, _ = exchange_unique_samples(self.onv, self.psi, None, disjoint=True, sizes=self.sizes)
        lut = WavefunctionLUT(uniq, wf, sorb, self.dev, rank=self.rank, world_size=self.world)
        eloc, psi0 = sample_space_energy_sharded(lut, self.h1e, self.h2e, sorb, nele, noa, nob)
        st = energy_statistics_amplitudes(eloc, psi0, lazy=True)
        self.lut = lut
        return eloc, psi0, st

    def load(self, onv: Tensor, psi: Tensor, non_blocking: bool = True) -> None:
        """copy a new sample set (device or pinned host tensors) into the step's input buffers"""
        self.onv.copy_(onv, non_blocking=non_blocking)
        self.psi.copy_(psi, non_blocking=non_blocking)

    def run(self) -> Tuple[Tensor, Tensor, PendingStatistics]:
        """the step on whatever the input buffers hold"""
        if self.graph is not None:
            self.graph.replay()
            eloc, psi0, st = self.out
            return eloc, psi0, PendingStatistics(st._allv, st._world, st._amplitude, st._cplx, st._counts)
        self._calls += 1
        if self._use_graph and self._calls > self._warmup:
            try:
                torch.cuda.synchronize(self.dev)
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    out = self._body()
                self.graph, self.out = g, out
                return self.run()
            except Exception as ex:  # noqa: BLE001  -- something in the step cannot be captured here: stay eager
                self._use_graph = False
                self.why_eager = f"capture failed: {type(ex).__name__}: {str(ex)[:160]}"
                torch.cuda.synchronize(self.dev)
        return self._body()

    def __call__(self, onv: Tensor, psi: Tensor) -> Tuple[Tensor, Tensor, PendingStatistics]:
        self.load(onv, psi)
        return self.run()
