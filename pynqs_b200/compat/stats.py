"""utils/stats/dist_stats.py and mc_stats.py of the reference, name for name: dist_mean / dist_var / dist_stats /
operator_statistics with the reference's conventions (prob already multiplied by the world size, results divided by it,
sample.py:772 + comm.py:65-67), on ONE moments kernel and ONE collective (pynqs_b200.distributed) instead of three
all-reduces, an all-gather and four barriers (dist_stats.py:37, 54, 74-75)."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Tuple, TypedDict

import torch
from torch import Tensor

from ..distributed import energy_statistics
from .distributed import get_world_size


def _tensor(v, like: Tensor) -> Tensor:
    if isinstance(v, complex):
        return torch.tensor(v, dtype=torch.complex128, device=like.device)
    return torch.tensor(v, dtype=torch.float64, device=like.device)


def dist_stats(x: Tensor, prob: Optional[Tensor] = None, counts: Optional[int] = None, world_size: int = 1) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    """'mean', 'var', 'sd', 'se' (dist_stats.py:59-79)"""
    assert x.dim() == 1 and prob is not None and prob.dim() == 1
    st = energy_statistics(x, prob.to(torch.float64), counts)
    return _tensor(st["mean"], x), _tensor(st["var"], x), _tensor(st["sd"], x), _tensor(st["se"], x)


def dist_var(x: Tensor, prob: Optional[Tensor] = None, world_size: int = 1) -> Tuple[Tensor, Tensor]:
    mean, var, _, _ = dist_stats(x, prob, 1, world_size)
    return mean, var


def dist_mean(x: Tensor, prob: Optional[Tensor] = None, world_size: int = 1) -> Tensor:
    return dist_stats(x, prob, 1, world_size)[0]


class StatsDict(TypedDict):
    mean: Tensor
    var: Tensor
    sd: Tensor
    se: Tensor


@dataclass
class operator_statistics:
    """mc_stats.py:19-54"""

    operator: str = "Ȏ"
    world_size: int = 1
    stats_dict: StatsDict = None

    def __init__(self, x: Tensor, prob: Tensor, counts: Optional[int] = None, operator: Optional[str] = None) -> None:
        self.world_size = get_world_size()
        mean, var, sd, se = dist_stats(x, prob, counts, self.world_size)
        self.stats_dict = {"mean": mean, "var": var, "sd": sd, "se": se}
        if operator is not None:
            self.operator = operator

    def __getitem__(self, key: str) -> Tensor:
        return self.stats_dict[key]

    def to_dict(self) -> StatsDict:
        return self.stats_dict

    def __repr__(self) -> str:
        mean, se, var = self["mean"], self["se"], self["var"]
        return f"<{self.operator}> = {mean.real:.9E} ± {se.real:.3E} [σ² = {var.real:.3E}]"
