"""Per-episode reward statistics: the only quantity that crosses ranks.

The reference logs `world.total_reward` once per epoch (sorrel/environment.py:191-199 ->
Logger.record_turn, sorrel/utils/logging.py:48-72).  With N instances sharded over R ranks the
logger is fed mean / std / min / max of the per-instance totals: each rank reduces its shard with
`sb_reward_stats` (deterministic CUDA reduction) and ONE tiny collective (40 bytes per rank,
all_gather over NCCL/NVLink; gloo in the CPU tests) combines the partials.  There is no collective
on the step / observation / buffer path.
"""
from __future__ import annotations

import ctypes as C

import torch

from ._abi import check, lib


def combine_partials(parts: torch.Tensor) -> torch.Tensor:
    """[R, 5] per-rank {sum, sum_sq, min, max, count} -> [5] global (any device)."""
    return torch.stack([parts[:, 0].sum(), parts[:, 1].sum(), parts[:, 2].min(), parts[:, 3].max(),
                        parts[:, 4].sum()])


def all_reduce_partial(partial: torch.Tensor, group=None) -> torch.Tensor:
    """One all_gather of the 5-vector across the process group, combined identically on every rank."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return combine_partials(partial[None])
    world = dist.get_world_size(group)
    parts = [torch.empty(5, dtype=partial.dtype, device=partial.device) for _ in range(world)]
    dist.all_gather(parts, partial.contiguous(), group=group)
    return combine_partials(torch.stack(parts))


def summarize(stats5: torch.Tensor) -> dict:
    """{sum, sum_sq, min, max, count} -> the numbers a Logger wants (synchronises)."""
    s, ss, mn, mx, n = (float(x) for x in stats5.tolist())
    mean = s / n
    var = max(ss / n - mean * mean, 0.0)
    return {"mean": mean, "std": var ** 0.5, "min": mn, "max": mx, "count": int(n), "sum": s}


class RewardStats:
    def __init__(self, device):
        self.lib = lib()
        self.device = torch.device(device)
        self.scratch = torch.zeros(self.lib.sb_stats_scratch_bytes(), dtype=torch.uint8, device=self.device)

    def local(self, total_reward: torch.Tensor) -> torch.Tensor:
        out = torch.empty(5, dtype=torch.float64, device=self.device)
        check(self.lib.sb_reward_stats(total_reward.data_ptr(), total_reward.numel(), out.data_ptr(),
                                       self.scratch.data_ptr(),
                                       torch.cuda.current_stream(self.device).cuda_stream), "sb_reward_stats")
        return out

    def all_reduce(self, total_reward: torch.Tensor, group=None) -> torch.Tensor:
        """Global {sum, sum_sq, min, max, count} as a device tensor (no host sync)."""
        return all_reduce_partial(self.local(total_reward), group)
