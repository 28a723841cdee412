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
    """One all_gather of the 5-vector across the process group, combined identically on every rank.  On CUDA the
    gather lands in one [R, 5] tensor and ONE small kernel (`sb_stats_combine`) folds it -- two stream-ordered launches
    per episode instead of a dozen elementwise ones; CPU tensors (gloo tests) are combined with torch."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return partial.clone() if partial.is_cuda else combine_partials(partial[None])
    world = dist.get_world_size(group)
    if not partial.is_cuda:
        parts = [torch.empty(5, dtype=partial.dtype, device=partial.device) for _ in range(world)]
        dist.all_gather(parts, partial.contiguous(), group=group)
        return combine_partials(torch.stack(parts))
    parts = torch.empty((world, 5), dtype=partial.dtype, device=partial.device)
    dist.all_gather_into_tensor(parts, partial.contiguous(), group=group)
    out = torch.empty(5, dtype=partial.dtype, device=partial.device)
    check(lib().sb_stats_combine(parts.data_ptr(), world, out.data_ptr(),
                                 torch.cuda.current_stream(partial.device).cuda_stream), "sb_stats_combine")
    return out


class PendingStats:
    """Handle of an all_gather in flight on the collective's own stream (`all_reduce_partial_async`);
    `result()` joins it and combines the partials (device tensor [5]).

    Measured on 2 x B200 (131 072 instances per GPU, boundary every 4 steps): leaving the NCCL gather un-joined
    costs ~10 ms per boundary (3.6 ms per step instead of 1.2 ms with the stream-ordered `all_reduce_partial`),
    so bench.py and Environment use the synchronous form; this one is kept for callers that want the statistics
    off the step stream and can afford the latency (e.g. gloo on CPU ranks)."""

    def __init__(self, parts, work):
        self._parts, self._work, self._out = parts, work, None

    def result(self) -> torch.Tensor:
        if self._out is None:
            if self._work is not None:
                self._work.wait()  # stream-level wait (NCCL) / host wait (gloo)
            self._out = combine_partials(torch.stack(self._parts) if isinstance(self._parts, list) else self._parts)
        return self._out


def all_reduce_partial_async(partial: torch.Tensor, group=None) -> PendingStats:
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return PendingStats(partial[None], None)
    world = dist.get_world_size(group)
    parts = [torch.empty(5, dtype=partial.dtype, device=partial.device) for _ in range(world)]
    work = dist.all_gather(parts, partial.contiguous(), group=group, async_op=True)
    return PendingStats(parts, work)


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

    def all_reduce_async(self, total_reward: torch.Tensor, group=None) -> PendingStats:
        """Same, without making the calling stream wait for the other ranks: the shard is reduced on the
        current stream (so a following `reset` may zero `total_reward`), the all_gather runs on the
        collective's stream and is joined by `PendingStats.result()`."""
        return all_reduce_partial_async(self.local(total_reward), group)
