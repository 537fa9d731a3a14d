"""Frame sharding across the GPUs of one box: the only multi-GPU logic on this path.

The backbone couples nothing across frames, so rank r of W simply takes frames r, r+W, ... exactly like the reference's
test-time sampler (pcdet/datasets/__init__.py:31-52: indices padded by wrap-around to a multiple of W, then
`indices[rank:total_size:num_replicas]`), runs them through its own engine, and the per-frame results are merged back in
dataset order after the timed region (pcdet/utils/common_utils.py:229-250 merge_results_dist: interleave the per-rank
lists, truncate to the dataset size).  There is no collective on the data path; the merge is one all_gather of a padded
fixed-shape tensor per rank (NCCL on GPUs, gloo in the CPU tests) instead of the reference's pickle files + barriers."""
from __future__ import annotations

from typing import List, Sequence

import torch
import torch.distributed as dist


def frames_for_rank(n_frames: int, rank: int, world: int) -> List[int]:
    """DistributedSampler(shuffle=False).__iter__ (pcdet/datasets/__init__.py:40-50)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    per = (n_frames + world - 1) // world
    total = per * world
    idx = list(range(n_frames))
    idx += idx[:total - n_frames]
    return idx[rank:total:world]


def merge_order(n_frames: int, world: int) -> List[tuple]:
    """(rank, local index) of dataset frame i after the interleaved merge (common_utils.py:244-248)."""
    return [(i % world, i // world) for i in range(n_frames)]


def gather_frame_results(local: torch.Tensor, n_frames: int) -> torch.Tensor:
    """local: [frames_on_this_rank, ...] fixed-shape per-frame results (padded detections, counts, ...), the same shape
    on every rank.  Returns [n_frames, ...] in dataset order on every rank.  Single process: returns local[:n_frames]."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local[:n_frames]
    world = dist.get_world_size()
    parts = [torch.empty_like(local) for _ in range(world)]
    dist.all_gather(parts, local.contiguous())
    stacked = torch.stack(parts, dim=1)                       # [local index, rank, ...] == interleaved dataset order
    return stacked.reshape((-1,) + tuple(local.shape[1:]))[:n_frames]


def aggregate_rate(units_per_rank: Sequence[int], seconds_max_over_ranks: float) -> float:
    """Whole-job throughput: all ranks' units over the slowest rank's device time."""
    return float(sum(units_per_rank)) / seconds_max_over_ranks
