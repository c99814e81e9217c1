"""Multi-GPU plumbing of the correspondence path: one process per GPU, torch.distributed.

The path shards by (scene, view) with no data-path communication (the reference already treats
every view independently: dataset/data_loader_infer.py:161, run/infer.py:428).  The single
exchange step exists only when ONE scene's views are split over ranks and a scene-level mask set
is pooled across all of them: every rank pools its own views into per-mask sums [K,C] and counts
[K], and one all-reduce(SUM) of K*(C+1) numbers combines them (308 KB at K=50, C=768 — latency
bound on NVLink; NCCL picks NVLS/in-switch reduction when available).  The analogous cross-view
reduction of the reference is the CPU vote histogram of run/infer.py:642-647.

All helpers work on CUDA tensors over `nccl` and on CPU tensors over `gloo` (used by the tests).
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def world() -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_scenes(weights: Sequence[float], world_size: int) -> List[List[int]]:
    """Size-balanced assignment of scenes to ranks (longest-processing-time greedy on the work
    estimate, e.g. N_scene * V_scene).  Deterministic: ties go to the lower rank / scene id."""
    order = sorted(range(len(weights)), key=lambda i: (-float(weights[i]), i))
    load = [0.0] * world_size
    out: List[List[int]] = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda j: (load[j], j))
        out[r].append(i)
        load[r] += float(weights[i])
    return [sorted(x) for x in out]


def shard_views(n_views: int, world_size: int, rank: int) -> range:
    """Contiguous block partition of one scene's views (sizes differ by at most one)."""
    base, rem = divmod(n_views, world_size)
    start = rank * base + min(rank, rem)
    return range(start, start + base + (1 if rank < rem else 0))


def allreduce_mask_sums(sums: torch.Tensor, counts: torch.Tensor, group=None,
                        async_op: bool = False):
    """sums [..., K, C] (per-view or already view-reduced) and counts [..., K] of THIS rank ->
    (scene sums float64 [K,C], scene counts int64 [K]) over all ranks.  Leading dimensions (views)
    are reduced locally in float64 first; the exchange is ONE all-reduce of a packed float64
    [K, C+1] buffer (counts are exact in float64 up to 2^53).  With async_op the (work, finish)
    pair lets the caller overlap the exchange with the next view's pooling."""
    k, c = sums.shape[-2], sums.shape[-1]
    s = sums.reshape(-1, k, c).to(torch.float64).sum(0)
    n = counts.reshape(-1, k).to(torch.float64).sum(0)
    packed = torch.cat([s, n.unsqueeze(1)], dim=1).contiguous()
    work = None
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        work = dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=group, async_op=async_op)

    def finish():
        if work is not None and async_op:
            work.wait()
        return packed[:, :c], packed[:, c].round().to(torch.int64)
    if async_op:
        return work, finish
    return finish()


def allreduce_votes(votes: torch.Tensor, counter: torch.Tensor, group=None):
    """Cross-view vote histogram of run/infer.py:642-647 when one scene's views live on several ranks: every
    rank accumulates `votes int32 [N, T]` / `counter int32 [N]` over ITS views (ops.accumulate_votes), one
    all-reduce(SUM) of the packed int32 [N, T+1] buffer gives every rank the scene's totals (integers: exact
    and order independent), after which ops.vote_argmax / ops.nn_fill run as on one GPU.  In place."""
    if not (dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1):
        return votes, counter
    packed = torch.cat([votes, counter.unsqueeze(1)], dim=1).contiguous()
    dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=group)
    votes.copy_(packed[:, :-1])
    counter.copy_(packed[:, -1])
    return votes, counter


def finalize_mean(sums: torch.Tensor, counts: torch.Tensor) -> torch.Tensor:
    """mean[k,:] = sums[k,:] / counts[k] (zeros for empty masks), float32 like torch.mean of the
    reference's float32 features (models/utils/criterion.py:152-157)."""
    den = counts.clamp(min=1).to(sums.dtype).unsqueeze(-1)
    mean = sums / den
    mean[counts == 0] = 0
    return mean.to(torch.float32)
