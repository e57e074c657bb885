"""Sharding of the attribution workload over ranks (one process per GPU).

The units are (image, guess) trajectories; they share only read-only state and never communicate
while they run (src/main.py:48-81).  The single exchange of the whole job is the gather of each
trajectory's final ``[loss, key logits, alpha]`` row, after which rank 0 reproduces the reference's
per-image ``argmin`` over the guesses (src/main.py:84-88).  NCCL on GPUs, gloo in the CPU tests.
"""
from __future__ import annotations

from typing import List, Tuple

import torch
import torch.distributed as dist


def trajectory_list(sample_size: int, n_guesses: int) -> List[Tuple[int, int]]:
    """Flat list of (image, guess) pairs in the reference's loop order (images outer, guesses inner)."""
    return [(i, g) for i in range(sample_size) for g in range(n_guesses)]


def partition(total: int, rank: int, world: int) -> range:
    """Contiguous, balanced slice of ``range(total)`` owned by ``rank`` (sizes differ by at most one)."""
    base, extra = divmod(total, world)
    start = rank * base + min(rank, extra)
    return range(start, start + base + (1 if rank < extra else 0))


def batches(indices: range, batch: int) -> List[range]:
    return [range(s, min(s + batch, indices.stop)) for s in range(indices.start, indices.stop, batch)]


def gather_rows(local: torch.Tensor, total: int, rank: int, world: int) -> torch.Tensor:
    """All-gather the per-trajectory result rows ``[len(partition), width]`` into ``[total, width]``
    (same order as ``trajectory_list``).  Uneven shards are padded to the largest shard."""
    if world == 1:
        return local
    width = local.shape[1]
    longest = (total + world - 1) // world
    buf = local.new_zeros(longest, width)
    buf[: local.shape[0]] = local
    parts = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(parts, buf)
    return torch.cat([parts[r][: len(partition(total, r, world))] for r in range(world)], 0)


def select_best(rows: torch.Tensor, n_guesses: int, key_len: int):
    """Per image: index of the guess with the smallest final loss, its key logits and alpha
    (src/main.py:84-88; ``list.index(min(...))`` = first minimum, as ``argmin`` on ties here)."""
    per_img = rows.reshape(-1, n_guesses, rows.shape[1])
    losses = per_img[:, :, 0]
    best = torch.argmin(losses, dim=1)
    picked = per_img[torch.arange(per_img.shape[0]), best]
    return best, picked[:, 1:1 + key_len], picked[:, 1 + key_len:]


def bit_accuracy(key_logits: torch.Tensor, true_key: torch.Tensor) -> torch.Tensor:
    """``mean(round(sigmoid(key)) == key_true)`` per image (src/utils.py:37-41, src/main.py:88)."""
    return (torch.round(torch.sigmoid(key_logits)) == true_key.to(key_logits.dtype)).float().mean(1)
