"""Hypothesis sharding over the GPUs of one box.

The reference never splits hypotheses across devices (it shards *frames* over ranks for evaluation and
exchanges predictions through files + barriers: src/megapose/datasets/samplers.py:41-55,
utils/tensor_collection.py:165-186).  Rows of every stage are independent
(inference/pose_estimator.py:139-141, 362-364 only chunk them), so here rank r owns a contiguous
slice of the stage's rows and one small all-gather per stage (logits after coarse/scoring, poses after
the refiner) makes the result identical on every rank.  NCCL collectives are enqueued from
torch.distributed without a host synchronisation; with the gloo backend (CPU tests) the same code
path runs on host tensors.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist


class HypothesisSharder:
    def __init__(self, group: Optional["dist.ProcessGroup"] = None, enabled: Optional[bool] = None):
        if enabled is None:
            enabled = dist.is_available() and dist.is_initialized()
        self.enabled = bool(enabled)
        self.group = group
        self.rank = dist.get_rank(group) if self.enabled else 0
        self.world = dist.get_world_size(group) if self.enabled else 1

    def span(self, n: int) -> Tuple[int, int]:
        """Contiguous slice [start, end) of n rows owned by this rank (ceil split, last ranks may be empty)."""
        per = (n + self.world - 1) // self.world if n > 0 else 0
        start = min(n, self.rank * per)
        return start, min(n, start + per)

    def gather_rows(self, local: torch.Tensor, n_total: int) -> torch.Tensor:
        """All-gather row slices produced by `span` into the full [n_total, ...] tensor on every rank."""
        if self.world == 1:
            return local
        per = (n_total + self.world - 1) // self.world
        tail = local.shape[1:]
        padded = torch.zeros((per,) + tuple(tail), dtype=local.dtype, device=local.device)
        padded[: local.shape[0]] = local
        out = torch.empty((self.world * per,) + tuple(tail), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, padded.contiguous(), group=self.group)
        return out[:n_total]
