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
        self._pad_bufs: dict = {}

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
        tail = tuple(local.shape[1:])
        out = torch.empty((self.world * per,) + tail, dtype=local.dtype, device=local.device)
        if per * self.world == n_total:
            # even split (576 rows over 1/2/4/8 ranks, one refiner row per rank, ...): one collective, nothing else -- the
            # stage's three gathers are latency-bound, a memset + copy + slice per call was a third of their cost
            dist.all_gather_into_tensor(out, local.contiguous(), group=self.group)
            return out
        key = (per, tail, local.dtype, local.device)
        padded = self._pad_bufs.get(key)
        if padded is None:
            if len(self._pad_bufs) >= 16:
                self._pad_bufs.clear()
            padded = self._pad_bufs[key] = torch.zeros((per,) + tail, dtype=local.dtype, device=local.device)
        padded[: local.shape[0]] = local  # rows past the slice keep their zeros (a shorter slice is always the last one)
        if local.shape[0] < per:
            padded[local.shape[0]:] = 0
        dist.all_gather_into_tensor(out, padded, group=self.group)
        return out[:n_total]
