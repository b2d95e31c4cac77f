"""Two frames in flight on one GPU.

One frame of the pipeline (inference/pose_estimator.py:511-641) is a throughput-bound head -- rasterising and scoring
576 hypotheses per detection, ~9 ms of kernels that fill the device -- followed by a latency-bound tail: five refiner
iterations and the scoring pass on ONE hypothesis per detection, ~300 small dependent launches that together keep a few
dozen SMs busy for ~3 ms.  The reference's evaluation loop (evaluation/prediction_runner.py:156-209) feeds frames one
after the other, so the tail of frame i and the head of frame i+1 never overlap.

FramePipeline alternates frames over `n_slots` PoseEstimators, each with its own CUDA stream, graphs, workspaces and mesh
database (nothing on the device is shared, so no two frames ever touch the same buffer), using
PoseEstimator.submit_inference_pipeline: the tail of one frame runs beside the head of the next, the heads themselves are
gated one behind the other (`serialize_heads`), and the host-side bookkeeping of a frame runs while the device works on
the next.  `set_reserved_sms` sizes the persistent grids for fewer than all SMs (measurement aid: it did not pay off,
DESIGN.md section 3.3).  Results come back in submission order and are, frame for frame, bit-identical to run_inference_pipeline's (same kernels,
same launch order within a frame).

    pipe = FramePipeline(lambda: build_estimator(...), n_slots=2)
    for obs, det in frames:
        done = pipe.submit(obs, det, n_refiner_iterations=5, n_pose_hypotheses=1)   # None until the pipe is full
        if done is not None: consume(*done)
    for done in pipe.drain(): consume(*done)
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple

import torch

from . import _abi

DEFAULT_RESERVE_SMS = 16


def set_reserved_sms(reserve: int) -> int:
    """Size every persistent grid for (device SMs - reserve) SMs; 0 restores the full device.  Process-wide, and recorded
    by mesh databases and captured graphs: call it before the estimators are built.  Returns the SM count now in use."""
    lib = _abi.lib()
    _abi.check(lib.mpx_set_sm_limit(0))
    total = lib.mpx_sm_count()
    if reserve > 0:
        limit = max(16, (total - reserve) // 2 * 2)
        _abi.check(lib.mpx_set_sm_limit(limit))
    return lib.mpx_sm_count()


class FramePipeline:
    def __init__(self, make_estimator: Callable[[], "torch.nn.Module"], n_slots: int = 2,
                 estimators: Optional[Sequence["torch.nn.Module"]] = None, device=None, tail_priority="high",
                 serialize_heads: bool = True):
        """`make_estimator` is called once per slot (each call must build its own models and mesh database); or pass the
        estimators themselves (fresh ones: `tail_priority` -- PoseEstimator.set_tail_priority -- is recorded by the graphs
        they capture on their first frames)."""
        ests = list(estimators) if estimators is not None else [make_estimator() for _ in range(n_slots)]
        for e in ests:
            e.set_tail_priority(tail_priority)
        assert len(ests) >= 1 and len({id(e) for e in ests}) == len(ests)
        for a in ests:
            for b in ests:
                if a is not b:
                    assert a.coarse_model is not b.coarse_model and a.refiner_model is not b.refiner_model, \
                        "the slots of a FramePipeline must not share models (their buffers and graphs are single-buffered)"
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        prio = -1 if tail_priority == "low" else 0
        self.slots = [dict(est=e, stream=torch.cuda.Stream(device=self.device, priority=prio), pending=None) for e in ests]
        self._next = 0
        # `serialize_heads`: frame i+1's coarse stage waits (on the device) for frame i's.  Without the gate the head of frame
        # i+2 -- enqueued behind the short tail of frame i on the same stream -- starts in the middle of frame i+1's head:
        # two sets of persistent, statically scheduled kernels then take SMs from each other at every kernel boundary and
        # each finishes when its most delayed CTA does.  Gated, the heads run back to back and each overlaps only the
        # latency-bound tail of its predecessor, which is what the pipeline is for.
        self.serialize_heads = bool(serialize_heads)
        self._last_head_done = None

    def __len__(self) -> int:
        return len(self.slots)

    def submit(self, observation, detections, **kwargs) -> Optional[Tuple]:
        """Enqueue one frame.  Returns the (final, extra_data) of the frame that occupied the slot before (the oldest frame
        in flight), or None while the pipe is filling."""
        slot = self.slots[self._next]
        self._next = (self._next + 1) % len(self.slots)
        old = slot["pending"]
        caller = torch.cuda.current_stream(self.device)
        slot["stream"].wait_stream(caller)  # inputs produced (copied) on the caller's stream
        with torch.cuda.stream(slot["stream"]):
            # enqueued BEFORE the older frame of this slot is waited for: the stream orders the two on the device, and the
            # host-side bookkeeping of the older frame (a few ms of pandas) then runs while the device has work queued
            if self.serialize_heads and self._last_head_done is not None and len(self.slots) > 1:
                slot["est"].__dict__["_head_gate"] = self._last_head_done
            slot["pending"] = slot["est"].submit_inference_pipeline(observation, detections, **kwargs)
            self._last_head_done = slot["est"].__dict__.pop("_head_done", None)
        return old.result() if old is not None else None

    def drain(self) -> List[Tuple]:
        """Wait for every frame in flight; results in submission order."""
        out = []
        for k in range(len(self.slots)):
            slot = self.slots[(self._next + k) % len(self.slots)]
            if slot["pending"] is not None:
                out.append(slot["pending"].result())
                slot["pending"] = None
        return out

    def join(self) -> None:
        """Make the caller's current stream wait for everything enqueued so far (for device-side timing)."""
        caller = torch.cuda.current_stream(self.device)
        for slot in self.slots:
            caller.wait_stream(slot["stream"])

    def run(self, frames, **kwargs):
        """Generator over (final, extra_data) of `frames` (an iterable of (observation, detections)), in order."""
        for obs, det in frames:
            done = self.submit(obs, det, **kwargs)
            if done is not None:
                yield done
        yield from self.drain()
