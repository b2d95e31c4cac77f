"""Device timeline of frames in flight: CUDA events at the start / end of every frame's coarse stage (head) and at the end of
its tail, for the FramePipeline the bench runs.  Prints one row per frame (ms relative to the first head's start)."""
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from megapose6d_b200 import pose_estimator as PE  # noqa: E402
from megapose6d_b200.frame_pipeline import FramePipeline  # noqa: E402
from megapose6d_b200.tensor_collection import PandasTensorCollection  # noqa: E402
from megapose6d_b200.types import ObservationTensor  # noqa: E402
from workloads import scenes  # noqa: E402


def main():
    n_slots = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    gate = (sys.argv[2] != "overlap") if len(sys.argv) > 2 else True
    n_frames = 16
    sc = scenes.bench_scene(1)
    images, K, det_df = sc["images"].cuda(), sc["K"].cuda(), sc["det_df"]
    bboxes = sc["bboxes"].cuda()
    ests = [scenes.build_estimator(sc) for _ in range(n_slots)]
    kw = dict(n_refiner_iterations=5, n_pose_hypotheses=1)

    def frame():
        return ObservationTensor(images, K), PandasTensorCollection(det_df.copy(), bboxes=bboxes)

    for e in ests:
        for _ in range(5):
            e.run_inference_pipeline(*frame()[:1], detections=frame()[1], **kw)
    torch.cuda.synchronize()
    marks = []
    orig = PE.PoseEstimator._coarse_stage_graphed

    def wrapped(self, *a, **k):
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        r = orig(self, *a, **k)
        e1.record()
        marks.append([e0, e1, None])
        return r

    PE.PoseEstimator._coarse_stage_graphed = wrapped
    pipe = FramePipeline(None, estimators=ests, serialize_heads=gate)
    for _ in range(4):
        pipe.submit(*frame(), **kw)
    pipe.drain()
    torch.cuda.synchronize()
    marks.clear()
    for i in range(n_frames):
        pipe.submit(*frame(), **kw)
        slot = pipe.slots[(pipe._next - 1) % len(pipe.slots)]
        e2 = torch.cuda.Event(enable_timing=True)
        e2.record(slot["stream"])
        marks[-1][2] = e2
    pipe.drain()
    torch.cuda.synchronize()
    t0 = marks[0][0]
    rows = []
    for i, (e0, e1, e2) in enumerate(marks):
        rows.append(dict(frame=i, head_start=t0.elapsed_time(e0), head_end=t0.elapsed_time(e1), tail_end=t0.elapsed_time(e2)))
    for a, b in zip(rows, rows[1:] + [None]):
        a["head_ms"] = a["head_end"] - a["head_start"]
        a["tail_ms"] = a["tail_end"] - a["head_end"]
        a["gap_to_next_head"] = (b["head_start"] - a["head_end"]) if b else None
    print(json.dumps(dict(slots=n_slots, gated=gate, frame_ms=(rows[-1]["head_start"] - rows[2]["head_start"]) / (len(rows) - 3),
                          rows=[{k: (round(v, 3) if isinstance(v, float) else v) for k, v in r.items()} for r in rows])))


if __name__ == "__main__":
    main()
