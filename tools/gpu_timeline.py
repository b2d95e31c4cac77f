"""Host vs device time per pipeline stage (async run, CUDA events at the stage boundaries + host clocks)."""
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from workloads import scenes  # noqa: E402
from megapose6d_b200 import pose_estimator  # noqa: E402
from megapose6d_b200.tensor_collection import PandasTensorCollection  # noqa: E402
from megapose6d_b200.types import ObservationTensor  # noqa: E402

sc = scenes.bench_scene(1)
images, K, det_df, bboxes = sc["images"], sc["K"], sc["det_df"], sc["bboxes"]
est = scenes.build_estimator(sc)
images_dev, K_dev, bboxes_dev = images.cuda(), K.cuda(), bboxes.cuda()
marks = []


def wrap(name):
    fn = getattr(pose_estimator.PoseEstimator, name)

    def inner(self, *a, **k):
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        t0 = time.perf_counter()
        r = fn(self, *a, **k)
        t1 = time.perf_counter()
        e1.record()
        marks.append((name, t0, t1, e0, e1))
        return r

    setattr(pose_estimator.PoseEstimator, name, inner)


for n in ("forward_coarse_model", "forward_refiner", "forward_scoring_model", "filter_pose_estimates"):
    wrap(n)


def step():
    obs = ObservationTensor(images_dev, K_dev)
    det = PandasTensorCollection(det_df.copy(), bboxes=bboxes_dev)
    return est.run_inference_pipeline(obs, detections=det, n_refiner_iterations=5, n_pose_hypotheses=1)


for _ in range(4):
    step()
torch.cuda.synchronize()
marks.clear()
N = 6
t_start = time.perf_counter()
e_start = torch.cuda.Event(enable_timing=True)
e_end = torch.cuda.Event(enable_timing=True)
e_start.record()
for _ in range(N):
    step()
e_end.record()
torch.cuda.synchronize()
wall = (time.perf_counter() - t_start) / N * 1e3
print(f"step: wall {wall:.2f} ms, device span {e_start.elapsed_time(e_end) / N:.2f} ms")
agg = {}
for name, t0, t1, e0, e1 in marks:
    h, d = agg.get(name, (0.0, 0.0))
    agg[name] = (h + (t1 - t0) * 1e3, d + e0.elapsed_time(e1))
tot_h = 0
for name, (h, d) in agg.items():
    print(f"  {name:28s} host {h / N:7.3f} ms   device span {d / N:7.3f} ms")
    tot_h += h / N
print(f"  outside the four stages: host {wall - tot_h:.3f} ms")
