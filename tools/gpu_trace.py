"""Kernel timeline of one pipeline step (torch.profiler / CUPTI): busy time, idle gaps and where they are."""
import sys
from pathlib import Path

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from workloads import scenes  # noqa: E402
from megapose6d_b200.tensor_collection import PandasTensorCollection  # noqa: E402
from megapose6d_b200.types import ObservationTensor  # noqa: E402


def main():
    sc = scenes.bench_scene(1)
    images, K, det_df, bboxes = sc["images"], sc["K"], sc["det_df"], sc["bboxes"]
    est = scenes.build_estimator(sc)
    images_dev, K_dev, bboxes_dev = images.cuda(), K.cuda(), bboxes.cuda()

    def step():
        obs = ObservationTensor(images_dev, K_dev)
        det = PandasTensorCollection(det_df.copy(), bboxes=bboxes_dev)
        return est.run_inference_pipeline(obs, detections=det, n_refiner_iterations=5, n_pose_hypotheses=1)

    for _ in range(4):
        step()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        step()
        torch.cuda.synchronize()
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    evs.sort(key=lambda e: e.time_range.start)
    t0, t1 = evs[0].time_range.start, max(e.time_range.end for e in evs)
    busy = sum(e.time_range.end - e.time_range.start for e in evs)
    print(f"{len(evs)} device activities, span {(t1 - t0) / 1e3:.3f} ms, busy {busy / 1e3:.3f} ms")
    gaps = []
    end = evs[0].time_range.end
    for prev, e in zip(evs, evs[1:]):
        g = e.time_range.start - end
        if g > 0:
            gaps.append((g, prev.name[:50], e.name[:50], (e.time_range.start - t0) / 1e3))
        end = max(end, e.time_range.end)
    print(f"idle total {sum(g[0] for g in gaps) / 1e3:.3f} ms in {len(gaps)} gaps")
    for g, a, b, at in sorted(gaps, reverse=True)[:25]:
        print(f"  {g:8.1f} us at {at:7.3f} ms   after {a:50s} before {b}")
    # coarse per-ms histogram of idle time
    bins = {}
    for g, a, b, at in gaps:
        bins[int(at)] = bins.get(int(at), 0) + g
    print("idle us per ms of the step:", {k: round(v) for k, v in sorted(bins.items())})


if __name__ == "__main__":
    main()
