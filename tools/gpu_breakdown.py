"""Where does a pipeline step go?  Wraps the engine's entry points with synchronising timers (diagnostic only:
the synchronisation removes all host/device overlap, so the parts add up to more than the asynchronous step)."""
import sys
import time
from collections import defaultdict
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from workloads import scenes  # noqa: E402
from megapose6d_b200 import _abi, backbone, lib3d, load_model, pose_estimator, pose_predictor, renderer  # noqa: E402
from megapose6d_b200.tensor_collection import PandasTensorCollection  # noqa: E402
from megapose6d_b200.types import ObservationTensor  # noqa: E402

acc = defaultdict(float)
cnt = defaultdict(int)


def wrap(obj, name, label=None):
    fn = getattr(obj, name)
    label = label or name

    def inner(*a, **k):
        torch.cuda.synchronize()
        t = time.perf_counter()
        r = fn(*a, **k)
        torch.cuda.synchronize()
        acc[label] += time.perf_counter() - t
        cnt[label] += 1
        return r

    setattr(obj, name, inner)


def main():
    sc = scenes.bench_scene(1)
    images, K, det_df, bboxes = sc["images"], sc["K"], sc["det_df"], sc["bboxes"]
    est = scenes.build_estimator(sc)
    images_dev, K_dev, bboxes_dev = images.cuda(), K.cuda(), bboxes.cuda()

    def step():
        obs = ObservationTensor(images_dev, K_dev)
        det = PandasTensorCollection(det_df.copy(), bboxes=bboxes_dev)
        return est.run_inference_pipeline(obs, detections=det, n_refiner_iterations=5, n_pose_hypotheses=1)

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(5):
        step()
    torch.cuda.synchronize()
    print(f"async step: {(time.perf_counter() - t) / 5 * 1e3:.2f} ms")

    for name in ("crop_geometry", "normalize_T", "make_TCO_multiview", "update_pose", "TCO_init_from_boxes_autodepth_with_R",
                 "image_to_nhwc4"):
        wrap(lib3d, name)
    wrap(renderer.BatchRenderer, "render_fused")
    wrap(backbone.ResNet34Engine, "forward", "net_forward")
    wrap(backbone.ResNet34Engine, "alloc_input")
    lib = _abi.lib()
    orig = lib.mpx_roi_align_fused

    class Roi:
        def __call__(self, *a):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            r = orig(*a)
            torch.cuda.synchronize()
            acc["roi_align_fused"] += time.perf_counter() - t0
            cnt["roi_align_fused"] += 1
            return r

    lib.mpx_roi_align_fused = Roi()
    for name in ("forward_coarse_model", "forward_refiner", "forward_scoring_model", "filter_pose_estimates"):
        wrap(pose_estimator.PoseEstimator, name, "STAGE " + name)
    n = 5
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(n):
        step()
    torch.cuda.synchronize()
    total = (time.perf_counter() - t) / n
    print(f"synchronised step: {total * 1e3:.2f} ms")
    for k in sorted(acc, key=lambda k: -acc[k]):
        print(f"  {k:45s} {acc[k] / n * 1e3:8.3f} ms/step  ({cnt[k] // n} calls)")


if __name__ == "__main__":
    main()
