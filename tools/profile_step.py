"""Run under `ncu --profile-from-start off ...`: warm the engine, then bracket ONE pipeline step (graphs off, so
every kernel is an ordinary launch) with cudaProfilerStart/Stop.  `--stage coarse` brackets only the 576-hypothesis
coarse forward (for the --set full capture)."""
import argparse
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from workloads import scenes  # noqa: E402
from megapose6d_b200 import _abi  # noqa: E402
from megapose6d_b200.tensor_collection import PandasTensorCollection  # noqa: E402
from megapose6d_b200.types import ObservationTensor  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--stage", default="step", choices=("step", "coarse"))
    args = ap.parse_args()
    sc = scenes.bench_scene(1)
    images, K, det_df, bboxes = sc["images"], sc["K"], sc["det_df"], sc["bboxes"]
    est = scenes.build_estimator(sc)
    _abi.lib().mpx_net_set_graphs(0)
    est.coarse_model.use_cuda_graphs = False
    est.refiner_model.use_cuda_graphs = False
    images_dev, K_dev, bboxes_dev = images.cuda(), K.cuda(), bboxes.cuda()

    def step():
        obs = ObservationTensor(images_dev, K_dev)
        det = PandasTensorCollection(det_df.copy(), bboxes=bboxes_dev)
        if args.stage == "coarse":
            return est.forward_coarse_model(obs, det)
        return est.run_inference_pipeline(obs, detections=det, n_refiner_iterations=5, n_pose_hypotheses=1)

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    step()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()


if __name__ == "__main__":
    main()
