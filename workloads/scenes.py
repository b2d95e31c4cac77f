"""Synthetic scenes and detections (no dataset is available offline): the inputs of bench.py, tools/ and tests/.

Every scenario is a dict with the same keys -- ds (RigidObjectDataset), images [1,C,480,640], K [1,3,3], labels,
bboxes [B,4], det_df, sd_coarse / sd_refiner (reference-format state dicts), cfg_refiner, grid, n_refiner_iterations,
n_pose_hypotheses, render_size -- so that the real reference (tools/make_golden.py), the CPU oracle and the CUDA path
can be driven from one description.  BASELINE.json configs[k] -> `baseline_config(k)`.
"""
from __future__ import annotations

import numpy as np
import torch

from megapose6d_b200 import procedural

from .weights import COARSE_CFG, REFINER_CFG, REFINER_RGBD_CFG, make_state_dict


def make_scene(n_objects=1, seed=0, h=480, w=640, with_depth=False, n_seg=100, n_lat=51):
    ds = procedural.make_object_dataset(n_objects, seed=seed, n_seg=n_seg, n_lat=n_lat)
    rng = np.random.RandomState(seed)
    rgb = torch.from_numpy(rng.randint(0, 256, size=(1, 3, h, w)).astype(np.float32) / 255.0)
    # low-pass the noise a little so that bilinear crops are not pure noise
    rgb = torch.nn.functional.avg_pool2d(rgb, 5, stride=1, padding=2)
    if with_depth:
        depth = torch.from_numpy(rng.uniform(0.3, 1.5, size=(1, 1, h, w)).astype(np.float32))
        depth[:, :, ::7, ::5] = 0.0  # invalid pixels
        images = torch.cat([rgb, depth], dim=1)
    else:
        images = rgb
    K = torch.from_numpy(procedural.example_camera(h, w)).float().unsqueeze(0)
    return ds, images.contiguous(), K


def detection_for_pose(K, TCO, points, pad=4.0):
    """Axis-aligned bbox (x1,y1,x2,y2) of the projected points."""
    P = (TCO[:3, :3] @ points.T + TCO[:3, 3:4])
    uv = (K @ P)
    uv = uv[:2] / uv[2:]
    return torch.tensor([uv[0].min() - pad, uv[1].min() - pad, uv[0].max() + pad, uv[1].max() + pad])



def _detections(ds, K, poses):
    import pandas as pd

    n = len(poses)
    labels = [ds[i % len(ds)].label for i in range(n)]
    bboxes = torch.stack([detection_for_pose(K[0], poses[i], torch.from_numpy(ds[i % len(ds)].mesh.vertices).float())
                          for i in range(n)])
    det_df = pd.DataFrame(dict(label=labels, batch_im_id=0))
    det_df["instance_id"] = det_df.groupby("label").cumcount().values
    return labels, bboxes, det_df


def pipeline_scenario():
    """Two objects, one RGB frame, one detection each; 72-rotation grid, 2 hypotheses, 2 refiner iterations.  The same
    scenario is run by the real reference (tools/make_golden.py -> tests/golden/pipeline.npz, and live in
    tests/test_oracle_vs_reference.py), by the oracle (tests/test_oracle_golden.py) and by the CUDA path
    (tests/test_gpu_pipeline.py)."""
    import pandas as pd

    ds, images, K = make_scene(2, seed=6)
    labels = [o.label for o in ds.list_objects]
    TCO_gt = torch.from_numpy(procedural.random_poses(2, 11)).float()
    TCO_gt[:, 2, 3] = torch.tensor([0.55, 0.7])
    bboxes = torch.stack([detection_for_pose(K[0], TCO_gt[i], torch.from_numpy(ds[i].mesh.vertices).float()) for i in range(2)])
    det_df = pd.DataFrame(dict(label=labels, batch_im_id=0, instance_id=np.arange(2)))
    return dict(ds=ds, images=images, K=K, labels=labels, bboxes=bboxes, det_df=det_df,
                sd_coarse=make_state_dict(COARSE_CFG, 5), sd_refiner=make_state_dict(REFINER_CFG, 6),
                grid=72, n_refiner_iterations=2, n_pose_hypotheses=2)



def bench_scene(n_objects: int, grid: int = 576, n_refiner_iterations: int = 5, render_size=(240, 320)):
    """BASELINE configs[1] per detection: one object x `grid` coarse hypotheses + 5 refiner iterations + scoring (RGB).
    bench.py runs it with one detection per GPU; tools/make_golden.py pins n_objects = 1 against the real reference
    (tests/golden/fullsize_rgb*.npz)."""
    ds, images, K = make_scene(n_objects, seed=0)
    poses = torch.from_numpy(procedural.random_poses(n_objects, 5, z_range=(0.5, 0.9), xy_range=0.1)).float()
    labels, bboxes, det_df = _detections(ds, K, poses)
    return dict(ds=ds, images=images, K=K, labels=labels, bboxes=bboxes, det_df=det_df, TCO_gt=poses,
                sd_coarse=make_state_dict(COARSE_CFG, 1), sd_refiner=make_state_dict(REFINER_CFG, 2),
                cfg_refiner=REFINER_CFG, model="megapose-1.0-RGB", grid=grid, n_refiner_iterations=n_refiner_iterations,
                n_pose_hypotheses=1, render_size=tuple(render_size))


def rgbd_scene(n_objects: int = 32, grid: int = 72, n_refiner_iterations: int = 5, seed: int = 21):
    """BASELINE configs[2]: RGB-D refiner (depth branch) on a 32-object batch x 5 refiner iterations; the coarse model
    stays RGB as in the zoo (utils/load_model.py:18-26).  32 distinct meshes of 5k-20k triangles; the depth channel is
    uniform noise with invalid (zero) pixels, which exercises the validity masking of the crop."""
    rng = np.random.RandomState(seed)
    sizes = [(int(rng.randint(50, 101)), int(rng.randint(51, 101))) for _ in range(n_objects)]  # 2*seg*(lat-1) triangles
    ds = procedural.make_object_dataset(n_objects, seed=seed, n_seg=[s[0] for s in sizes], n_lat=[s[1] for s in sizes])
    _, images, K = make_scene(1, seed=seed, with_depth=True)
    poses = torch.from_numpy(procedural.random_poses(n_objects, seed + 1, z_range=(0.45, 0.9), xy_range=0.12)).float()
    labels, bboxes, det_df = _detections(ds, K, poses)
    return dict(ds=ds, images=images, K=K, labels=labels, bboxes=bboxes, det_df=det_df, TCO_gt=poses,
                sd_coarse=make_state_dict(COARSE_CFG, 3), sd_refiner=make_state_dict(REFINER_RGBD_CFG, 4),
                cfg_refiner=REFINER_RGBD_CFG, model="megapose-1.0-RGBD", grid=grid,
                n_refiner_iterations=n_refiner_iterations, n_pose_hypotheses=1, render_size=(240, 320))


def ycbv_scene(n_objects: int = 21, grid: int = 576, n_refiner_iterations: int = 5, seed: int = 31):
    """BASELINE configs[3]: a YCB-V-style frame with 21 objects x 576 hypotheses (12 096 coarse rows)."""
    ds, images, K = make_scene(n_objects, seed=seed)
    poses = torch.from_numpy(procedural.random_poses(n_objects, seed + 1, z_range=(0.5, 1.0), xy_range=0.15)).float()
    labels, bboxes, det_df = _detections(ds, K, poses)
    return dict(ds=ds, images=images, K=K, labels=labels, bboxes=bboxes, det_df=det_df, TCO_gt=poses,
                sd_coarse=make_state_dict(COARSE_CFG, 1), sd_refiner=make_state_dict(REFINER_CFG, 2),
                cfg_refiner=REFINER_CFG, model="megapose-1.0-RGB", grid=grid, n_refiner_iterations=n_refiner_iterations,
                n_pose_hypotheses=1, render_size=(240, 320))


FULLSIZE = {
    # name -> builder; each has a fixture tests/golden/<name>.npz written by tools/make_golden.py from the REAL reference
    "fullsize_rgb": lambda: bench_scene(1),
    "fullsize_rgb_224": lambda: bench_scene(1, render_size=(224, 224)),
    "fullsize_rgbd32": lambda: rgbd_scene(32),
    "fullsize_ycbv21": lambda: ycbv_scene(21),
}


def build_estimator(sc: dict, sharder=None):
    """The product's PoseEstimator for a scenario: the weights go through the model-zoo directory layout
    (<root>/<run_id>/{config.yaml, checkpoint.pth.tar}) and `load_named_model`, like a downloaded checkpoint would."""
    import tempfile
    from pathlib import Path

    from megapose6d_b200 import load_model

    named = load_model.NAMED_MODELS[sc["model"]]
    with tempfile.TemporaryDirectory() as tmp:
        load_model.write_run(tmp, named["coarse_run_id"], sc["sd_coarse"])
        load_model.write_run(tmp, named["refiner_run_id"], sc["sd_refiner"])
        est = load_model.load_named_model(sc["model"], sc["ds"], models_root=Path(tmp), render_size=sc["render_size"])
    if sc["grid"] != 576:
        est.load_SO3_grid(sc["grid"])
    if sharder is not None:
        est.sharder = sharder
    return est
