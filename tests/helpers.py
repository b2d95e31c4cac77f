"""Shared builders for the tests: synthetic scenes, oracle objects, reference adapters."""
from __future__ import annotations

import numpy as np
import torch

from megapose6d_b200 import procedural
from oracle import pipeline_ref, resnet_ref


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


def ref_meshes_from_dataset(ds) -> pipeline_ref.RefMeshes:
    labels, v, n, c, f = [], [], [], [], []
    for obj in ds.list_objects:
        m = obj.mesh.with_defaults()
        labels.append(obj.label)
        v.append(np.asarray(m.vertices, np.float64) * obj.scale)
        n.append(m.vertex_normals)
        c.append(np.clip(m.vertex_colors, 0, 1))
        f.append(m.faces)
    return pipeline_ref.RefMeshes(labels, v, n, c, f)


COARSE_CFG = dict(n_rendered_views=1, multiview_type="TCO", render_normals=True, render_depth=False, input_depth=False,
                  predict_rendered_views_logits=True, predict_pose_update=False, remove_TCO_rendering=False,
                  depth_normalization_type="tCR_scale_clamp_center")
REFINER_CFG = dict(n_rendered_views=4, multiview_type="TCO+front_3views", render_normals=True, render_depth=False,
                   input_depth=False, predict_rendered_views_logits=False, predict_pose_update=True,
                   remove_TCO_rendering=False, depth_normalization_type="tCR_scale_clamp_center")
REFINER_RGBD_CFG = dict(REFINER_CFG, render_depth=True, input_depth=True)


def n_inputs(cfg):
    return (3 + int(cfg["input_depth"])) + (6 + int(cfg["render_depth"])) * cfg["n_rendered_views"]


def make_state_dict(cfg, seed=0):
    head = "pose_fc" if cfg["predict_pose_update"] else "views_logits_head"
    dim = 9 if cfg["predict_pose_update"] else cfg["n_rendered_views"]
    sd = resnet_ref.init_state_dict(n_inputs(cfg), head, dim, seed=seed)
    if cfg["predict_pose_update"]:
        # keep random-weight pose updates small and well-conditioned: R ~ I, vz ~ 1
        sd["pose_fc.weight"] = sd["pose_fc.weight"] * 0.05
        sd["pose_fc.bias"] = torch.tensor([1.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0, 1.0]) + 0.02 * sd["pose_fc.bias"]
    return sd


def detection_for_pose(K, TCO, points, pad=4.0):
    """Axis-aligned bbox (x1,y1,x2,y2) of the projected points."""
    P = (TCO[:3, :3] @ points.T + TCO[:3, 3:4])
    uv = (K @ P)
    uv = uv[:2] / uv[2:]
    return torch.tensor([uv[0].min() - pad, uv[1].min() - pad, uv[0].max() + pad, uv[1].max() + pad])
