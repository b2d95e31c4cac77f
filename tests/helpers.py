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
    labels, v, n, c, f, uv, tex, mod = [], [], [], [], [], [], [], []
    for obj in ds.list_objects:
        m = obj.mesh.with_defaults()
        labels.append(obj.label)
        v.append(np.asarray(m.vertices, np.float64) * obj.scale)
        n.append(m.vertex_normals)
        c.append(np.clip(m.vertex_colors, 0, 1))
        f.append(m.faces)
        uv.append(m.uv)
        tex.append(m.texture)
        mod.append(1 if m.texture_modulate else 0)
    return pipeline_ref.RefMeshes(labels, v, n, c, f, uv, tex, mod)


COARSE_CFG = dict(n_rendered_views=1, multiview_type="TCO", render_normals=True, render_depth=False, input_depth=False,
                  predict_rendered_views_logits=True, predict_pose_update=False, remove_TCO_rendering=False,
                  depth_normalization_type="tCR_scale_clamp_center")
REFINER_CFG = dict(n_rendered_views=4, multiview_type="TCO+front_3views", render_normals=True, render_depth=False,
                   input_depth=False, predict_rendered_views_logits=False, predict_pose_update=True,
                   remove_TCO_rendering=False, depth_normalization_type="tCR_scale_clamp_center")
REFINER_RGBD_CFG = dict(REFINER_CFG, render_depth=True, input_depth=True)


def n_inputs(cfg):
    return (3 + int(cfg["input_depth"])) + (6 + int(cfg["render_depth"])) * cfg["n_rendered_views"]


def _calibration_batch(c, seed, n=4, h=240, w=320):
    """Smooth images in [0,1]; half of them with the render channels masked to a blob on black, like real inputs."""
    g = torch.Generator().manual_seed(1000 + seed)
    x = torch.rand(n, c, h // 8, w // 8, generator=g)
    x = torch.nn.functional.interpolate(x, size=(h, w), mode="bilinear", align_corners=False)
    yy, xx = torch.meshgrid(torch.linspace(-1, 1, h), torch.linspace(-1, 1, w), indexing="ij")
    blob = ((xx ** 2 + yy ** 2) < 0.4).float()
    x[n // 2:, 3:] *= blob
    return x.clamp(0, 1)


_SD_CACHE = {}


def make_state_dict(cfg, seed=0):
    """Seeded random weights in the checkpoint layout with a conditioned head: the head is made orthogonal to the
    dominant feature direction of a calibration batch and scaled so that coarse logits are O(1) and pose updates are
    small (R ~ I, v_z ~ 1) -- random heads otherwise produce |logit| ~ 300 and 20x depth jumps."""
    key = (tuple(sorted(cfg.items())), seed)
    if key in _SD_CACHE:
        return dict(_SD_CACHE[key])
    head = "pose_fc" if cfg["predict_pose_update"] else "views_logits_head"
    dim = 9 if cfg["predict_pose_update"] else cfg["n_rendered_views"]
    c = n_inputs(cfg)
    sd = resnet_ref.init_state_dict(c, head, dim, seed=seed)
    with torch.no_grad():
        pooled = resnet_ref.pooled_features(sd, _calibration_batch(c, seed))
        feats = torch.nn.functional.linear(pooled, sd["backbone.fc.weight"], sd["backbone.fc.bias"])
    v = torch.linalg.svd(feats, full_matrices=False)[2][0]
    W = sd[head + ".weight"]
    W = W - (W @ v).unsqueeze(1) * v.unsqueeze(0)
    raw = feats @ W.t()
    W = W * ((0.02 if cfg["predict_pose_update"] else 1.5) / (raw - raw.mean(0)).std().clamp_min(1e-12))
    sd[head + ".weight"] = W
    offset = (feats @ W.t()).mean(0)
    if cfg["predict_pose_update"]:
        sd[head + ".bias"] = torch.tensor([1.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0, 1.0]) - offset
    else:
        sd[head + ".bias"] = -offset
    _SD_CACHE[key] = dict(sd)
    return sd


def detection_for_pose(K, TCO, points, pad=4.0):
    """Axis-aligned bbox (x1,y1,x2,y2) of the projected points."""
    P = (TCO[:3, :3] @ points.T + TCO[:3, 3:4])
    uv = (K @ P)
    uv = uv[:2] / uv[2:]
    return torch.tensor([uv[0].min() - pad, uv[1].min() - pad, uv[0].max() + pad, uv[1].max() + pad])


# ------------------------------------------------------------------------- the shared two-object pipeline scenario
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


def check_pipeline_against_golden(golden, coarse_poses, coarse_logits, kept_hypotheses, exact_network: bool,
                                  final_labels=None, final_hypotheses=None, final_poses=None):
    """`golden`: tests/golden/pipeline.npz (outputs of the reference's own PoseEstimator.run_inference_pipeline).
    coarse_poses [2*72,4,4] / coarse_logits [2*72] in the reference's row order (detection-major, grid order);
    kept_hypotheses: per detection, the set of hypothesis ids that survived the coarse filter.

    Tolerances: the initial poses are fp32 geometry -> rtol 2e-5 / atol 2e-6 (the oracle reproduces the reference to
    1e-5 / 1e-6, the CUDA path the oracle to the same).  Logits: an fp32 network (`exact_network`) must reproduce them to
    rtol 1e-4 / atol 1e-5 and give the same final result; the bf16 CUDA network is checked through what the logits are
    used for -- the same survivors wherever the reference's ranking margin is well above the observed noise."""
    want_poses = torch.from_numpy(golden["coarse_poses"])
    assert torch.allclose(coarse_poses.float().cpu(), want_poses, rtol=2e-5, atol=2e-6)
    want = torch.from_numpy(golden["coarse_logit"]).float()
    got = torch.as_tensor(np.array(torch.as_tensor(coarse_logits).cpu() if torch.is_tensor(coarse_logits) else coarse_logits, dtype=np.float32))
    err = (got - want).abs()
    if exact_network:
        assert torch.allclose(got, want, rtol=1e-4, atol=1e-5), f"coarse logits: max err {err.max():.3g}"
        tol = 1e-4
    else:
        tol = 4.0 * err.median().item() + 0.05
    n_det, m = golden["kept_hypotheses"].shape[0], want.numel() // golden["kept_hypotheses"].shape[0]
    k = golden["kept_hypotheses"].shape[1]
    checked = 0
    for det in range(n_det):
        ranked = torch.sort(want[det * m:(det + 1) * m], descending=True).values
        if (ranked[k - 1] - ranked[k]).item() > 2 * tol:
            assert set(int(h) for h in kept_hypotheses[det]) == set(int(h) for h in golden["kept_hypotheses"][det]), det
            checked += 1
    if exact_network:
        assert checked == n_det, "the scenario was chosen with clear margins"
        order_w, order_g = np.argsort(golden["final_label"]), np.argsort(np.asarray(final_labels))
        assert [int(h) for h in np.asarray(final_hypotheses)[order_g]] == [int(h) for h in golden["final_hypothesis"][order_w]]
        assert torch.allclose(final_poses.float().cpu()[order_g], torch.from_numpy(golden["final_poses"])[order_w],
                              rtol=1e-4, atol=1e-5)
    return dict(max_logit_err=err.max().item(), survivors_checked=checked)
