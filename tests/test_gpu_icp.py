"""Depth refinement: full-resolution depth render through the CUDA rasteriser + point-to-plane ICP behind the
reference's `ICPRefiner.refine_poses` interface (inference/icp_refiner.py:197-262).  The registration itself is OpenCV's
in the reference (absent here: parity unpinned, see megapose6d_b200/icp_refiner.py), so the checks are the contract's:
the rendered depth equals the oracle rasteriser's, an exact depth map pulls a perturbed pose back onto the true one, too few
points leave the pose untouched, and `poses_input` holds the incoming poses."""
import numpy as np
import pandas as pd
import pytest
import torch

from megapose6d_b200 import load_model, procedural
from megapose6d_b200.icp_refiner import ICPRefiner, compute_masks
from megapose6d_b200.tensor_collection import PandasTensorCollection
from oracle import pipeline_ref
from tests import helpers

pytestmark = pytest.mark.gpu


def _pose_err(a, b):
    dR = a[:3, :3].double().T @ b[:3, :3].double()
    ang = torch.rad2deg(torch.acos(((dR.trace() - 1) / 2).clamp(-1, 1)))
    return ang.item(), (a[:3, 3] - b[:3, 3]).norm().item() * 1000.0


def test_icp_refiner_pulls_perturbed_poses_back(tmp_path):
    ds, _, K = helpers.make_scene(2, seed=12)
    load_model.write_run(tmp_path, "coarse-rgb-906902141", helpers.make_state_dict(helpers.COARSE_CFG, 5))
    load_model.write_run(tmp_path, "refiner-rgb-653307694", helpers.make_state_dict(helpers.REFINER_CFG, 6))
    est = load_model.load_named_model("megapose-1.0-RGB-multi-hypothesis-icp", ds, models_root=tmp_path)
    assert isinstance(est.depth_refiner, ICPRefiner)
    labels = [ds[0].label, ds[1].label, ds[0].label]
    T_true = torch.from_numpy(procedural.random_poses(3, 23, z_range=(0.35, 0.6), xy_range=0.05)).float()
    Kn = K.repeat(3, 1, 1)
    # "measured" depth: one frame per object, rendered by the ORACLE rasteriser at full resolution
    rm = helpers.ref_meshes_from_dataset(ds)
    ref = pipeline_ref.RefRenderer(rm).render(labels, T_true, Kn, None, (480, 640), render_depth=True)["depths"][:, 0]
    got = est.depth_refiner.renderer.render(labels, T_true.cuda(), Kn.cuda(), None, (480, 640), render_depth=True).depths[:, 0]
    assert torch.equal(got.cpu(), ref)  # the depth render of the refiner is the contract's, bit for bit
    depth = ref.clone().cuda()                                    # [B = 3, H, W] metres
    # predictions: the true poses perturbed by ~1.5 degrees and a few millimetres
    rng = np.random.RandomState(3)
    T_pred = T_true.clone()
    for i in range(3):
        w = torch.from_numpy(rng.randn(3)).float()
        w = w / w.norm() * np.deg2rad(1.5)
        Kx = torch.tensor([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])
        T_pred[i, :3, :3] = torch.matrix_exp(Kx) @ T_true[i, :3, :3]
        T_pred[i, :3, 3] += torch.from_numpy(rng.uniform(-0.004, 0.004, 3)).float()
    infos = pd.DataFrame(dict(label=labels, batch_im_id=[0, 1, 2], instance_id=[0, 0, 0]))
    preds = PandasTensorCollection(infos, poses=T_pred.cuda())
    refined, extra = est.depth_refiner.refine_poses(preds, depth=depth, K=K.repeat(3, 1, 1).cuda())
    assert torch.equal(refined.poses_input.cpu(), T_pred) and extra["n_accepted"] == 3
    for i in range(3):
        r0, t0 = _pose_err(T_pred[i], T_true[i])
        r1, t1 = _pose_err(refined.poses[i].cpu(), T_true[i])
        print(f"object {i}: {r0:.3f} deg / {t0:.2f} mm -> {r1:.3f} deg / {t1:.2f} mm")
        assert t1 < 0.35 * t0 and t1 < 1.0 and r1 < 0.6 * r0 + 0.1
    # too few valid points (depth image empty): poses stay, nothing accepted
    empty, extra = est.depth_refiner.refine_poses(preds, depth=torch.zeros_like(depth), K=K.repeat(3, 1, 1).cuda())
    assert torch.equal(empty.poses.cpu(), T_pred) and extra["n_accepted"] == 0
    # the estimator's hook (inference/pose_estimator.py:485-508)
    m, _ = compute_masks("threshold", got[0], depth[0], 0.1)
    assert m.sum() > 1000
