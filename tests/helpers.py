"""Shared builders for the tests: the synthetic workloads (re-exported from workloads/), oracle objects, reference
adapters and the checker of the reference's pipeline fixtures."""
from __future__ import annotations

import numpy as np
import torch

from megapose6d_b200 import procedural
from oracle import pipeline_ref, resnet_ref  # noqa: F401
from workloads.scenes import (FULLSIZE, bench_scene, detection_for_pose, make_scene, pipeline_scenario,  # noqa: F401
                              rgbd_scene, ycbv_scene)
from workloads.weights import (COARSE_CFG, REFINER_CFG, REFINER_RGBD_CFG, calibration_batch, make_state_dict,  # noqa: F401
                               n_inputs)

_calibration_batch = calibration_batch


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
