"""The oracle restatements must reproduce the REAL reference code (loaded by path from /root/reference).

Runs only in the build container (marker `reference`); the fixtures it pins are re-checked everywhere by
tests/test_oracle_golden.py.
"""
import numpy as np
import pandas as pd
import pytest
import torch

from oracle import lib3d_ref as L
from oracle import pipeline_ref, refload, resnet_ref, so3_ref
from tests import helpers

pytestmark = pytest.mark.reference


@pytest.fixture(scope="module")
def ref():
    return refload.load()


def _rand_pose(n, seed):
    from megapose6d_b200.procedural import random_poses

    return torch.from_numpy(random_poses(n, seed)).float()


def test_lib3d_functions_match(ref):
    g = torch.Generator().manual_seed(0)
    n = 7
    TCO = _rand_pose(n, 1)
    K = torch.from_numpy(helpers.procedural.example_camera()).float().unsqueeze(0).repeat(n, 1, 1)
    pts = torch.randn(n, 300, 3, generator=g) * 0.05
    assert torch.equal(L.project_points_robust(pts, K, TCO), ref.camera_geometry.project_points_robust(pts, K, TCO))
    uv = L.project_points_robust(pts, K, TCO)
    assert torch.equal(L.boxes_from_uv(uv), ref.camera_geometry.boxes_from_uv(uv))
    boxes = L.boxes_from_uv(uv)
    assert torch.equal(L.get_K_crop_resize(K, boxes, (240, 320)),
                       ref.camera_geometry.get_K_crop_resize(K.clone(), boxes, (480, 640), (240, 320)))
    p9 = torch.randn(n, 9, generator=g)
    assert torch.equal(L.compute_rotation_matrix_from_ortho6d(p9[:, :6]),
                       ref.rotations.compute_rotation_matrix_from_ortho6d(p9[:, :6]))
    Tn = TCO + 0.01 * torch.randn(n, 4, 4, generator=g)
    assert torch.equal(L.normalize_T(Tn), ref.transform_ops.normalize_T(Tn))
    assert torch.equal(L.invert_transform_matrices(TCO), ref.transform_ops.invert_transform_matrices(TCO))
    tCR = TCO[:, :3, 3] + 0.01
    dR = L.compute_rotation_matrix_from_ortho6d(p9[:, :6])
    Kc = L.get_K_crop_resize(K, boxes, (240, 320))
    assert torch.equal(L.pose_update_with_reference_point(TCO, Kc, p9[:, 6:], dR, tCR),
                       ref.cosypose_ops.pose_update_with_reference_point(TCO, Kc, p9[:, 6:], dR, tCR))
    bb = torch.tensor([[384.0, 234, 522, 455]]).repeat(n, 1) + torch.arange(n).view(-1, 1)
    R = L.compute_rotation_matrix_from_ortho6d(torch.randn(n, 6, generator=g))
    assert torch.equal(L.TCO_init_from_boxes_autodepth_with_R(bb, pts, K, R),
                       ref.cosypose_ops.TCO_init_from_boxes_autodepth_with_R(bb, pts, K, R))


def test_crops_match(ref):
    ds, images, K = helpers.make_scene(1, seed=3, with_depth=True)
    n = 3
    TCO = _rand_pose(n, 5)
    Kn = K.repeat(n, 1, 1)
    pts = torch.from_numpy(ds[0].mesh.vertices[:2000]).float().unsqueeze(0).repeat(n, 1, 1)
    uv = L.project_points_robust(pts, Kn, TCO)
    obs = L.boxes_from_uv(uv)
    imgs = images.repeat(n, 1, 1, 1)
    b1, c1 = L.deepim_crops_robust(imgs, obs, Kn, TCO, TCO[:, :3, 3], pts, (240, 320))
    b2, c2 = ref.cropping.deepim_crops_robust(images=imgs, obs_boxes=obs, K=Kn, TCO_pred=TCO, tCR_in=TCO[:, :3, 3],
                                              O_vertices=pts, output_size=(240, 320), lamb=1.4)
    assert torch.equal(b1, b2) and torch.equal(c1, c2)


def test_sample_and_pad(ref):
    pts = torch.arange(5002 * 3, dtype=torch.float32).view(1, 5002, 3)
    ids = L.sample_point_ids(5002, 2000)
    assert torch.equal(pts[:, ids], ref.mesh_ops.sample_points(pts, 2000, deterministic=True))


def test_so3_asset_matches_reference_file():
    from megapose6d_b200.so3 import load_SO3_grid

    for n in (72, 576, 4608):
        assert torch.equal(load_SO3_grid(n), so3_ref.load_SO3_grid_reference(n))


def test_resnet_restatement_matches_reference_module(ref):
    for cfg in (helpers.COARSE_CFG, helpers.REFINER_CFG):
        sd = helpers.make_state_dict(cfg, seed=1)
        c = helpers.n_inputs(cfg)
        net = ref.torchvision_resnet.resnet34(num_classes=512, n_input_channels=c)
        net.load_state_dict({k[len("backbone."):]: v for k, v in sd.items() if k.startswith("backbone.")})
        net.eval()
        x = torch.rand(2, c, 64, 96, generator=torch.Generator().manual_seed(2))
        head = resnet_ref.head_name(sd)
        with torch.no_grad():
            want = torch.nn.functional.linear(net(x), sd[head + ".weight"], sd[head + ".bias"])
            got = resnet_ref.forward(sd, x)
        assert torch.allclose(got, want, rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("backbone_str,cls", [("resnet34", "WideResNet34"), ("resnet18", "WideResNet18")])
def test_wide_resnet_restatement_matches_reference_module(ref, backbone_str, cls):
    """models/wide_resnet.py:59-126 (pre-activation blocks, bare 1x1 downsample, no fc) + the spatial mean of
    PosePredictor.net_forward (models/pose_rigid.py:323-328) + head."""
    sd = resnet_ref.init_state_dict_wide(9, "views_logits_head", 2, seed=5, backbone_str=backbone_str)
    net = getattr(ref.wide_resnet, cls)(n_inputs=9)
    net.load_state_dict({k[len("backbone."):]: v for k, v in sd.items() if k.startswith("backbone.")})
    net.eval()
    x = torch.rand(2, 9, 64, 96, generator=torch.Generator().manual_seed(2))
    with torch.no_grad():
        want = torch.nn.functional.linear(net(x).flatten(2).mean(dim=-1), sd["views_logits_head.weight"], sd["views_logits_head.bias"])
        got = resnet_ref.forward_wide(sd, x)
        emu = resnet_ref.forward_wide_act16_emulated(sd, x)
        bound = resnet_ref.act16_forward_error_bound(sd, x)
    assert torch.allclose(got, want, rtol=1e-5, atol=1e-5)
    assert ((emu - want).abs() <= bound).all()
    from workloads import weights
    sd_w = weights.init_state_dict_wide(9, "views_logits_head", 2, seed=5, backbone_str=backbone_str)
    assert all(torch.equal(sd[k], sd_w[k]) for k in sd)  # the workload generator and the oracle agree on the layout


class _MeshDbAdapter:
    """mesh_db.select(labels).{points, sample_points} as the reference's PosePredictor expects."""

    def __init__(self, meshes, ref):
        self.m, self.ref = meshes, ref

    def select(self, labels):
        pts = self.m.select_points(labels)
        ref = self.ref

        class _Sel:
            points = pts

            @staticmethod
            def sample_points(n, deterministic=False):
                return ref.mesh_ops.sample_points(pts, n, deterministic=deterministic)

        return _Sel()


def _reference_predictor(ref, cfg, sd, meshes, render_size=(240, 320), n_threads=None):
    class Renderer(ref.Panda3dBatchRenderer):
        def __init__(self, inner):
            self.inner = inner

        def render(self, labels, TCO, K, light_datas, resolution, render_depth=False, render_mask=False,
                   render_normals=False):
            d = self.inner.render(labels, TCO, K, None, resolution, render_depth, render_mask, render_normals)
            return type("Out", (), d)

    c = helpers.n_inputs(cfg)
    backbone = ref.torchvision_resnet.resnet34(num_classes=512, n_input_channels=c)
    backbone.n_features = 512
    model = ref.pose_rigid.PosePredictor(
        backbone=backbone, renderer=Renderer(pipeline_ref.RefRenderer(meshes, **({} if n_threads is None else dict(n_threads=n_threads)))),
        mesh_db=_MeshDbAdapter(meshes, ref), render_size=tuple(render_size), n_rendered_views=cfg["n_rendered_views"], multiview_type=cfg["multiview_type"],
        render_normals=True, render_depth=cfg["render_depth"], input_depth=cfg["input_depth"],
        predict_rendered_views_logits=cfg["predict_rendered_views_logits"], remove_TCO_rendering=False,
        predict_pose_update=cfg["predict_pose_update"], depth_normalization_type=cfg["depth_normalization_type"])
    model.load_state_dict(sd)
    return model.eval()


@pytest.mark.parametrize("cfg_name", ["coarse", "refiner", "refiner_rgbd"])
def test_pose_predictor_matches_reference(ref, cfg_name):
    cfg = dict(coarse=helpers.COARSE_CFG, refiner=helpers.REFINER_CFG, refiner_rgbd=helpers.REFINER_RGBD_CFG)[cfg_name]
    ds, images, K = helpers.make_scene(2, seed=4, with_depth=cfg["input_depth"])
    meshes = helpers.ref_meshes_from_dataset(ds)
    sd = helpers.make_state_dict(cfg, seed=3)
    labels = [ds[0].label, ds[1].label]
    TCO = _rand_pose(2, 8)
    imgs, Kn = images.repeat(2, 1, 1, 1), K.repeat(2, 1, 1)
    oracle = pipeline_ref.RefPosePredictor(sd, cfg, meshes, pipeline_ref.RefRenderer(meshes))
    model = _reference_predictor(ref, cfg, sd, meshes)
    with torch.no_grad():
        if cfg["predict_pose_update"]:
            want = model(images=imgs, K=Kn, labels=labels, TCO=TCO, n_iterations=2)
            got = oracle.forward(imgs, Kn, labels, TCO, n_iterations=2)
            for it in ("iteration=1", "iteration=2"):
                w, g = want[it], got[it]
                assert torch.allclose(g["TCO_output"], w.TCO_output, rtol=1e-5, atol=1e-6)
                assert torch.equal(g["K_crop"], w.K_crop) or torch.allclose(g["K_crop"], w.K_crop, rtol=1e-6, atol=1e-4)
                assert torch.allclose(g["KV_crop"], w.KV_crop, rtol=1e-6, atol=1e-4)
                assert torch.allclose(g["renders"], w.renders, atol=1e-6)
                assert torch.allclose(g["images_crop"], w.images_crop, atol=1e-6)
                assert torch.allclose(g["network_output"], w.network_outputs["pose"], rtol=1e-4, atol=1e-5)
        else:
            want = model.forward_coarse(images=imgs, K=Kn, labels=labels, TCO_input=TCO, return_debug_data=True)
            got = oracle.forward_coarse(imgs, Kn, labels, TCO)
            assert torch.allclose(got["renders"], want["renders"], atol=1e-6)
            assert torch.allclose(got["images_crop"], want["images_crop"], atol=1e-6)
            assert torch.allclose(got["logits"], want["logits"], rtol=1e-4, atol=1e-5)


def test_pose_estimator_pipeline_matches_reference(ref):
    """The reference's own PoseEstimator.run_inference_pipeline vs the oracle pipeline (72-rotation grid)."""
    ds, images, K = helpers.make_scene(2, seed=6)
    meshes = helpers.ref_meshes_from_dataset(ds)
    sd_c, sd_r = helpers.make_state_dict(helpers.COARSE_CFG, 5), helpers.make_state_dict(helpers.REFINER_CFG, 6)
    labels = [o.label for o in ds.list_objects]
    TCO_gt = _rand_pose(2, 11)
    TCO_gt[:, 2, 3] = torch.tensor([0.55, 0.7])
    bboxes = torch.stack([helpers.detection_for_pose(K[0], TCO_gt[i], torch.from_numpy(ds[i].mesh.vertices).float())
                          for i in range(2)])
    det_df = pd.DataFrame(dict(label=labels, batch_im_id=0, instance_id=np.arange(2)))
    coarse = _reference_predictor(ref, helpers.COARSE_CFG, sd_c, meshes)
    refiner = _reference_predictor(ref, helpers.REFINER_CFG, sd_r, meshes)
    coarse.cfg = refiner.cfg = None
    with refload.cpu_cuda_patch():
        est = ref.pose_estimator.PoseEstimator(refiner_model=refiner, coarse_model=coarse, bsz_objects=2, bsz_images=64,
                                               SO3_grid_size=72)
        detections = ref.tensor_collection.PandasTensorCollection(det_df.copy(), bboxes=bboxes)
        obs = ref.types.ObservationTensor(images, K)
        final, extra = est.run_inference_pipeline(obs, detections=detections, n_refiner_iterations=2, n_pose_hypotheses=2)
    oc = pipeline_ref.RefPosePredictor(sd_c, helpers.COARSE_CFG, meshes, pipeline_ref.RefRenderer(meshes))
    orf = pipeline_ref.RefPosePredictor(sd_r, helpers.REFINER_CFG, meshes, pipeline_ref.RefRenderer(meshes))
    oest = pipeline_ref.RefPoseEstimator(oc, orf, bsz_images=64, bsz_objects=2, SO3_grid_size=72)
    got = oest.run_inference_pipeline(images, K, det_df.copy(), bboxes, n_refiner_iterations=2, n_pose_hypotheses=2)
    want_coarse = extra["coarse"]["preds"]
    assert np.allclose(got["coarse_df"]["coarse_logit"].values, want_coarse.infos["coarse_logit"].values, rtol=1e-4, atol=1e-5)
    assert torch.allclose(got["coarse_poses"], want_coarse.poses, rtol=1e-5, atol=1e-6)
    assert sorted(got["final_df"]["hypothesis_id"].tolist()) == sorted(final.infos["hypothesis_id"].tolist())
    order_w = np.argsort(final.infos["label"].values)
    order_g = np.argsort(got["final_df"]["label"].values)
    assert torch.allclose(got["final_poses"][order_g], final.poses[order_w], rtol=1e-4, atol=1e-5)
