"""Model-configuration branches beyond the three released models (training/pose_models_cfg.py:36-138 builds any of them
from a run's config.yaml): multiview types `TCO+front_1view` / `sphere_26views` and `views_inplane_rotations`
(lib3d/multiview.py:165-246), every `depth_normalization_type` (models/pose_rigid.py:466-496), and `render_normals=False`
(rgb-only renders under make_scene_lights(), models/pose_rigid.py:374-378).  Each variant goes through the model-zoo
directory layout (config.yaml + checkpoint) and is compared with the oracle's PosePredictor on the same inputs; tolerances
as in tests/test_gpu_pipeline.py."""
import pytest
import torch

from megapose6d_b200 import _abi, lib3d, load_model, procedural
from megapose6d_b200.renderer import BatchRenderer, make_scene_lights
from oracle import lib3d_ref as L
from oracle import pipeline_ref, resnet_ref
from tests import helpers

pytestmark = pytest.mark.gpu
ACT = _abi.act_dtype() if torch.cuda.is_available() else torch.float16


@pytest.fixture(scope="module")
def scene():
    ds, images, K = helpers.make_scene(2, seed=8, with_depth=True)
    return ds, images, K, helpers.ref_meshes_from_dataset(ds)


@pytest.mark.parametrize("kind,n_views,remove,inplane", [("TCO+front_1view", 2, False, False), ("sphere_26views", 27, False, False),
                                                         ("TCO+front_3views", 3, True, False), ("TCO+front_1view", 4, True, True)])
def test_multiview_camera_variants(kind, n_views, remove, inplane):
    n = 7
    TCO = torch.from_numpy(procedural.random_poses(n, 3)).float()
    tCR = TCO[:, :3, 3].contiguous()
    got = lib3d.make_TCO_multiview(TCO.cuda(), tCR.cuda(), kind, n_views, remove, inplane).cpu()
    want = L.make_TCO_multiview(TCO, tCR, kind, n_views, remove, inplane)
    assert got.shape == want.shape == (n, n_views, 4, 4)
    assert torch.allclose(got, want, rtol=1e-5, atol=2e-6)


@pytest.mark.parametrize("big", [False, True], ids=["scatter", "tiled"])
def test_point_light_render_bit_exact_vs_oracle(scene, big):
    ds, images, K, rm = scene
    n = 40 if big else 6
    labels = [ds[i % 2].label for i in range(n)]
    TCO = torch.from_numpy(procedural.random_poses(n, 14, z_range=(0.3, 0.8))).float()
    Kc = torch.tensor([[1100.0, 0, 160], [0, 1100, 120], [0, 0, 1]]).repeat(n, 1, 1)
    r = BatchRenderer(object_dataset=ds)
    out = r.render(labels, TCO.cuda(), Kc.cuda(), [make_scene_lights() for _ in labels], (240, 320), render_depth=True)
    ref = pipeline_ref.RefRenderer(rm).render(labels, TCO, Kc, None, (240, 320), render_depth=True, point_lights=True)
    amb = pipeline_ref.RefRenderer(rm).render(labels, TCO, Kc, None, (240, 320))
    assert out.normals is None
    assert torch.equal(out.rgbs.cpu(), ref["rgbs"]) and torch.equal(out.depths.cpu(), ref["depths"])
    covered = ref["depths"][:, 0] > 0
    ratio = (ref["rgbs"].sum(1)[covered] / amb["rgbs"].sum(1)[covered].clamp_min(1e-3))
    # far lights on the six axes: 0.1 + 0.4 * (|nx| + |ny| + |nz|) of the visible side, i.e. between 0.5 and 0.1 + 0.4 sqrt(3)
    assert 0.3 < ratio.median() < 0.8 and covered.float().mean() > 0.03


def _variant_models(scene, tmp_path, cfg_over, seed):
    """(product PosePredictor, oracle RefPosePredictor, cfg) for the refiner configuration REFINER_RGBD_CFG + cfg_over."""
    ds, images, K, rm = scene
    cfg = dict(helpers.REFINER_RGBD_CFG if cfg_over.pop("_rgbd", False) else helpers.REFINER_CFG, **cfg_over)
    sd = helpers.make_state_dict(cfg, seed)
    zoo = dict(load_model.ZOO_CONFIGS["refiner-rgb-653307694"], depth_augmentation=True,
               **{k: v for k, v in cfg.items()})
    load_model.write_run(tmp_path, "variant", sd, cfg=zoo)
    load_model.write_run(tmp_path, "coarse-rgb-906902141", helpers.make_state_dict(helpers.COARSE_CFG, 5))
    coarse, refiner, _ = load_model.load_pose_models("coarse-rgb-906902141", "variant", ds, models_root=tmp_path)
    oracle = pipeline_ref.RefPosePredictor(sd, cfg, rm, pipeline_ref.RefRenderer(rm))
    return refiner, oracle, cfg, sd


VARIANTS = {
    "front_1view": dict(multiview_type="TCO+front_1view", n_rendered_views=2),
    "sphere_26views": dict(multiview_type="sphere_26views", n_rendered_views=27),
    "inplane_rotations": dict(multiview_type="TCO+front_1view", n_rendered_views=4, remove_TCO_rendering=True,
                              views_inplane_rotations=True),
    "no_normals_point_lights": dict(render_normals=False),
    "rgbd_tCR_scale": dict(_rgbd=True, depth_normalization_type="tCR_scale"),
    "rgbd_tCR_center_clamp": dict(_rgbd=True, depth_normalization_type="tCR_center_clamp"),
    "rgbd_none": dict(_rgbd=True, depth_normalization_type="none"),
    "rgbd_no_normals": dict(_rgbd=True, render_normals=False),
    "wide_resnet34": dict(backbone_str="resnet34"),
    "wide_resnet18_rgbd": dict(_rgbd=True, backbone_str="resnet18"),
}


@pytest.mark.parametrize("name", list(VARIANTS))
def test_refiner_variant_matches_oracle(scene, tmp_path, name):
    ds, images, K, rm = scene
    model, oracle, cfg, sd = _variant_models(scene, tmp_path, dict(VARIANTS[name]), seed=20 + len(name))
    model.keep_images = True
    rgbd = cfg["input_depth"]
    n = 3
    labels = [ds[i % 2].label for i in range(n)]
    TCO = torch.from_numpy(procedural.random_poses(n, 19, z_range=(0.4, 0.8))).float()
    imgs = images if rgbd else images[:, :3].contiguous()
    Kn = K.repeat(n, 1, 1)
    got = model(images=imgs.cuda(), K=Kn.cuda(), labels=labels, TCO=TCO.cuda(), n_iterations=2,
                batch_im_ids=torch.zeros(n, dtype=torch.long))
    imgs_n = imgs.repeat(n, 1, 1, 1)
    for it in (1, 2):
        g = got[f"iteration={it}"]
        r = oracle.forward(imgs_n, Kn, labels, g.TCO_input.cpu(), n_iterations=1)["iteration=1"]
        assert g.renders.shape == r["renders"].shape and g.images_crop.shape == r["images_crop"].shape
        assert torch.allclose(g.TCV_O_input.cpu(), r["TCV_O_input"], rtol=1e-5, atol=2e-6)
        assert torch.allclose(g.KV_crop.cpu(), r["KV_crop"], rtol=2e-5, atol=2e-3)
        frac = (g.renders.cpu() != r["renders"]).float().mean().item()
        mean = (g.renders.cpu() - r["renders"]).abs().mean().item()
        assert frac < 0.05 and mean < 2e-3, f"renders: {frac:.3e} of values differ, mean |diff| {mean:.3e}"
        assert torch.allclose(g.images_crop.cpu()[:, :3], r["images_crop"][:, :3], atol=3e-5)
        if rgbd:
            bad = ((g.images_crop.cpu()[:, 3] - r["images_crop"][:, 3]).abs() > 1e-4).float().mean().item()
            assert bad < 3e-3, f"{bad:.2e} of crop depth values differ"
        out_g, out_r = g.network_outputs["pose"].cpu(), r["network_output"]
        bound = resnet_ref.act16_forward_error_bound(sd, r["x"], dtype=ACT)
        err = (out_g - out_r).abs()
        print(f"{name} it {it}: max|pose9 err|={err.max():.4g} (bound {bound.min():.3g}..{bound.max():.3g})")
        assert (err <= bound + 1e-3).all()
        forced = L.update_pose(r["TCO_input"], r["K_crop"], out_g, r["tCR"])
        assert torch.allclose(g.TCO_output.cpu(), forced, rtol=1e-4, atol=1e-5)


def test_msaa4_render_bit_exact_vs_oracle(scene):
    """4x anti-aliased colour / normal render (the reference's offscreen buffer has 4x MSAA,
    panda3d_scene_renderer.py:73-74): four renders at the multisample positions, quantised per sample, rounded mean."""
    ds, images, K, rm = scene
    n = 5
    labels = [ds[i % 2].label for i in range(n)]
    TCO = torch.from_numpy(procedural.random_poses(n, 15, z_range=(0.3, 0.8))).float()
    Kc = torch.tensor([[1100.0, 0, 160], [0, 1100, 120], [0, 0, 1]]).repeat(n, 1, 1)
    r = BatchRenderer(object_dataset=ds, msaa4=True)
    out = r.render(labels, TCO.cuda(), Kc.cuda(), None, (240, 320), render_depth=True, render_normals=True)
    ref = pipeline_ref.RefRenderer(rm).render(labels, TCO, Kc, None, (240, 320), render_depth=True, render_normals=True, msaa4=True)
    one = pipeline_ref.RefRenderer(rm).render(labels, TCO, Kc, None, (240, 320), render_depth=True, render_normals=True)
    assert torch.equal(out.rgbs.cpu(), ref["rgbs"]) and torch.equal(out.normals.cpu(), ref["normals"])
    assert torch.equal(out.depths.cpu(), ref["depths"]) and torch.equal(ref["depths"], one["depths"])
    # silhouette pixels are blends with the black background: strictly between 0 and the single-sample value somewhere
    edge = (ref["rgbs"].sum(1) > 0) & (one["depths"][:, 0] == 0)
    assert edge.sum() > 50


def test_refiner_with_antialiased_renders_matches_oracle(scene, tmp_path):
    """BatchRenderer(msaa4=True) switches the PosePredictor to the un-fused input path (fp32 crops and anti-aliased renders,
    normalised, concatenated, packed): same outputs as the oracle predictor over the oracle's anti-aliased renderer."""
    ds, images, K, rm = scene
    model, _, cfg, sd = _variant_models(scene, tmp_path, dict(), seed=31)
    model.renderer.msaa4 = True
    try:
        oracle = pipeline_ref.RefPosePredictor(sd, cfg, rm, pipeline_ref.RefRenderer(rm, msaa4=True))
        n = 2
        labels = [ds[i % 2].label for i in range(n)]
        TCO = torch.from_numpy(procedural.random_poses(n, 29, z_range=(0.4, 0.8))).float()
        imgs = images[:, :3].contiguous()
        Kn = K.repeat(n, 1, 1)
        model.keep_images = True
        g = model(images=imgs.cuda(), K=Kn.cuda(), labels=labels, TCO=TCO.cuda(), n_iterations=1,
                  batch_im_ids=torch.zeros(n, dtype=torch.long))["iteration=1"]
        r = oracle.forward(imgs.repeat(n, 1, 1, 1), Kn, labels, TCO, n_iterations=1)["iteration=1"]
        frac = (g.renders.cpu() != r["renders"]).float().mean().item()
        assert frac < 0.05, f"anti-aliased renders: {frac:.3e} of values differ"
        bound = resnet_ref.act16_forward_error_bound(sd, r["x"], dtype=ACT)
        assert ((g.network_outputs["pose"].cpu() - r["network_output"]).abs() <= bound + 1e-3).all()
    finally:
        model.renderer.msaa4 = False
