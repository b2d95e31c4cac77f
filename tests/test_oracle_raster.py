"""Properties of the CPU rasteriser (the contract the CUDA kernel is held to)."""
import numpy as np
import torch

from megapose6d_b200 import procedural
from megapose6d_b200.meshes import TriMesh, compute_vertex_normals
from megapose6d_b200.object_dataset import RigidObject, RigidObjectDataset
from oracle import pipeline_ref
from tests import helpers


def _sphere_ds(radius=0.05):
    m = procedural.bumpy_sphere(n_seg=64, n_lat=33, radius=radius, bump=0.0, squash=(1, 1, 1))
    return RigidObjectDataset([RigidObject("s", mesh=m)])


def test_sphere_depth_silhouette_and_normals():
    ds = _sphere_ds()
    rm = helpers.ref_meshes_from_dataset(ds)
    T = torch.eye(4).unsqueeze(0)
    T[0, 2, 3] = 0.5
    K = torch.tensor([[[800.0, 0, 160], [0, 800, 120], [0, 0, 1]]])
    out = pipeline_ref.RefRenderer(rm, quantize8=False).render(["s"], T, K, None, (240, 320), render_depth=True, render_normals=True)
    depth = out["depths"][0, 0]
    mask = depth > 0
    # silhouette radius of a sphere of radius r at distance d: f * r / sqrt(d^2 - r^2)
    r_px = 800 * 0.05 / np.sqrt(0.25 - 0.0025)
    area = np.pi * r_px ** 2
    assert abs(mask.sum().item() - area) / area < 0.03
    # the pixel next to the principal point sees the near pole: depth ~ d - r
    assert abs(depth[119, 159].item() - 0.45) < 2e-4
    # eye normals: analytic sphere normal at an oblique pixel, encoded with the 32-level wrapped texture
    def enc(v):
        u = v * 32 - 0.5
        k0 = int(np.floor(u))
        f = u - k0
        t0, t1 = ((k0 % 32) * 255 // 32) / 255, (((k0 + 1) % 32) * 255 // 32) / 255
        return t0 + f * (t1 - t0)

    i, j = 105, 185
    ray = np.array([(j + 0.5 - 160) / 800, (i + 0.5 - 120) / 800, 1.0])
    ray /= np.linalg.norm(ray)
    c = np.array([0, 0, 0.5])
    b = ray @ c
    t = b - np.sqrt(b * b - (c @ c - 0.05 ** 2))
    n_cv = (t * ray - c) / 0.05                      # OpenCV camera axes
    n_panda = np.array([n_cv[0], n_cv[2], -n_cv[1]])  # x right, y forward, z up
    got = out["normals"][0, :, i, j].numpy()
    assert abs(depth[i, j].item() - t * ray[2]) < 3e-4
    assert np.allclose(got, [enc(v) for v in n_panda], atol=0.04), (got, n_panda)
    # background is exactly zero everywhere
    assert out["rgbs"][0][:, ~mask].abs().sum() == 0 and out["normals"][0][:, ~mask].abs().sum() == 0


def test_invalid_pose_is_black_and_quantisation_levels():
    ds = _sphere_ds()
    rm = helpers.ref_meshes_from_dataset(ds)
    T = torch.eye(4).unsqueeze(0).repeat(2, 1, 1)
    T[:, 2, 3] = 0.4
    T[1, 0, 0] = float("nan")
    K = torch.tensor([[600.0, 0, 160], [0, 600, 120], [0, 0, 1]]).repeat(2, 1, 1)
    out = pipeline_ref.RefRenderer(rm).render(["s", "s"], T, K, None, (240, 320), render_depth=True, render_normals=True)
    assert out["rgbs"][1].abs().sum() == 0 and out["depths"][1].abs().sum() == 0 and out["normals"][1].abs().sum() == 0
    v = out["rgbs"][0] * 255
    assert torch.allclose(v, v.round(), atol=1e-4)  # k/255 levels (uint8 read-back of the reference)


def test_nearest_surface_wins_and_order_independent():
    # two parallel quads, the nearer one must win regardless of triangle order
    def quad(z, col):
        v = np.array([[-.05, -.05, z], [.05, -.05, z], [.05, .05, z], [-.05, .05, z]])
        return v, np.array([[0, 1, 2], [0, 2, 3]]), np.tile(col, (4, 1))

    v1, f1, c1 = quad(0.0, [1, 0, 0])
    v2, f2, c2 = quad(0.02, [0, 1, 0])
    T = torch.eye(4).unsqueeze(0)
    T[0, 2, 3] = 0.5
    K = torch.tensor([[[600.0, 0, 160], [0, 600, 120], [0, 0, 1]]])
    outs = []
    for order in (0, 1):
        vs = np.concatenate([v1, v2]) if order == 0 else np.concatenate([v2, v1])
        fs = np.concatenate([f1, f2 + 4]) if order == 0 else np.concatenate([f2, f1 + 4])
        cs = np.concatenate([c1, c2]) if order == 0 else np.concatenate([c2, c1])
        mesh = TriMesh(vs, fs.astype(np.int32), compute_vertex_normals(vs, fs), cs.astype(float))
        rm = helpers.ref_meshes_from_dataset(RigidObjectDataset([RigidObject("q", mesh=mesh)]))
        outs.append(pipeline_ref.RefRenderer(rm).render(["q"], T, K, None, (240, 320), render_depth=True, render_normals=True))
    assert torch.equal(outs[0]["rgbs"], outs[1]["rgbs"]) and torch.equal(outs[0]["depths"], outs[1]["depths"])
    centre = outs[0]["rgbs"][0, :, 120, 160]
    assert centre[0] == 1 and centre[1] == 0  # the red quad (z = 0.5) hides the green one (z = 0.52)
    assert abs(outs[0]["depths"][0, 0, 120, 160].item() - 0.5) < 1e-6


def test_oracle_point_lights_and_multiview_variants():
    """The contract's render_normals=False light set (ambient 0.1 + six axis lights of 0.4 at 10 radii): for far lights the
    Lambert sum is |nx| + |ny| + |nz| of the normal, so lit / albedo lies in [0.5, 0.1 + 0.4 sqrt(3)]; and the extra
    multiview types (26-view sphere, in-plane rotations) keep every view a rigid transform looking at the object."""
    import numpy as np
    import torch

    from megapose6d_b200 import procedural
    from oracle import lib3d_ref as L
    from oracle import pipeline_ref
    from tests import helpers

    ds, _, _ = helpers.make_scene(1, seed=2)
    rm = helpers.ref_meshes_from_dataset(ds)
    TCO = torch.from_numpy(procedural.random_poses(2, 4, z_range=(0.3, 0.5), xy_range=0.01)).float()
    K = torch.tensor([[600.0, 0, 80], [0, 600, 60], [0, 0, 1]]).repeat(2, 1, 1)
    labels = [ds[0].label] * 2
    r = pipeline_ref.RefRenderer(rm, quantize8=False)
    amb = r.render(labels, TCO, K, None, (120, 160), render_depth=True)
    lit = r.render(labels, TCO, K, None, (120, 160), render_depth=True, point_lights=True)
    assert torch.equal(amb["depths"], lit["depths"])
    cov = (amb["depths"][:, 0] > 0) & (amb["rgbs"].min(1).values > 0.05)
    ratio = lit["rgbs"].permute(0, 2, 3, 1)[cov] / amb["rgbs"].permute(0, 2, 3, 1)[cov]
    assert cov.sum() > 2000 and ratio.min() > 0.45 and ratio.max() < 0.1 + 0.4 * 3 ** 0.5 + 0.02
    assert (ratio.max(1).values - ratio.min(1).values).max() < 1e-4  # one shade per pixel, applied to all three channels

    T = torch.from_numpy(procedural.random_poses(3, 5)).float()
    tCR = T[:, :3, 3].contiguous()
    V = L.make_TCO_multiview(T, tCR, "sphere_26views", 27, False)
    assert V.shape == (3, 27, 4, 4) and torch.allclose(V[:, 0], T)
    R = V[..., :3, :3].double()
    assert torch.allclose(R.transpose(-1, -2) @ R, torch.eye(3, dtype=torch.float64).expand_as(R), atol=1e-5)
    # every extra camera sees the object origin (tCR = t, reference point = origin) on its optical axis
    t = V[:, 1:, :3, 3]
    assert (t[..., :2].abs().max() < 1e-5) and (t[..., 2] > 0).all()
    W = L.make_TCO_multiview(T, tCR, "TCO+front_1view", 4, True, True)
    assert W.shape == (3, 4, 4, 4)
    Rz = torch.tensor([[0.0, -1.0, 0.0], [1.0, 0.0, 0.0], [0.0, 0.0, 1.0]])
    assert torch.allclose(W[:, 1, :3, :3], Rz @ W[:, 0, :3, :3], atol=1e-6) and torch.equal(W[:, 1, :3, 3], W[:, 0, :3, 3])


def test_oracle_msaa4_is_the_mean_of_four_offset_renders():
    import torch

    from megapose6d_b200 import procedural
    from oracle import pipeline_ref
    from tests import helpers

    ds, _, _ = helpers.make_scene(1, seed=2)
    rm = helpers.ref_meshes_from_dataset(ds)
    T = torch.from_numpy(procedural.random_poses(1, 4, z_range=(0.3, 0.4), xy_range=0.01)).float()
    K = torch.tensor([[600.0, 0, 80], [0, 600, 60], [0, 0, 1]]).unsqueeze(0)
    r = pipeline_ref.RefRenderer(rm)
    one = r.render([ds[0].label], T, K, None, (120, 160), render_normals=True, render_depth=True)
    aa = r.render([ds[0].label], T, K, None, (120, 160), render_normals=True, render_depth=True, msaa4=True)
    assert torch.equal(one["depths"], aa["depths"])
    lv = aa["rgbs"] * 255
    assert (lv - lv.round()).abs().max() < 1e-3  # 8-bit levels
    inside = (one["depths"][0, 0] > 0)
    # far from the silhouette the four samples see the same surface: within two 8-bit levels of the single-sample render
    core = torch.nn.functional.avg_pool2d(inside[None, None].float(), 5, 1, 2)[0, 0] == 1
    assert ((aa["rgbs"] - one["rgbs"])[0][:, core].abs().max() <= 3.5 / 255)
    # on the silhouette the pixel is a blend with the black background
    rim = inside & ~core
    assert (aa["rgbs"][0].sum(0)[rim] < one["rgbs"][0].sum(0)[rim] - 1e-3).float().mean() > 0.2


def test_near_plane_clips_per_sample_instead_of_dropping_triangles():
    """A long quad receding from z = 0.04 to z = 0.5 crosses the lens' near plane (z = 0.1, types.py:63-64).  Its triangles
    are kept and every sample nearer than the plane is rejected: depth is > 0 exactly where the surface lies beyond 0.1 m.
    A triangle with a vertex behind the eye plane has no projection and is dropped."""
    v = np.array([[-.03, .02, 0.04], [.03, .02, 0.04], [.03, -.25, 0.5], [-.03, -.25, 0.5]])
    f = np.array([[0, 1, 2], [0, 2, 3]], dtype=np.int32)
    mesh = TriMesh(v, f, compute_vertex_normals(v, f), np.tile([0.2, 0.6, 1.0], (4, 1)))
    rm = helpers.ref_meshes_from_dataset(RigidObjectDataset([RigidObject("q", mesh=mesh)]))
    T = torch.eye(4).unsqueeze(0)
    K = torch.tensor([[[300.0, 0, 160], [0, 300, 200], [0, 0, 1]]])
    out = pipeline_ref.RefRenderer(rm).render(["q"], T, K, None, (240, 320), render_depth=True, render_normals=True)
    d = out["depths"][0, 0]
    assert (d > 0).sum() > 500, "the part beyond the near plane must be rendered"
    assert d[d > 0].min() >= 0.1 - 1e-6 and d[d > 0].min() < 0.103  # cut at the plane, not at a triangle boundary
    assert d.max() <= 0.5 + 1e-6
    # the same quad pushed so that two vertices lie behind the eye: nothing can be projected, nothing is drawn
    T2 = T.clone()
    T2[0, 2, 3] = -0.05
    out2 = pipeline_ref.RefRenderer(rm).render(["q"], T2, K, None, (240, 320), render_depth=True, render_normals=True)
    assert out2["depths"].abs().sum() == 0
