"""Texture support: file readers (OBJ + MTL + image, PLY with per-vertex or per-corner texture coordinates)
and the texture part of the renderer contract in the CPU oracle."""
import numpy as np
import torch
from PIL import Image

from megapose6d_b200 import meshes, procedural
from megapose6d_b200.meshes import TriMesh
from megapose6d_b200.object_dataset import RigidObject, RigidObjectDataset
from oracle import pipeline_ref
from tests import helpers


def _bilinear(tex, u, v):
    """Independent (float64) restatement of the sampling rule: repeat wrap, v up, texel centres, byte / 255."""
    th, tw = tex.shape[:2]
    u, v = u - np.floor(u), v - np.floor(v)
    x, y = u * tw - 0.5, (1.0 - v) * th - 0.5
    x0, y0 = int(np.floor(x)), int(np.floor(y))
    fx, fy = x - x0, y - y0
    t = tex.astype(np.float64) / 255.0
    c = lambda r, i: t[r % th, i % tw]  # noqa: E731
    top = c(y0, x0) + fx * (c(y0, x0 + 1) - c(y0, x0))
    bot = c(y0 + 1, x0) + fx * (c(y0 + 1, x0 + 1) - c(y0 + 1, x0))
    return top + fy * (bot - top)


def _quad(tex, uv, colors=None):
    """Unit-ish quad in the z = 0 plane facing the camera (-z), uv per corner."""
    v = np.array([[-0.1, -0.075, 0], [0.1, -0.075, 0], [0.1, 0.075, 0], [-0.1, 0.075, 0]], np.float64)
    f = np.array([[0, 1, 2], [0, 2, 3]], np.int32)
    n = np.tile([[0, 0, -1.0]], (4, 1))
    return TriMesh(v, f, n, colors, np.asarray(uv, np.float64), tex, texture_modulate=colors is not None)


def test_oracle_texture_sampling_matches_the_rule():
    tex = procedural.checker_texture(12, 16, cell=3, seed=4)
    uv = [[-0.25, 1.3], [1.4, 1.3], [1.4, -0.2], [-0.25, -0.2]]  # beyond [0,1]: exercises the repeat wrap
    ds = RigidObjectDataset([RigidObject("q", mesh=_quad(tex, uv))])
    rm = helpers.ref_meshes_from_dataset(ds)
    T = torch.eye(4).unsqueeze(0)
    T[0, 2, 3] = 0.5
    K = torch.tensor([[[600.0, 0, 160], [0, 600, 120], [0, 0, 1]]])
    out = pipeline_ref.RefRenderer(rm, quantize8=False).render(["q"], T, K, None, (240, 320), render_depth=True)
    rgb, depth = out["rgbs"][0].numpy(), out["depths"][0, 0].numpy()
    ii, jj = np.nonzero(depth > 0)
    assert len(ii) > 20000
    rs = np.random.RandomState(0)
    for k in rs.choice(len(ii), 400, replace=False):
        i, j = ii[k], jj[k]
        # the quad is fronto-parallel: object coordinates are affine in the pixel
        X, Y = (j + 0.5 - 160) / 600 * 0.5, (i + 0.5 - 120) / 600 * 0.5
        a, b = (X + 0.1) / 0.2, (Y + 0.075) / 0.15
        u = uv[0][0] + a * (uv[1][0] - uv[0][0])
        v = uv[0][1] + b * (uv[3][1] - uv[0][1])
        assert np.allclose(rgb[:, i, j], _bilinear(tex, u, v), atol=2e-3), (i, j, rgb[:, i, j], _bilinear(tex, u, v))


def test_oracle_texture_modulates_vertex_colours_and_quantises():
    tex = procedural.checker_texture(8, 8, cell=2, seed=1)
    uv = [[0, 1], [1, 1], [1, 0], [0, 0]]
    cols = np.array([[1.0, 0.5, 0.25]] * 4)
    T = torch.eye(4).unsqueeze(0)
    T[0, 2, 3] = 0.5
    K = torch.tensor([[[600.0, 0, 160], [0, 600, 120], [0, 0, 1]]])
    plain = pipeline_ref.RefRenderer(helpers.ref_meshes_from_dataset(RigidObjectDataset([RigidObject("q", mesh=_quad(tex, uv))])),
                                     quantize8=False).render(["q"], T, K, None, (240, 320))["rgbs"][0]
    modded = pipeline_ref.RefRenderer(helpers.ref_meshes_from_dataset(RigidObjectDataset([RigidObject("q", mesh=_quad(tex, uv, cols))])),
                                      quantize8=True).render(["q"], T, K, None, (240, 320))["rgbs"][0]
    want = (plain * torch.tensor([1.0, 0.5, 0.25]).view(3, 1, 1) * 255).round() / 255
    assert torch.allclose(modded, want, atol=1.01 / 255)
    levels = (modded * 255).round() / 255
    assert torch.equal(levels, modded)  # quantised to k / 255


def test_obj_mtl_texture_reader(tmp_path):
    tex = procedural.checker_texture(10, 14, cell=2, seed=3)
    Image.fromarray(tex).save(tmp_path / "diffuse.png")
    (tmp_path / "model.mtl").write_text("newmtl skin\nKd 0.8 0.8 0.8\nmap_Kd -s 1 1 1 diffuse.png\n")
    (tmp_path / "model.obj").write_text(
        "mtllib model.mtl\nusemtl skin\n"
        "v 0 0 0\nv 1 0 0\nv 1 1 0\nv 0 1 0\n"
        "vt 0 0\nvt 1 0\nvt 1 1\nvt 0 1\nvt 0.5 0.5\n"
        "vn 0 0 1\n"
        "f 1/1/1 2/2/1 3/3/1\n"        # corner 3 with vt 3 ...
        "f 1/1/1 3/5/1 4/4/1\n")       # ... and with vt 5: position 3 must be split
    m = meshes.load_mesh(tmp_path / "model.obj")
    assert m.texture is not None and np.array_equal(m.texture, tex)
    assert m.vertices.shape == (5, 3) and m.uv.shape == (5, 2) and m.faces.shape == (2, 3)
    assert np.allclose(m.vertices[m.faces[0]], [[0, 0, 0], [1, 0, 0], [1, 1, 0]])
    assert np.allclose(m.uv[m.faces[1]], [[0, 0], [0.5, 0.5], [0, 1]])
    assert np.allclose(m.vertex_normals, [[0, 0, 1]] * 5)
    assert m.with_defaults().texture_modulate is False
    # a missing image leaves an untextured mesh (vertex colours / default albedo)
    (tmp_path / "model.mtl").write_text("newmtl skin\nmap_Kd nowhere.png\n")
    m2 = meshes.load_mesh(tmp_path / "model.obj").with_defaults()
    assert m2.texture is None and m2.uv is None


def _write_ply(path, header_extra, vertex_props, vertices, face_props, faces):
    lines = ["ply", "format ascii 1.0"] + header_extra + [f"element vertex {len(vertices)}"]
    lines += [f"property float {p}" for p in vertex_props]
    lines += [f"element face {len(faces)}"] + face_props + ["end_header"]
    lines += [" ".join(str(x) for x in v) for v in vertices] + faces
    path.write_text("\n".join(lines) + "\n")


def test_ply_texture_readers(tmp_path):
    tex = procedural.checker_texture(6, 6, cell=1, seed=2)
    Image.fromarray(tex).save(tmp_path / "tex.png")
    # per-vertex texture coordinates (BOP models: texture_u / texture_v + `comment TextureFile`)
    _write_ply(tmp_path / "a.ply", ["comment TextureFile tex.png"], ["x", "y", "z", "texture_u", "texture_v"],
               [(0, 0, 0, 0, 0), (1, 0, 0, 1, 0), (1, 1, 0, 1, 1), (0, 1, 0, 0, 1)],
               ["property list uchar int vertex_indices"], ["3 0 1 2", "3 0 2 3"])
    a = meshes.load_mesh(tmp_path / "a.ply")
    assert np.array_equal(a.texture, tex) and np.allclose(a.uv, [[0, 0], [1, 0], [1, 1], [0, 1]]) and a.faces.shape == (2, 3)
    # per-corner texture coordinates (MeshLab export): shared positions with different uv are split
    _write_ply(tmp_path / "b.ply", ["comment TextureFile tex.png"], ["x", "y", "z"],
               [(0, 0, 0), (1, 0, 0), (1, 1, 0), (0, 1, 0)],
               ["property list uchar int vertex_indices", "property list uchar float texcoord"],
               ["3 0 1 2 6 0 0 1 0 1 1", "3 0 2 3 6 0 0 0.5 0.5 0 1"])
    b = meshes.load_mesh(tmp_path / "b.ply")
    assert b.vertices.shape == (5, 3) and np.allclose(b.uv[b.faces[1]], [[0, 0], [0.5, 0.5], [0, 1]])
    assert np.allclose(b.vertices[b.faces[1]], [[0, 0, 0], [1, 1, 0], [0, 1, 0]])
    # no texture file: coordinates are dropped, the mesh renders untextured
    _write_ply(tmp_path / "c.ply", [], ["x", "y", "z", "texture_u", "texture_v"],
               [(0, 0, 0, 0, 0), (1, 0, 0, 1, 0), (1, 1, 0, 1, 1)], ["property list uchar int vertex_indices"], ["3 0 1 2"])
    c = meshes.load_mesh(tmp_path / "c.ply")
    assert c.texture is None and c.uv is None


def test_textured_box_object_goes_through_the_mesh_database():
    ds = RigidObjectDataset([RigidObject("box", mesh=procedural.textured_box()),
                             RigidObject("ball", mesh=procedural.bumpy_sphere(n_seg=24, n_lat=13))])
    rm = helpers.ref_meshes_from_dataset(ds)
    assert rm.tex_dims.tolist() == [[48, 64], [0, 0]] and rm.tex_offsets.tolist() == [0, 48 * 64 * 3, 48 * 64 * 3]
    T = torch.from_numpy(procedural.random_poses(2, 5, z_range=(0.3, 0.4))).float()
    K = torch.tensor([[600.0, 0, 160], [0, 600, 120], [0, 0, 1]]).repeat(2, 1, 1)
    out = pipeline_ref.RefRenderer(rm).render(["box", "ball"], T, K, None, (240, 320), render_depth=True)
    box = out["rgbs"][0][:, out["depths"][0, 0] > 0]
    assert box.shape[1] > 3000 and box.std() > 0.1  # the checkerboard, not a flat albedo
