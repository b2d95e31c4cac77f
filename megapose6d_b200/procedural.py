"""Procedurally generated objects and scenes (no dataset or checkpoint is available offline).

Used by bench.py, __graft_entry__.smoke() and the tests for the synthetic configurations of
BASELINE.json: bumpy UV spheres with per-vertex normals and colours, >= 2000
vertices as the reference's point sampling requires (lib3d/mesh_ops.py:79).
"""
from __future__ import annotations

from typing import List, Tuple

import numpy as np

from .meshes import TriMesh, compute_vertex_normals
from .object_dataset import RigidObject, RigidObjectDataset


def bumpy_sphere(n_seg: int = 100, n_lat: int = 51, radius: float = 0.05, bump: float = 0.25, seed: int = 0,
                 squash: Tuple[float, float, float] = (1.0, 0.8, 0.6)) -> TriMesh:
    """UV sphere with smooth radial bumps; 2*n_seg*(n_lat-1) triangles, n_seg*(n_lat-1)+2 vertices.

    The default (100, 51) gives exactly 10 000 triangles / 5 002 vertices (BASELINE config 5).
    """
    rng = np.random.RandomState(seed)
    th = np.linspace(0, np.pi, n_lat + 1)[1:-1]  # polar angles of the interior rings
    ph = np.arange(n_seg) * (2 * np.pi / n_seg)
    dirs = [np.array([0.0, 0.0, 1.0])]
    for t in th:
        for p in ph:
            dirs.append(np.array([np.sin(t) * np.cos(p), np.sin(t) * np.sin(p), np.cos(t)]))
    dirs.append(np.array([0.0, 0.0, -1.0]))
    dirs = np.asarray(dirs)
    # low-frequency bumps: a few random plane waves evaluated on the unit directions
    r = np.ones(len(dirs))
    for _ in range(6):
        k = rng.randn(3) * 2.5
        r += bump / 6 * np.sin(dirs @ k + rng.uniform(0, 2 * np.pi))
    verts = dirs * r[:, None] * radius * np.asarray(squash)[None]
    nring = len(th)
    faces: List[List[int]] = []
    for s in range(n_seg):
        faces.append([0, 1 + s, 1 + (s + 1) % n_seg])
    for ring in range(nring - 1):
        a0 = 1 + ring * n_seg
        b0 = a0 + n_seg
        for s in range(n_seg):
            s1 = (s + 1) % n_seg
            faces.append([a0 + s, b0 + s, b0 + s1])
            faces.append([a0 + s, b0 + s1, a0 + s1])
    last = len(verts) - 1
    a0 = 1 + (nring - 1) * n_seg
    for s in range(n_seg):
        faces.append([a0 + s, last, a0 + (s + 1) % n_seg])
    faces_a = np.asarray(faces, dtype=np.int32)
    # colours: smooth function of direction, quantised to uint8 levels like a vertex-colour PLY
    col = 0.5 + 0.5 * np.stack([np.sin(3 * dirs[:, 0] + seed), np.sin(4 * dirs[:, 1] + 1.0), np.sin(5 * dirs[:, 2] + 2.0)], 1)
    col = np.round(col * 255) / 255
    return TriMesh(verts, faces_a, compute_vertex_normals(verts, faces_a), col)


def checker_texture(th: int = 48, tw: int = 64, cell: int = 8, seed: int = 0) -> np.ndarray:
    """RGB uint8 [th,tw,3]: a coloured checkerboard over a smooth gradient (every texel distinct from its neighbours)."""
    rs = np.random.RandomState(seed)
    yy, xx = np.mgrid[0:th, 0:tw]
    check = ((yy // cell + xx // cell) % 2).astype(np.float64)
    base = np.stack([xx / max(tw - 1, 1), yy / max(th - 1, 1), 1.0 - xx / max(tw - 1, 1)], axis=-1)
    tint = rs.uniform(0.35, 1.0, size=(3,))
    img = (0.25 + 0.75 * check[..., None]) * (0.4 + 0.6 * base) * tint
    return np.clip(np.round(img * 255.0), 0, 255).astype(np.uint8)


def textured_box(size=(0.10, 0.07, 0.05), seed: int = 0, with_vertex_colors: bool = False) -> TriMesh:
    """Axis-aligned box, 4 vertices per face (24 in all) so that every face carries its own patch of the texture."""
    sx, sy, sz = (0.5 * s for s in size)
    faces_def = [  # (normal, 4 corners counter-clockwise seen from outside)
        ((1, 0, 0), [(sx, -sy, -sz), (sx, sy, -sz), (sx, sy, sz), (sx, -sy, sz)]),
        ((-1, 0, 0), [(-sx, sy, -sz), (-sx, -sy, -sz), (-sx, -sy, sz), (-sx, sy, sz)]),
        ((0, 1, 0), [(sx, sy, -sz), (-sx, sy, -sz), (-sx, sy, sz), (sx, sy, sz)]),
        ((0, -1, 0), [(-sx, -sy, -sz), (sx, -sy, -sz), (sx, -sy, sz), (-sx, -sy, sz)]),
        ((0, 0, 1), [(-sx, -sy, sz), (sx, -sy, sz), (sx, sy, sz), (-sx, sy, sz)]),
        ((0, 0, -1), [(-sx, sy, -sz), (sx, sy, -sz), (sx, -sy, -sz), (-sx, -sy, -sz)]),
    ]
    v, n, uv, f = [], [], [], []
    for k, (nrm, corners) in enumerate(faces_def):
        u0, v0 = (k % 3) / 3.0, (k // 3) / 2.0  # 3 x 2 atlas; deliberately reaching outside [0,1] on the last column
        quad_uv = [(u0, v0), (u0 + 0.4, v0), (u0 + 0.4, v0 + 0.55), (u0, v0 + 0.55)]
        base = len(v)
        v += corners
        n += [nrm] * 4
        uv += quad_uv
        f += [[base, base + 1, base + 2], [base, base + 2, base + 3]]
    rs = np.random.RandomState(seed + 100)
    colors = rs.uniform(0.3, 1.0, size=(len(v), 3)).round(3) if with_vertex_colors else None
    return TriMesh(np.asarray(v, np.float64), np.asarray(f, np.int32), np.asarray(n, np.float64), colors,
                   np.asarray(uv, np.float64), checker_texture(seed=seed), texture_modulate=with_vertex_colors)


def textured_sphere(n_seg: int = 100, n_lat: int = 51, radius: float = 0.05, seed: int = 0) -> TriMesh:
    """`bumpy_sphere` with longitude / latitude texture coordinates and the checker texture instead of vertex colours.  No
    seam duplication: the last column interpolates u from (n_seg-1)/n_seg back to 0 across the texture, which is fine for
    tests (both implementations must agree on it)."""
    m = bumpy_sphere(n_seg, n_lat, radius, seed=seed)
    d = m.vertices / (radius * np.array([1.0, 0.8, 0.6]))
    d = d / np.linalg.norm(d, axis=1, keepdims=True)
    u = (np.arctan2(d[:, 1], d[:, 0]) / (2 * np.pi)) % 1.0
    v = 1.0 - np.arccos(np.clip(d[:, 2], -1, 1)) / np.pi
    return TriMesh(m.vertices, m.faces, m.vertex_normals, None, np.stack([u, v], 1), checker_texture(64, 128, 8, seed))


def make_object_dataset(n_objects: int = 1, seed: int = 0, n_seg=100, n_lat=51) -> RigidObjectDataset:
    """`n_seg` / `n_lat`: one int for all objects or one per object (2 * n_seg * (n_lat - 1) triangles each)."""
    segs = list(n_seg) if isinstance(n_seg, (list, tuple)) else [n_seg] * n_objects
    lats = list(n_lat) if isinstance(n_lat, (list, tuple)) else [n_lat] * n_objects
    objs = []
    for i in range(n_objects):
        n_seg, n_lat = segs[i], lats[i]
        rng = np.random.RandomState(seed + 17 * i)
        squash = tuple(0.6 + 0.4 * rng.rand(3))
        mesh = bumpy_sphere(n_seg=n_seg, n_lat=n_lat, radius=0.04 + 0.03 * rng.rand(), seed=seed + i, squash=squash)
        objs.append(RigidObject(label=f"obj_{i:06d}", mesh=mesh, mesh_units="m"))
    return RigidObjectDataset(objs)


def example_camera(h: int = 480, w: int = 640) -> np.ndarray:
    """Intrinsics of the reference's barbecue-sauce example (README.md:226)."""
    K = np.array([[605.9547119140625, 0.0, 319.029052734375], [0.0, 605.006591796875, 249.67617797851562], [0.0, 0.0, 1.0]])
    if (h, w) != (480, 640):
        K = K.copy()
        K[0] *= w / 640.0
        K[1] *= h / 480.0
    return K


def random_poses(n: int, seed: int = 0, z_range=(0.4, 1.2), xy_range=0.08) -> np.ndarray:
    """Random object poses in front of the camera: [n,4,4] float64."""
    rng = np.random.RandomState(seed)
    q = rng.randn(n, 4)
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    x, y, z, w = q.T
    R = np.stack([
        np.stack([1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)], -1),
        np.stack([2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)], -1),
        np.stack([2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)], -1),
    ], -2)
    T = np.tile(np.eye(4), (n, 1, 1))
    T[:, :3, :3] = R
    T[:, 0, 3] = rng.uniform(-xy_range, xy_range, n)
    T[:, 1, 3] = rng.uniform(-xy_range, xy_range, n)
    T[:, 2, 3] = rng.uniform(*z_range, n)
    return T
