"""Triangle meshes, point database and the device mesh store.

Host-side mirror of the reference's MeshDataBase / BatchedMeshes / Meshes
(src/megapose/lib3d/rigid_mesh_database.py:57-200, point sampling lib3d/mesh_ops.py:77-87) plus
what the Panda3D side of the reference does when it loads a model for rendering
(src/megapose/panda3d_renderer/panda3d_scene_renderer.py:195-208: scale to metres, apply
`ypr_offset_deg`).  The reference reads mesh files with trimesh / Assimp; here a small PLY/OBJ reader
covers vertex positions, normals, colours, texture coordinates (per vertex or per face corner) and the
diffuse texture named by the OBJ's MTL (`map_Kd`) or the PLY's `comment TextureFile` (the two forms the
BOP / YCB-V / HOPE model sets use).  Corners that share a position but not a texture coordinate or
normal become separate vertices, as they do in trimesh and Assimp.
"""
from __future__ import annotations

import ctypes
import struct
from copy import deepcopy
from dataclasses import dataclass
from pathlib import Path
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import _abi
from .object_dataset import RigidObject, RigidObjectDataset


@dataclass
class TriMesh:
    vertices: np.ndarray  # [nv,3] float64, mesh units
    faces: np.ndarray  # [nf,3] int32
    vertex_normals: Optional[np.ndarray] = None  # [nv,3]
    vertex_colors: Optional[np.ndarray] = None  # [nv,3] in [0,1]
    uv: Optional[np.ndarray] = None  # [nv,2] texture coordinates (v up)
    texture: Optional[np.ndarray] = None  # [th,tw,3] uint8, row 0 = top of the image
    texture_modulate: bool = False  # multiply the texture with the vertex colours (set when the file has both)

    def with_defaults(self) -> "TriMesh":
        normals = self.vertex_normals if self.vertex_normals is not None else compute_vertex_normals(self.vertices, self.faces)
        colors = self.vertex_colors if self.vertex_colors is not None else np.full((len(self.vertices), 3), 0.8)
        textured = self.texture is not None and self.uv is not None
        return TriMesh(np.asarray(self.vertices, np.float64), np.asarray(self.faces, np.int32), np.asarray(normals, np.float64),
                       np.asarray(colors, np.float64), np.asarray(self.uv, np.float64) if textured else None,
                       np.ascontiguousarray(self.texture[..., :3], dtype=np.uint8) if textured else None,
                       bool(textured and self.vertex_colors is not None and self.texture_modulate))


def compute_vertex_normals(vertices: np.ndarray, faces: np.ndarray) -> np.ndarray:
    """Area-weighted average of the incident face normals."""
    v = np.asarray(vertices, np.float64)
    f = np.asarray(faces, np.int64)
    fn = np.cross(v[f[:, 1]] - v[f[:, 0]], v[f[:, 2]] - v[f[:, 0]])
    n = np.zeros_like(v)
    for k in range(3):
        np.add.at(n, f[:, k], fn)
    length = np.linalg.norm(n, axis=1, keepdims=True)
    length[length == 0] = 1.0
    return n / length


# ---------------------------------------------------------------------------------------------
# file readers
# ---------------------------------------------------------------------------------------------
_PLY_TYPES = {"char": "b", "int8": "b", "uchar": "B", "uint8": "B", "short": "h", "int16": "h", "ushort": "H",
              "uint16": "H", "int": "i", "int32": "i", "uint": "I", "uint32": "I", "float": "f", "float32": "f",
              "double": "d", "float64": "d"}


def load_ply(path: Path) -> TriMesh:
    data = Path(path).read_bytes()
    end = data.index(b"end_header")
    end = data.index(b"\n", end) + 1
    header = data[:end].decode("ascii", "replace").splitlines()
    fmt = "ascii"
    elements: List[dict] = []
    for line in header:
        tok = line.split()
        if not tok:
            continue
        if tok[0] == "format":
            fmt = tok[1]
        elif tok[0] == "element":
            elements.append({"name": tok[1], "count": int(tok[2]), "props": []})
        elif tok[0] == "property":
            if tok[1] == "list":
                elements[-1]["props"].append(("list", tok[2], tok[3], tok[4]))
            else:
                elements[-1]["props"].append(("scalar", tok[1], tok[2]))
    body = data[end:]
    verts: Dict[str, np.ndarray] = {}
    faces: List[List[int]] = []
    face_uv: List[Optional[List[float]]] = []  # per-face texcoord list (u0 v0 u1 v1 ...), PLY `texcoord` property
    texture_file = None
    for line in header:
        tok = line.split()
        if len(tok) >= 3 and tok[0] == "comment" and tok[1].lower() == "texturefile":
            texture_file = line.split(None, 2)[2].strip()
    if fmt == "ascii":
        lines = body.decode("ascii", "replace").split("\n")
        li = 0
        for el in elements:
            if el["name"] == "vertex":
                names = [p[2] for p in el["props"]]
                arr = np.array([lines[li + i].split() for i in range(el["count"])], dtype=np.float64)
                verts = {n: arr[:, k] for k, n in enumerate(names)}
            elif el["name"] == "face":
                list_names = [p[3] for p in el["props"] if p[0] == "list"]
                for i in range(el["count"]):
                    tok = lines[li + i].split()
                    pos, row, tc = 0, None, None
                    for name in list_names:
                        n = int(tok[pos])
                        vals = tok[pos + 1:pos + 1 + n]
                        pos += 1 + n
                        if name in ("vertex_indices", "vertex_index"):
                            row = [int(t) for t in vals]
                        elif name == "texcoord":
                            tc = [float(t) for t in vals]
                    if row is not None:
                        faces.append(row)
                        face_uv.append(tc)
            li += el["count"]
    else:
        endian = "<" if fmt == "binary_little_endian" else ">"
        off = 0
        for el in elements:
            if all(p[0] == "scalar" for p in el["props"]):
                dt = np.dtype([(p[2], endian + _PLY_TYPES[p[1]]) for p in el["props"]])
                arr = np.frombuffer(body, dtype=dt, count=el["count"], offset=off)
                off += dt.itemsize * el["count"]
                if el["name"] == "vertex":
                    verts = {n: arr[n].astype(np.float64) for n in arr.dtype.names}
            else:
                for _ in range(el["count"]):
                    row, tc = None, None
                    for p in el["props"]:
                        if p[0] == "list":
                            (n,) = struct.unpack_from(endian + _PLY_TYPES[p[1]], body, off)
                            off += struct.calcsize(_PLY_TYPES[p[1]])
                            vals = struct.unpack_from(endian + _PLY_TYPES[p[2]] * n, body, off)
                            off += struct.calcsize(_PLY_TYPES[p[2]]) * n
                            if p[3] in ("vertex_indices", "vertex_index"):
                                row = list(vals)
                            elif p[3] == "texcoord":
                                tc = list(vals)
                        else:
                            off += struct.calcsize(_PLY_TYPES[p[1]])
                    if el["name"] == "face" and row is not None:
                        faces.append(row)
                        face_uv.append(tc)
    v = np.stack([verts["x"], verts["y"], verts["z"]], axis=1)
    normals = np.stack([verts["nx"], verts["ny"], verts["nz"]], axis=1) if "nx" in verts else None
    colors = None
    if "red" in verts:
        colors = np.stack([verts["red"], verts["green"], verts["blue"]], axis=1) / 255.0
    uv = None
    for un, vn in (("texture_u", "texture_v"), ("s", "t"), ("u", "v")):
        if un in verts and vn in verts:
            uv = np.stack([verts[un], verts[vn]], axis=1)
            break
    texture = _load_texture(Path(path).parent / texture_file) if texture_file else None
    if uv is None and texture is not None and faces and all(t is not None and len(t) == 2 * len(f) for t, f in zip(face_uv, faces)):
        # per-corner texture coordinates: one vertex per distinct (position index, u, v)
        corners = [(f[k], float(t[2 * k]), float(t[2 * k + 1])) for f, t in zip(faces, face_uv) for k in range(len(f))]
        new_faces, src, uv = _split_corners(corners, [len(f) for f in faces])
        v = v[src]
        normals = normals[src] if normals is not None else None
        colors = colors[src] if colors is not None else None
        faces = new_faces
        uv = np.asarray(uv, np.float64)
    if texture is None:
        uv = None
    return TriMesh(v, _triangulate(faces), normals, colors, uv, texture, texture_modulate=colors is not None)


def _load_texture(path: Path) -> Optional[np.ndarray]:
    """RGB uint8 [th,tw,3], row 0 = top; None when the file is missing (the mesh then renders with its vertex colours)."""
    path = Path(path)
    if not path.is_file():
        return None
    from PIL import Image  # local import: only textured meshes need it

    with Image.open(path) as im:
        return np.ascontiguousarray(np.asarray(im.convert("RGB"), dtype=np.uint8))


def _split_corners(corners: Sequence[tuple], face_sizes: Sequence[int]):
    """Distinct corner keys -> new vertex ids (first-seen order).  Returns (faces as lists of new ids, source position index
    per new vertex, the remaining key fields per new vertex)."""
    ids: Dict[tuple, int] = {}
    src: List[int] = []
    rest: List[tuple] = []
    flat: List[int] = []
    for key in corners:
        k = ids.get(key)
        if k is None:
            k = ids[key] = len(src)
            src.append(key[0])
            rest.append(key[1:])
        flat.append(k)
    faces, pos = [], 0
    for n in face_sizes:
        faces.append(flat[pos:pos + n])
        pos += n
    return faces, np.asarray(src, np.int64), rest


def _parse_mtl(path: Path) -> Dict[str, dict]:
    mats: Dict[str, dict] = {}
    cur = None
    if not Path(path).is_file():
        return mats
    for line in Path(path).read_text(errors="replace").splitlines():
        tok = line.split()
        if not tok or tok[0].startswith("#"):
            continue
        if tok[0] == "newmtl":
            cur = mats.setdefault(" ".join(tok[1:]), {})
        elif cur is not None and tok[0] == "Kd" and len(tok) >= 4:
            cur["Kd"] = [float(t) for t in tok[1:4]]
        elif cur is not None and tok[0] == "map_Kd":
            cur["map_Kd"] = line.split(None, 1)[1].strip().split()[-1]  # options (-s, -o ...) precede the file name
    return mats


def load_obj(path: Path) -> TriMesh:
    """Wavefront OBJ: v (optionally with r g b), vt, vn, f with v, v/vt, v//vn or v/vt/vn corners, mtllib / usemtl.  One
    diffuse texture per mesh (the first material with a `map_Kd`, as the single-object model files of BOP / HOPE have)."""
    path = Path(path)
    vs, cols, vts, vns = [], [], [], []
    corners: List[tuple] = []
    sizes: List[int] = []
    mats: Dict[str, dict] = {}
    used: List[str] = []
    for line in path.read_text(errors="replace").splitlines():
        tok = line.split()
        if not tok:
            continue
        if tok[0] == "v":
            vs.append([float(t) for t in tok[1:4]])
            if len(tok) >= 7:
                cols.append([float(t) for t in tok[4:7]])
        elif tok[0] == "vt":
            vts.append([float(tok[1]), float(tok[2]) if len(tok) > 2 else 0.0])
        elif tok[0] == "vn":
            vns.append([float(t) for t in tok[1:4]])
        elif tok[0] == "mtllib":
            mats.update(_parse_mtl(path.parent / line.split(None, 1)[1].strip()))
        elif tok[0] == "usemtl":
            used.append(" ".join(tok[1:]))
        elif tok[0] == "f":
            n = 0
            for t in tok[1:]:
                parts = (t.split("/") + ["", ""])[:3]
                vi = int(parts[0])
                ti = int(parts[1]) if parts[1] else 0
                ni = int(parts[2]) if parts[2] else 0
                corners.append((vi - 1 if vi > 0 else len(vs) + vi,
                                (ti - 1 if ti > 0 else len(vts) + ti) if ti else -1,
                                (ni - 1 if ni > 0 else len(vns) + ni) if ni else -1))
                n += 1
            sizes.append(n)
    v = np.asarray(vs, np.float64).reshape(-1, 3)
    colors = np.asarray(cols) if len(cols) == len(vs) and vs else None
    texture = None
    for name in used + list(mats):
        m = mats.get(name)
        if m and "map_Kd" in m:
            texture = _load_texture(path.parent / m["map_Kd"])
            if texture is not None:
                break
    has_vt = bool(vts) and all(c[1] >= 0 for c in corners)
    has_vn = bool(vns) and all(c[2] >= 0 for c in corners)
    if not (has_vt and texture is not None) and not has_vn:
        faces, pos = [], 0
        for n in sizes:
            faces.append([c[0] for c in corners[pos:pos + n]])
            pos += n
        return TriMesh(v, _triangulate(faces), None, colors)
    keys = [(c[0], c[1] if has_vt and texture is not None else -1, c[2] if has_vn else -1) for c in corners]
    faces, src, rest = _split_corners(keys, sizes)
    rest = np.asarray(rest, np.int64).reshape(-1, 2)
    uv = np.asarray(vts, np.float64)[rest[:, 0]] if has_vt and texture is not None else None
    normals = np.asarray(vns, np.float64)[rest[:, 1]] if has_vn else None
    return TriMesh(v[src], _triangulate(faces), normals, colors[src] if colors is not None else None, uv, texture,
                   texture_modulate=colors is not None)


def _triangulate(faces: Sequence[Sequence[int]]) -> np.ndarray:
    tris = []
    for f in faces:
        for k in range(1, len(f) - 1):
            tris.append([f[0], f[k], f[k + 1]])
    return np.asarray(tris, dtype=np.int32).reshape(-1, 3)


def load_mesh(path: Path) -> TriMesh:
    path = Path(path)
    if path.is_dir():
        cands = sorted(list(path.glob("*.ply")) + list(path.glob("*.obj")))
        assert cands, f"no mesh file in {path}"
        path = cands[0]
    if path.suffix.lower() == ".ply":
        return load_ply(path)
    if path.suffix.lower() == ".obj":
        return load_obj(path)
    raise ValueError(f"unsupported mesh format: {path}")


def _panda_hpr_matrix(ypr_deg) -> np.ndarray:
    """Rotation of Panda3D NodePath.setHpr(h, p, r): heading about Z(up), pitch about X(right), roll
    about Y(forward); parity unpinned (Panda3D absent), identity for the default (0,0,0)."""
    h, p, r = [np.deg2rad(a) for a in ypr_deg]
    Rz = np.array([[np.cos(h), -np.sin(h), 0], [np.sin(h), np.cos(h), 0], [0, 0, 1]])
    Rx = np.array([[1, 0, 0], [0, np.cos(p), -np.sin(p)], [0, np.sin(p), np.cos(p)]])
    Ry = np.array([[np.cos(r), 0, np.sin(r)], [0, 1, 0], [-np.sin(r), 0, np.cos(r)]])
    return Rz @ Rx @ Ry


# ---------------------------------------------------------------------------------------------
# database
# ---------------------------------------------------------------------------------------------
def _sample_ids(n_total: int, n_points: int) -> np.ndarray:
    # identical to np.random.RandomState(0).choice in mesh_ops.py:77-87 (deterministic=True)
    assert n_points <= n_total, f"meshes need at least {n_points} (padded) vertices, have {n_total}"
    return np.random.RandomState(0).choice(n_total, size=n_points, replace=False)


class MeshDataBase:
    def __init__(self, obj_list: List[RigidObject]):
        self.obj_dict = {obj.label: obj for obj in obj_list}
        self.obj_list = obj_list
        self.infos = {obj.label: dict() for obj in obj_list}
        self.meshes: Dict[str, TriMesh] = {}
        for label, obj in self.obj_dict.items():
            mesh = obj.mesh if getattr(obj, "mesh", None) is not None else load_mesh(obj.mesh_path)
            self.meshes[label] = mesh.with_defaults()
            if obj.diameter_meters is None:
                pts = self.meshes[label].vertices * obj.scale
                obj.diameter_meters = float(np.linalg.norm(pts.max(0) - pts.min(0)))

    @staticmethod
    def from_object_ds(object_ds: RigidObjectDataset) -> "MeshDataBase":
        return MeshDataBase([object_ds[n] for n in range(len(object_ds))])

    def batched(self) -> "BatchedMeshes":
        labels, points = [], []
        infos = deepcopy(self.infos)
        for label, mesh in self.meshes.items():
            pts = torch.tensor(mesh.vertices) * self.obj_dict[label].scale
            infos[label]["n_points"] = pts.shape[0]
            points.append(pts)
            labels.append(label)
        # pad_stack_tensors(fill="select_random", deterministic=True), rigid_mesh_database.py:171-200
        n_max = max(p.shape[0] for p in points)
        rs = np.random.RandomState(0)
        padded = []
        for p in points:
            n_pad = n_max - len(p)
            if n_pad > 0:
                ids = rs.choice(np.arange(len(p)), size=n_pad)
                p = torch.cat((p, p[ids]), dim=0)
            padded.append(p)
        return BatchedMeshes(infos, labels, torch.stack(padded).float(), self)


class BatchedMeshes:
    """points [L, Nv, 3] float32 (metres) + the device triangle store used by the rasteriser."""

    def __init__(self, infos, labels, points: torch.Tensor, mesh_db: MeshDataBase):
        self.infos = infos
        self.labels = np.asarray(labels)
        self.label_to_id = {label: n for n, label in enumerate(labels)}
        self.points = points
        self._mesh_db = mesh_db
        self._handle: Optional[ctypes.c_void_p] = None
        self._point_subsets: Dict[int, torch.Tensor] = {}

    # --- reference API
    def select(self, labels: Sequence[str]) -> "Meshes":
        ids = [self.label_to_id[l] for l in labels]
        return Meshes([self.infos[l] for l in labels], self.labels[ids], self.points[ids])

    def cuda(self) -> "BatchedMeshes":
        self.points = self.points.cuda()
        return self

    def float(self) -> "BatchedMeshes":
        self.points = self.points.float()
        return self

    # --- engine side
    def label_ids(self, labels: Sequence[str], device) -> torch.Tensor:
        return torch.tensor([self.label_to_id[l] for l in labels], dtype=torch.int32, device=device)

    def point_subset(self, n_points: int) -> torch.Tensor:
        """[L, n_points, 3] deterministic subset (Meshes.sample_points(n, deterministic=True))."""
        if n_points not in self._point_subsets:
            ids = torch.as_tensor(_sample_ids(self.points.shape[1], n_points), device=self.points.device)
            self._point_subsets[n_points] = torch.index_select(self.points, 1, ids).contiguous()
        sub = self._point_subsets[n_points]
        if sub.device != self.points.device:
            sub = sub.to(self.points.device)
            self._point_subsets[n_points] = sub
        return sub

    @property
    def handle(self) -> ctypes.c_void_p:
        """Device mesh store (mpx_meshdb), created on first use."""
        if self._handle is None:
            verts, normals, colors, faces = [], [], [], []
            v_off, f_off = [0], [0]
            for label in self.labels:
                mesh = self._mesh_db.meshes[str(label)]
                obj = self._mesh_db.obj_dict[str(label)]
                Rm = _panda_hpr_matrix(obj.ypr_offset_deg)
                verts.append((mesh.vertices * obj.scale) @ Rm.T)
                normals.append(mesh.vertex_normals @ Rm.T)
                colors.append(np.clip(mesh.vertex_colors, 0.0, 1.0))
                faces.append(mesh.faces)
                v_off.append(v_off[-1] + len(mesh.vertices))
                f_off.append(f_off[-1] + len(mesh.faces))
            v = np.ascontiguousarray(np.concatenate(verts), np.float32)
            n = np.ascontiguousarray(np.concatenate(normals), np.float32)
            c = np.ascontiguousarray(np.concatenate(colors), np.float32)
            f = np.ascontiguousarray(np.concatenate(faces), np.int32)
            vo = np.asarray(v_off, np.int64)
            fo = np.asarray(f_off, np.int64)
            out = ctypes.c_void_p()
            _abi.check(_abi.lib().mpx_meshdb_create(
                len(self.labels), v.ctypes.data, n.ctypes.data, c.ctypes.data, vo.ctypes.data, f.ctypes.data,
                fo.ctypes.data, ctypes.byref(out)))
            self._handle = out
            self.host_arrays = dict(verts=v, normals=n, colors=c, faces=f, vert_offsets=vo, face_offsets=fo)
            meshes = [self._mesh_db.meshes[str(label)] for label in self.labels]
            if any(m.texture is not None for m in meshes):
                uv_all, tex_all, t_off, dims, mod = [], [], [0], [], []
                for m in meshes:
                    has = m.texture is not None and m.uv is not None
                    uv_all.append(np.asarray(m.uv, np.float32) if has else np.zeros((len(m.vertices), 2), np.float32))
                    if has:
                        tex_all.append(np.ascontiguousarray(m.texture, np.uint8).reshape(-1))
                    dims.append(m.texture.shape[:2] if has else (0, 0))
                    t_off.append(t_off[-1] + (tex_all[-1].size if has else 0))
                    mod.append(1 if (has and m.texture_modulate) else 0)
                uv = np.ascontiguousarray(np.concatenate(uv_all), np.float32)
                tex = np.ascontiguousarray(np.concatenate(tex_all), np.uint8)
                to = np.asarray(t_off, np.int64)
                td = np.ascontiguousarray(np.asarray(dims, np.int32))
                tm = np.ascontiguousarray(np.asarray(mod, np.int32))
                _abi.check(_abi.lib().mpx_meshdb_set_textures(out, uv.ctypes.data, tex.ctypes.data, to.ctypes.data,
                                                              td.ctypes.data, tm.ctypes.data))
                self.host_arrays.update(uv=uv, tex=tex, tex_offsets=to, tex_dims=td, tex_modulate=tm)
        return self._handle

    def __del__(self):
        try:
            if self._handle is not None:
                _abi.lib().mpx_meshdb_destroy(self._handle)
        except Exception:  # noqa: BLE001
            pass


class Meshes:
    def __init__(self, infos, labels, points: torch.Tensor):
        self.infos = infos
        self.labels = np.asarray(labels)
        self.points = points

    def sample_points(self, n_points: int, deterministic: bool = False) -> torch.Tensor:
        if deterministic:
            ids = _sample_ids(self.points.shape[1], n_points)
        else:
            ids = np.random.choice(self.points.shape[1], size=n_points, replace=False)
        return torch.index_select(self.points, 1, torch.as_tensor(ids, device=self.points.device))
