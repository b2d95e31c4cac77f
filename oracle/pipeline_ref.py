"""oracle/pipeline_ref.py -- CPU restatement of the render-and-compare pipeline.

TEST INFRASTRUCTURE ONLY (see oracle/lib3d_ref.py header): the checker for the CUDA path and the timed
CPU baseline of bench.py.  Never imported by megapose6d_b200/.

Restates, in plain torch fp32 on the host:
  * PosePredictor.forward / forward_coarse / crop_inputs / compute_crops_multiview /
    render_images_multiview / normalize_images   (src/megapose/models/pose_rigid.py:180-708)
  * PoseEstimator.forward_coarse_model / forward_refiner / forward_scoring_model /
    filter_pose_estimates / run_inference_pipeline  (src/megapose/inference/pose_estimator.py:102-667)
with the C rasteriser of oracle/raster_ref.c standing in for Panda3D and the closed-form
make_TCO_multiview of oracle/lib3d_ref.py.  tests/test_oracle_vs_reference.py runs the reference's own
PosePredictor / PoseEstimator (oracle/refload.py) on the same inputs and requires identical outputs.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
import time
from collections import defaultdict
from pathlib import Path
from typing import Dict, List, Optional, Sequence

import numpy as np
import pandas as pd
import torch

from . import lib3d_ref as L
from . import resnet_ref
from .so3_ref import load_SO3_grid_reference

_HERE = Path(__file__).resolve().parent
_LIB: Optional[ctypes.CDLL] = None


def build_raster_lib(force: bool = False) -> Path:
    out = _HERE / "_build" / "libraster_ref.so"
    src = _HERE / "raster_ref.c"
    if force or not out.exists() or out.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["make", "-C", str(_HERE), "-s"], check=True)
    return out


def raster_lib() -> ctypes.CDLL:
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(str(build_raster_lib()))
        _LIB.raster_ref_render_batch.restype = ctypes.c_int
        _LIB.raster_ref_render_batch_tex.restype = ctypes.c_int
        _LIB.raster_ref_render.restype = ctypes.c_int
    return _LIB


class RefMeshes:
    """Host arrays in the layout of mpx_meshdb_create + the [L, Nv, 3] point database."""

    def __init__(self, labels: Sequence[str], verts: List[np.ndarray], normals: List[np.ndarray],
                 colors: List[np.ndarray], faces: List[np.ndarray], uvs: Optional[List[Optional[np.ndarray]]] = None,
                 textures: Optional[List[Optional[np.ndarray]]] = None, modulate: Optional[Sequence[bool]] = None):
        self.labels = list(labels)
        self.set_textures(verts, uvs, textures, modulate)
        self.label_to_id = {l: i for i, l in enumerate(self.labels)}
        self.verts = np.ascontiguousarray(np.concatenate(verts), np.float32)
        self.normals = np.ascontiguousarray(np.concatenate(normals), np.float32)
        self.colors = np.ascontiguousarray(np.concatenate(colors), np.float32)
        self.faces = np.ascontiguousarray(np.concatenate(faces), np.int32)
        self.vert_offsets = np.asarray(np.cumsum([0] + [len(v) for v in verts]), np.int64)
        self.face_offsets = np.asarray(np.cumsum([0] + [len(f) for f in faces]), np.int64)
        # points database as MeshDataBase.batched() (lib3d/rigid_mesh_database.py:90-130): float64 scale, pad, float32
        self.points = L.pad_stack_points([torch.tensor(np.asarray(v, np.float64)) for v in verts]).float()

    def set_textures(self, verts, uvs, textures, modulate) -> None:
        """Optional per-mesh texture: uv [nv,2] and an RGB uint8 image [th,tw,3] (row 0 = top); layout of
        mpx_meshdb_set_textures."""
        n = len(verts)
        self.uv = self.tex = self.tex_offsets = self.tex_dims = self.tex_modulate = None
        if not textures or all(t is None for t in textures):
            return
        uv_all, tex_all, offs, dims = [], [], [0], []
        for i in range(n):
            has = textures[i] is not None and uvs is not None and uvs[i] is not None
            uv_all.append(np.asarray(uvs[i], np.float32) if has else np.zeros((len(verts[i]), 2), np.float32))
            if has:
                t = np.ascontiguousarray(textures[i][..., :3], np.uint8)
                tex_all.append(t.reshape(-1))
                dims.append(t.shape[:2])
            else:
                dims.append((0, 0))
            offs.append(offs[-1] + (tex_all[-1].size if has else 0))
        self.uv = np.ascontiguousarray(np.concatenate(uv_all), np.float32)
        self.tex = np.ascontiguousarray(np.concatenate(tex_all), np.uint8)
        self.tex_offsets = np.asarray(offs, np.int64)
        self.tex_dims = np.ascontiguousarray(np.asarray(dims, np.int32))
        self.tex_modulate = np.ascontiguousarray(np.asarray(modulate if modulate is not None else [0] * n, np.int32))

    @staticmethod
    def from_host_arrays(labels, arrays: Dict[str, np.ndarray], points: torch.Tensor) -> "RefMeshes":
        m = RefMeshes.__new__(RefMeshes)
        m.uv = arrays.get("uv")
        m.tex, m.tex_offsets = arrays.get("tex"), arrays.get("tex_offsets")
        m.tex_dims, m.tex_modulate = arrays.get("tex_dims"), arrays.get("tex_modulate")
        m.labels = list(labels)
        m.label_to_id = {l: i for i, l in enumerate(m.labels)}
        m.verts, m.normals, m.colors, m.faces = arrays["verts"], arrays["normals"], arrays["colors"], arrays["faces"]
        m.vert_offsets, m.face_offsets = arrays["vert_offsets"], arrays["face_offsets"]
        m.points = points.detach().cpu().float()
        return m

    def select_points(self, labels: Sequence[str]) -> torch.Tensor:
        return self.points[[self.label_to_id[l] for l in labels]]

    def sample_points(self, labels: Sequence[str], n: int) -> torch.Tensor:
        ids = torch.as_tensor(L.sample_point_ids(self.points.shape[1], n))
        return torch.index_select(self.select_points(labels), 1, ids)


class RefRenderer:
    """`.render()` with the contract of Panda3dBatchRenderer.render (panda3d_batch_renderer.py:217-282)."""

    def __init__(self, meshes: RefMeshes, quantize8: bool = True, normals_gl_axes: bool = False, n_threads: Optional[int] = None,
                 msaa4: bool = False):
        self.meshes = meshes
        self.msaa4 = msaa4
        self.flags = (1 if quantize8 else 0) | (2 if normals_gl_axes else 0)
        self.n_threads = n_threads or os.cpu_count() or 1

    MSAA4_OFFSETS = ((-2 / 16, -6 / 16), (6 / 16, -2 / 16), (-6 / 16, 2 / 16), (2 / 16, 6 / 16))

    def render(self, labels, TCO, K, light_datas=None, resolution=(240, 320), render_depth=False, render_mask=False,
               render_normals=False, point_lights=False, msaa4=False):
        if msaa4 or self.msaa4:
            return self._render_msaa4(labels, TCO, K, resolution, render_depth, render_normals, point_lights)
        return self._render(labels, TCO, K, light_datas, resolution, render_depth, render_mask, render_normals, point_lights)

    def _render_msaa4(self, labels, TCO, K, resolution, render_depth, render_normals, point_lights):
        """Contract of the 4x anti-aliased render (the reference's offscreen buffer has 4x MSAA,
        panda3d_scene_renderer.py:73-74; PARITY UNPINNED like the rest of the renderer): one render per sample position of
        the standard 4x pattern (principal point moved by the sample offset, float32), every sample shaded and quantised on
        its own, pixel = rounded mean of the four 8-bit samples ((k0 + k1 + k2 + k3 + 2) >> 2), or the float mean
        ((s0 + s1) + (s2 + s3)) / 4 without quantisation; depth = the pixel-centre depth."""
        centre = self._render(labels, TCO, K, None, resolution, render_depth, False, render_normals, point_lights)
        parts = []
        for ox, oy in self.MSAA4_OFFSETS:
            Ks = K.detach().float().clone()
            Ks[:, 0, 2] = Ks[:, 0, 2] - np.float32(ox)
            Ks[:, 1, 2] = Ks[:, 1, 2] - np.float32(oy)
            parts.append(self._render(labels, TCO, Ks, None, resolution, False, False, render_normals, point_lights))

        def resolve(key):
            ps = [p[key] for p in parts]
            if self.flags & 1:
                k = sum((p * 255.0).round().to(torch.int32) for p in ps)
                return ((k + 2) >> 2).float() / 255.0
            return ((ps[0] + ps[1]) + (ps[2] + ps[3])) * 0.25

        return dict(rgbs=resolve("rgbs"), normals=resolve("normals") if render_normals else None, depths=centre["depths"])

    def _render(self, labels, TCO, K, light_datas=None, resolution=(240, 320), render_depth=False, render_mask=False,
                render_normals=False, point_lights=False):
        """`point_lights`: shade the albedo under make_scene_lights() (panda3d_scene_renderer.py:104-136) instead of white
        ambient light -- what models with render_normals=False are fed (models/pose_rigid.py:374-378)."""
        if render_mask:
            raise NotImplementedError
        n = len(labels)
        h, w = resolution
        m = self.meshes
        idx = np.ascontiguousarray([m.label_to_id[l] for l in labels], np.int32)
        T = np.ascontiguousarray(TCO.detach().cpu().float().numpy().reshape(n, 16))
        Kn = np.ascontiguousarray(K.detach().cpu().float().numpy().reshape(n, 9))
        rgb = np.zeros((n, 3, h, w), np.float32)
        nrm = np.zeros((n, 3, h, w), np.float32) if render_normals else None
        dep = np.zeros((n, 1, h, w), np.float32) if render_depth else None
        p = lambda a: a.ctypes.data_as(ctypes.c_void_p) if a is not None else None  # noqa: E731
        rc = raster_lib().raster_ref_render_batch_tex(
            ctypes.c_int(len(m.labels)), p(m.verts), p(m.normals), p(m.colors), p(m.vert_offsets), p(m.faces),
            p(m.face_offsets), p(getattr(m, "uv", None)), p(getattr(m, "tex", None)), p(getattr(m, "tex_offsets", None)),
            p(getattr(m, "tex_dims", None)), p(getattr(m, "tex_modulate", None)), p(idx), p(T), p(Kn), ctypes.c_int(n),
            ctypes.c_int(h), ctypes.c_int(w), ctypes.c_uint(self.flags | (4 if point_lights else 0)), p(rgb), p(nrm), p(dep),
            ctypes.c_int(self.n_threads))
        assert rc == 0, f"raster_ref_render_batch_tex failed ({rc})"
        return dict(rgbs=torch.from_numpy(rgb), normals=torch.from_numpy(nrm) if nrm is not None else None,
                    depths=torch.from_numpy(dep) if dep is not None else None)


class RefPosePredictor:
    def __init__(self, sd: Dict[str, torch.Tensor], cfg: dict, meshes: RefMeshes, renderer: RefRenderer,
                 render_size=(240, 320), net=None):
        self.sd = {k: v.detach().cpu() for k, v in sd.items()}
        self.cfg = cfg
        self.meshes = meshes
        self.renderer = renderer
        self.render_size = tuple(render_size)
        self.n_views = cfg["n_rendered_views"]
        self.input_depth = cfg.get("input_depth", False)
        self.render_depth = cfg.get("render_depth", False)
        self.norm_type = cfg.get("depth_normalization_type")
        self.multiview_type = cfg.get("multiview_type", "TCO")
        self.remove_TCO_rendering = cfg.get("remove_TCO_rendering", False)
        self.predict_pose_update = cfg.get("predict_pose_update", True)
        self.render_normals = cfg.get("render_normals", True)
        self.views_inplane_rotations = cfg.get("views_inplane_rotations", False)
        self.n_render_ch = 3 + (3 if self.render_normals else 0) + (1 if self.render_depth else 0)
        self.net = net if net is not None else (lambda x: resnet_ref.forward_any(self.sd, x))

    # pose_rigid.py:180-247
    def crop_inputs(self, images, K, TCO, tCR, labels):
        points = self.meshes.sample_points(labels, 2000)
        uv = L.project_points_robust(points, K, TCO)
        boxes_rend = L.boxes_from_uv(uv)
        boxes_crop, crops = L.deepim_crops_robust(images, boxes_rend, K, TCO, tCR, points, self.render_size, lamb=1.4)
        K_crop = L.get_K_crop_resize(K.clone(), boxes_crop, self.render_size)
        return crops, K_crop, boxes_rend, boxes_crop

    # pose_rigid.py:249-303
    def compute_crops_multiview(self, images, K, TCV_O, tCR, labels):
        bsz, n_views = TCV_O.shape[:2]
        labels_mv = [labels[n] for n in range(bsz) for _ in range(n_views)]
        T = TCV_O.flatten(0, 1)
        tcr = tCR.flatten(0, 1)
        Kr = K.unsqueeze(1).repeat(1, n_views, 1, 1).flatten(0, 1)
        points = self.meshes.sample_points(labels_mv, 200)
        uv = L.project_points_robust(points, Kr, T)
        boxes_rend = L.boxes_from_uv(uv)
        boxes_crop, _ = L.deepim_crops_robust(images, boxes_rend, Kr, T, tcr, points, self.render_size, lamb=1.4,
                                              return_crops=False)
        return L.get_K_crop_resize(Kr.clone(), boxes_crop, self.render_size).view(bsz, n_views, 3, 3)

    # pose_rigid.py:336-408
    def render_images_multiview(self, labels, TCV_O, KV):
        bsz, n_views = TCV_O.shape[:2]
        labels_mv = [labels[n] for n in range(bsz) for _ in range(n_views)]
        d = self.renderer.render(labels_mv, TCV_O.flatten(0, 1), KV.flatten(0, 1), resolution=self.render_size,
                                 render_normals=self.render_normals, render_depth=self.render_depth,
                                 point_lights=not self.render_normals)
        cat = [d["rgbs"]] + ([d["normals"]] if self.render_normals else []) + ([d["depths"]] if self.render_depth else [])
        r = torch.cat(cat, dim=1)
        return r.view(bsz, n_views, r.shape[1], *r.shape[-2:]).flatten(1, 2)

    # pose_rigid.py:410-464
    def normalize_images(self, images, renders, tCR):
        images, renders = images.clone(), renders.clone()
        if self.input_depth:
            images[:, [3]] = L.normalize_depth(images[:, [3]], tCR, self.norm_type)
        if self.render_depth:
            n_views = renders.shape[1] // self.n_render_ch
            dims = (self.n_render_ch - 1) + self.n_render_ch * torch.arange(0, n_views)
            renders[:, dims] = L.normalize_depth(renders[:, dims], tCR, self.norm_type)
        return images, renders

    # pose_rigid.py:498-604
    def forward(self, images, K, labels, TCO, n_iterations=1):
        if not self.input_depth:
            images = images[:, [0, 1, 2]]
        bsz = images.shape[0]
        outputs = dict()
        TCO_input = TCO
        for n in range(n_iterations):
            TCO_input = L.normalize_T(TCO_input)
            tCR = TCO_input[..., :3, -1].clone()
            TCV_O = L.make_TCO_multiview(TCO_input, tCR, self.multiview_type, self.n_views, self.remove_TCO_rendering,
                                         self.views_inplane_rotations)
            tCV_R = TCV_O[..., :3, -1].clone()
            images_crop, K_crop, boxes_rend, boxes_crop = self.crop_inputs(images, K, TCO_input, tCR, labels)
            KV_crop = self.compute_crops_multiview(images, K, TCV_O, tCV_R, labels)
            if not self.remove_TCO_rendering:
                KV_crop[:, 0] = K_crop
            renders = self.render_images_multiview(labels, TCV_O, KV_crop)
            images_crop, renders = self.normalize_images(images_crop, renders, tCR)
            x = torch.cat((images_crop, renders), dim=1)
            out = self.net(x)
            if self.predict_pose_update:
                TCO_output = L.update_pose(TCO_input, K_crop, out, tCR)
            else:
                TCO_output = TCO_input.clone()
            outputs[f"iteration={n + 1}"] = dict(TCO_input=TCO_input, TCO_output=TCO_output, TCV_O_input=TCV_O, tCR=tCR,
                                                 K=K, K_crop=K_crop, KV_crop=KV_crop, boxes_rend=boxes_rend,
                                                 boxes_crop=boxes_crop, network_output=out, renders=renders,
                                                 images_crop=images_crop, x=x)
            TCO_input = TCO_output
        return outputs

    # pose_rigid.py:634-708
    def forward_coarse(self, images, K, labels, TCO_input):
        if not self.input_depth:
            images = images[:, [0, 1, 2]]
        TCO_input = L.normalize_T(TCO_input)
        tCR = TCO_input[..., :3, -1]
        images_crop, K_crop, boxes_rend, boxes_crop = self.crop_inputs(images, K, TCO_input, tCR, labels)
        renders = self.render_images_multiview(labels, TCO_input.unsqueeze(1), K_crop.unsqueeze(1))
        images_crop, renders = self.normalize_images(images_crop, renders, tCR)
        x = torch.cat((images_crop, renders), dim=1)
        logits = self.net(x)
        return dict(logits=logits, scores=torch.sigmoid(logits), images_crop=images_crop, renders=renders, K_crop=K_crop,
                    boxes_crop=boxes_crop, boxes_rend=boxes_rend, x=x)


class RefPoseEstimator:
    """pose_estimator.py:52-667 on host tensors; infos are plain DataFrames + tensors in dicts."""

    def __init__(self, coarse: RefPosePredictor, refiner: RefPosePredictor, bsz_images: int = 128, bsz_objects: int = 8,
                 SO3_grid_size: int = 576):
        self.coarse, self.refiner = coarse, refiner
        self.bsz_images, self.bsz_objects = bsz_images, bsz_objects
        self.SO3_grid = load_SO3_grid_reference(SO3_grid_size)

    def forward_coarse_model(self, images, K, det_df: pd.DataFrame, bboxes: torch.Tensor, max_hypotheses: Optional[int] = None):
        B, M = len(det_df), self.SO3_grid.shape[0]
        df = det_df.loc[det_df.index.repeat(M)].copy()
        df["hypothesis_id"] = np.tile(np.arange(M), B)
        df["bbox_id"] = np.repeat(np.arange(B), M)
        if max_hypotheses is not None:  # bounded sample for the timed CPU baseline
            df = df.iloc[:max_hypotheses]
        n = len(df)
        logits, TCO_all = [], []
        for s in range(0, n, self.bsz_images):
            d = df.iloc[s:s + self.bsz_images]
            im_ids = torch.as_tensor(d["batch_im_id"].to_numpy(copy=True))
            labels = d["label"].tolist()
            K_ = K[im_ids]
            TCO_init = L.TCO_init_from_boxes_autodepth_with_R(bboxes[torch.as_tensor(d["bbox_id"].to_numpy(copy=True))].float(),
                                                              self.coarse.meshes.select_points(labels), K_,
                                                              self.SO3_grid[torch.as_tensor(d["hypothesis_id"].to_numpy(copy=True))])
            out = self.coarse.forward_coarse(images[im_ids], K_, labels, TCO_init)
            logits.append(out["logits"])
            TCO_all.append(TCO_init)
        df = df.reset_index(drop=True)
        logits = torch.cat(logits)
        df["coarse_logit"] = logits.flatten().numpy()
        df["coarse_score"] = torch.sigmoid(logits).flatten().numpy()
        return df, torch.cat(TCO_all)

    @staticmethod
    def filter_pose_estimates(df: pd.DataFrame, top_K: int, field: str):
        keep = df.sort_values(field, ascending=False, kind="stable").groupby(["batch_im_id", "label", "instance_id"]).head(top_K)
        return keep.index.tolist()

    def forward_refiner(self, images, K, df: pd.DataFrame, TCO: torch.Tensor, n_iterations: int):
        outs = defaultdict(list)
        for s in range(0, len(df), self.bsz_objects):
            d = df.iloc[s:s + self.bsz_objects]
            im_ids = torch.as_tensor(d["batch_im_id"].to_numpy(copy=True))
            o = self.refiner.forward(images[im_ids], K[im_ids], d["label"].tolist(), TCO[s:s + self.bsz_objects], n_iterations)
            for k, v in o.items():
                outs[k].append(v)
        merged = dict()
        for k, lst in outs.items():
            merged[k] = {f: torch.cat([x[f] for x in lst]) for f in ("TCO_output", "TCO_input", "K_crop", "boxes_rend", "boxes_crop")}
        return merged

    def forward_scoring_model(self, images, K, df: pd.DataFrame, TCO: torch.Tensor):
        logits = []
        for s in range(0, len(df), self.bsz_images):
            d = df.iloc[s:s + self.bsz_images]
            im_ids = torch.as_tensor(d["batch_im_id"].to_numpy(copy=True))
            logits.append(self.coarse.forward_coarse(images[im_ids], K[im_ids], d["label"].tolist(), TCO[s:s + self.bsz_images])["logits"])
        return torch.cat(logits)

    def run_inference_pipeline(self, images, K, det_df: pd.DataFrame, bboxes: torch.Tensor, n_refiner_iterations=5,
                               n_pose_hypotheses=1):
        t0 = time.time()
        det_df = det_df.reset_index(drop=True)
        if "instance_id" not in det_df:
            det_df["instance_id"] = det_df.groupby(["batch_im_id", "label"]).cumcount().values
        df_c, TCO_c = self.forward_coarse_model(images, K, det_df, bboxes)
        keep = self.filter_pose_estimates(df_c, n_pose_hypotheses, "coarse_logit")
        df_f, TCO_f = df_c.iloc[keep].reset_index(drop=True), TCO_c[keep]
        ref = self.forward_refiner(images, K, df_f, TCO_f, n_refiner_iterations)
        TCO_r = ref[f"iteration={n_refiner_iterations}"]["TCO_output"]
        logits = self.forward_scoring_model(images, K, df_f, TCO_r)
        df_s = df_f.copy()
        df_s["pose_logit"] = logits.flatten().numpy()
        df_s["pose_score"] = torch.sigmoid(logits).flatten().numpy()
        keep2 = self.filter_pose_estimates(df_s, 1, "pose_logit")
        return dict(final_df=df_s.iloc[keep2].reset_index(drop=True), final_poses=TCO_r[keep2], coarse_df=df_c,
                    coarse_poses=TCO_c, filtered_df=df_f, filtered_poses=TCO_f, refiner=ref, scored_df=df_s,
                    time=time.time() - t0)
