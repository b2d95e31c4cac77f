"""Depth refinement after the render-and-compare pipeline: full-resolution depth render + point-to-plane ICP.

Drop-in for the reference's `ICPRefiner` (src/megapose/inference/icp_refiner.py:197-262, base class
inference/depth_refiner.py:28-56): same constructor, same `refine_poses(predictions, masks, depth, K) ->
(predictions_refined, extra_data)` contract (`poses_input` holds the incoming pose, `poses` the refined one when the
registration is accepted), same pre-processing:

  * the object is rendered at the resolution of the depth image with the predicted pose (icp_refiner.py:219-229) -- here
    by the CUDA rasteriser through `BatchRenderer.render(render_depth=True)`, one view per prediction;
  * masks = measured and rendered depth both valid and within 0.1 m of each other (refiner_utils.compute_masks,
    "threshold"), depth kept in (0.2, 5) m (icp_refiner.py:141-142);
  * back-projection with `getXYZ` (icp_refiner.py:104-127: x = (u - cx) z / fx with INTEGER-truncated u - cx, v - cy, as the
    int16 uv table of the reference does), normals from smoothed depth gradients (`get_normal`, :38-101);
  * fewer than 1000 points on either side => pose left untouched (:150-151); centroid pre-alignment (:155-160);
  * registration accepted when 0 <= residual <= 0.05 (:162-172).

What is NOT the reference's arithmetic: the registration itself.  The reference calls OpenCV's
`cv2.ppf_match_3d_ICP(100, tolerence=0.05, numLevels=4).registerModelToScene` and `cv2.inpaint` before the normal
estimation; OpenCV is not installed here and neither algorithm is in /root/reference, so parity of this stage is
UNPINNED.  This module implements the published scheme those calls follow -- coarse-to-fine (4 levels, points subsampled by
2^level), nearest-neighbour correspondences, rejection of pairs beyond a robust distance threshold, linearised
point-to-plane minimisation, at most 100 iterations in total, residual = mean point-to-plane distance of the inliers -- on
the GPU with batched torch primitives (nearest neighbours by blocked distance matrices; a few thousand points per object).
Holes of the measured depth are filled by normalised Gaussian smoothing instead of Navier-Stokes inpainting.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from .meshes import BatchedMeshes
from .renderer import BatchRenderer, Panda3dLightData


class DepthRefiner:
    """inference/depth_refiner.py:28-56."""

    def refine_poses(self, predictions, masks: Optional[torch.Tensor] = None, depth: Optional[torch.Tensor] = None,
                     K: Optional[torch.Tensor] = None):
        raise NotImplementedError


def compute_masks(mask_type: str, depth_rendered: torch.Tensor, depth_measured: torch.Tensor,
                  depth_delta_thresh: float = 0.1) -> Tuple[torch.Tensor, torch.Tensor]:
    """inference/refiner_utils.py:29-57 on tensors."""
    mask_measured = (depth_measured > 0) & (depth_rendered > 0)
    if mask_type == "threshold":
        mask_measured = mask_measured & ~((depth_measured - depth_rendered).abs() > depth_delta_thresh)
    elif mask_type != "simple":
        raise ValueError(f"Unknown mask type {mask_type}")
    return mask_measured, mask_measured


def get_xyz(depth: torch.Tensor, K: torch.Tensor) -> torch.Tensor:
    """getXYZ (icp_refiner.py:104-127): [H,W] depth -> [H,W,3]; the pixel offsets are truncated to integers like the
    reference's int16 uv table."""
    h, w = depth.shape
    fx, fy, cx, cy = K[0, 0], K[1, 1], K[0, 2], K[1, 2]
    u = torch.trunc(torch.arange(w, device=depth.device, dtype=torch.float32) - cx)
    v = torch.trunc(torch.arange(h, device=depth.device, dtype=torch.float32) - cy)
    return torch.stack((u[None, :] * depth / fx, v[:, None] * depth / fy, depth), dim=-1)


def _gaussian_blur(img: torch.Tensor, sigma: float) -> torch.Tensor:
    radius = int(4 * sigma + 0.5)  # scipy.ndimage.gaussian_filter truncate=4
    x = torch.arange(-radius, radius + 1, device=img.device, dtype=torch.float32)
    k = torch.exp(-0.5 * (x / sigma) ** 2)
    k = k / k.sum()
    h, w = img.shape
    # separable filter as sums of shifted slices over a reflect-padded copy (scipy's default boundary mode)
    ix = torch.arange(-radius, w + radius, device=img.device).abs()
    ix = torch.where(ix >= w, 2 * w - 1 - ix, ix).clamp(0, w - 1)
    t = img[:, ix]
    t = sum(k[j] * t[:, j:j + w] for j in range(2 * radius + 1))
    iy = torch.arange(-radius, h + radius, device=img.device).abs()
    iy = torch.where(iy >= h, 2 * h - 1 - iy, iy).clamp(0, h - 1)
    t = t[iy]
    return sum(k[j] * t[j:j + h] for j in range(2 * radius + 1))


def get_normal(depth: torch.Tensor, K: torch.Tensor) -> torch.Tensor:
    """get_normal (icp_refiner.py:38-101) with refine=True: hole filling + Gaussian smoothing (sigma 2), then the cross
    product of the back-projected image-axis tangents built from np.gradient(depth, 2)."""
    fx, fy, cx, cy = K[0, 0], K[1, 1], K[0, 2], K[1, 2]
    depth = torch.nan_to_num(depth.float())
    valid = (depth != 0).float()
    # holes: normalised convolution (the reference inpaints with cv2.INPAINT_NS, radius 2)
    num, den = _gaussian_blur(depth * valid, 2.0), _gaussian_blur(valid, 2.0)
    filled = torch.where(valid > 0, depth, num / den.clamp_min(1e-6))
    d = _gaussian_blur(filled, 2.0)
    h, w = d.shape
    u = torch.trunc(torch.arange(w, device=d.device, dtype=torch.float32) - cx)[None, :].expand(h, w)
    v = torch.trunc(torch.arange(h, device=d.device, dtype=torch.float32) - cy)[:, None].expand(h, w)
    g0, g1 = torch.gradient(d, spacing=2.0, edge_order=2)  # d/dy, d/dx like np.gradient(depth, 2, edge_order=2)
    v_y = torch.stack((u / fx * g0, d / fy + v / fy * g0, g0), dim=-1)
    v_x = torch.stack((d / fx + u / fx * g1, v / fy * g1, g1), dim=-1)
    n = torch.cross(v_x, v_y, dim=-1)
    norm = n.norm(dim=-1, keepdim=True)
    n = n / torch.where(norm == 0, torch.ones_like(norm), norm)
    return torch.nan_to_num(n)


def _nearest(src: torch.Tensor, tgt: torch.Tensor, block: int = 4096) -> Tuple[torch.Tensor, torch.Tensor]:
    """Index into tgt of the nearest neighbour of every src point and its distance (blocked distance matrices)."""
    idx, dist = [], []
    for s in range(0, src.shape[0], block):
        d = torch.cdist(src[s:s + block], tgt)
        m = d.min(dim=1)
        idx.append(m.indices)
        dist.append(m.values)
    return torch.cat(idx), torch.cat(dist)


def _small_transform(x: torch.Tensor) -> torch.Tensor:
    """Rigid transform from the 6-vector (rotation vector, translation) of the linearised update (exact exponential)."""
    w, t = x[:3].double(), x[3:].double()
    theta = w.norm()
    Kx = torch.zeros(3, 3, dtype=torch.float64, device=x.device)
    Kx[0, 1], Kx[0, 2], Kx[1, 0], Kx[1, 2], Kx[2, 0], Kx[2, 1] = -w[2], w[1], w[2], -w[0], -w[1], w[0]
    eye = torch.eye(3, dtype=torch.float64, device=x.device)
    if theta < 1e-12:
        R = eye + Kx
    else:
        R = eye + torch.sin(theta) / theta * Kx + (1 - torch.cos(theta)) / theta ** 2 * (Kx @ Kx)
    T = torch.eye(4, dtype=torch.float64, device=x.device)
    T[:3, :3], T[:3, 3] = R, t
    return T


def register_point_to_plane(src: torch.Tensor, tgt: torch.Tensor, n_iterations: int = 100, tolerance: float = 0.05,
                            n_levels: int = 4, rejection_scale: float = 2.5) -> Tuple[int, float, torch.Tensor]:
    """Coarse-to-fine point-to-plane ICP of `src` [Ns,6] (xyz, normal) onto `tgt` [Nt,6]: returns (retval, residual, pose)
    like `cv2.ppf_match_3d_ICP.registerModelToScene` (retval 0 on success; pose [4,4] float64 maps src onto tgt)."""
    dev = src.device
    pose = torch.eye(4, dtype=torch.float64, device=dev)
    p_all = src[:, :3].double()
    q_all, nq_all = tgt[:, :3].double(), tgt[:, 3:].double()
    residual = float("inf")
    iters_per_level = max(1, n_iterations // n_levels)
    for level in range(n_levels - 1, -1, -1):
        step = 2 ** level
        p0 = p_all[::step]
        q, nq = q_all[::step], nq_all[::step]
        prev = float("inf")
        for _ in range(iters_per_level):
            p = p0 @ pose[:3, :3].T + pose[:3, 3]
            idx, dist = _nearest(p.float(), q.float())
            thresh = rejection_scale * max(float(torch.median(dist)) * 1.4826, 1e-4) + float(torch.median(dist))
            keep = dist <= thresh
            if int(keep.sum()) < 6:
                return -1, -1.0, pose
            pk, qk, nk = p[keep], q[idx[keep]], nq[idx[keep]]
            r = ((pk - qk) * nk).sum(-1)                              # signed point-to-plane distances
            A = torch.cat((torch.cross(pk, nk, dim=-1), nk), dim=-1)  # d r / d (omega, t)
            H = A.T @ A + 1e-9 * torch.eye(6, dtype=torch.float64, device=dev)
            x = torch.linalg.solve(H, -(A.T @ r))
            pose = _small_transform(x) @ pose
            residual = float(r.abs().mean())
            if abs(prev - residual) < 1e-7 * max(1.0, residual):
                break
            prev = residual
    retval = 0 if 0 <= residual <= tolerance else -1
    return retval, residual, pose


def icp_refinement(depth_measured: torch.Tensor, depth_rendered: torch.Tensor, object_mask_measured: torch.Tensor,
                   cam_K: torch.Tensor, TCO_pred: torch.Tensor, n_min_points: int = 1000) -> Tuple[torch.Tensor, int]:
    """icp_refiner.py:130-175 on device tensors."""
    xyz_t, n_t = get_xyz(depth_measured, cam_K), get_normal(depth_measured, cam_K)
    depth_valid = (depth_measured > 0.2) & (depth_measured < 5) & object_mask_measured
    points_tgt = torch.cat((xyz_t, n_t), dim=-1)[depth_valid]
    xyz_s, n_s = get_xyz(depth_rendered, cam_K), get_normal(depth_rendered, cam_K)
    points_src = torch.cat((xyz_s, n_s), dim=-1)[depth_valid & (depth_rendered > 0)]
    if len(points_tgt) < n_min_points or len(points_src) < n_min_points:
        return torch.full((4, 4), float("nan"), device=TCO_pred.device), -1
    TCO = TCO_pred.double().clone()
    shift = points_tgt[:, :3].mean(0) - points_src[:, :3].mean(0)  # centroid pre-alignment (:155-160)
    TCO[:3, 3] += shift.double()
    points_src = points_src.clone()
    points_src[:, :3] += shift
    tolerance = 0.05
    retval, residual, pose = register_point_to_plane(points_src, points_tgt, 100, tolerance, 4)
    TCO = pose @ TCO
    if residual > tolerance or residual < 0:
        retval = -1
    return TCO.float(), retval


class ICPRefiner(DepthRefiner):
    def __init__(self, mesh_db: BatchedMeshes, renderer: BatchRenderer) -> None:
        self.mesh_db = mesh_db
        self.renderer = renderer
        self.light_datas = [Panda3dLightData("ambient")]

    @torch.no_grad()
    def refine_poses(self, predictions, masks: Optional[torch.Tensor] = None, depth: Optional[torch.Tensor] = None,
                     K: Optional[torch.Tensor] = None):
        """icp_refiner.py:208-262."""
        assert depth is not None and K is not None
        predictions_refined = predictions.clone()
        if "poses_input" not in predictions_refined.tensors:
            predictions_refined.register_tensor("poses_input", predictions.poses.clone())
        resolution = tuple(depth.shape[-2:])
        df = predictions.infos
        labels = df.label.tolist()
        batch_im_ids = torch.as_tensor(df.batch_im_id.to_numpy().copy(), device=K.device)
        N = len(predictions)
        TCO_ = predictions.poses
        K_ = K[batch_im_ids]
        out = self.renderer.render(labels, TCO=TCO_, K=K_, light_datas=[self.light_datas] * N, resolution=resolution,
                                   render_depth=True)
        all_depth_rendered = out.depths
        n_accepted = 0
        for n in range(N):
            view_id = int(batch_im_ids[n])
            depth_measured = depth[view_id].reshape(resolution).float()
            depth_rendered = all_depth_rendered[n].reshape(resolution)
            if masks is None:
                _, mask = compute_masks("threshold", depth_rendered, depth_measured, 0.1)
            else:
                mask = masks[view_id].reshape(resolution).bool()
            TCO_refined, retval = icp_refinement(depth_measured, depth_rendered, mask, K_[n].float(), predictions.poses[n])
            predictions_refined.poses_input[n] = predictions.poses[n].clone()
            if retval != -1:
                predictions_refined.poses[n] = TCO_refined
                n_accepted += 1
        return predictions_refined, dict(n_accepted=n_accepted)
