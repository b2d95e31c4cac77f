"""oracle/lib3d_ref.py -- CPU (torch fp32 / numpy fp64) restatement of the reference's 3-D math.

TEST INFRASTRUCTURE ONLY.  Nothing under megapose6d_b200/ may import this module; it is used by
tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs as the checker / timed baseline.

Each function cites the reference code it restates (paths relative to
/root/reference/src/megapose).  The restatement is validated against the reference itself
(loaded by file path, see oracle/refload.py) in tests/test_oracle_vs_reference.py -- that test runs
only where /root/reference exists -- and against the committed fixtures in tests/golden/.
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np
import torch


# ---------------------------------------------------------------------------------------------
# rotations / transforms
# ---------------------------------------------------------------------------------------------
def unitquat_to_rotmat(quat: torch.Tensor) -> torch.Tensor:
    """xyzw unit quaternion -> rotation matrix.

    Restates roma.unitquat_to_rotmat (third-party `roma`, unpinned in conda/environment_full.yaml,
    absent here) as called by utils/transform_utils.py:48-49.
    """
    x, y, z, w = quat[..., 0], quat[..., 1], quat[..., 2], quat[..., 3]
    tx, ty, tz = 2 * x, 2 * y, 2 * z
    twx, twy, twz = tx * w, ty * w, tz * w
    txx, txy, txz = tx * x, ty * x, tz * x
    tyy, tyz, tzz = ty * y, tz * y, tz * z
    m = torch.empty(quat.shape[:-1] + (3, 3), dtype=quat.dtype)
    m[..., 0, 0] = 1 - (tyy + tzz)
    m[..., 0, 1] = txy - twz
    m[..., 0, 2] = txz + twy
    m[..., 1, 0] = txy + twz
    m[..., 1, 1] = 1 - (txx + tzz)
    m[..., 1, 2] = tyz - twx
    m[..., 2, 0] = txz - twy
    m[..., 2, 1] = tyz + twx
    m[..., 2, 2] = 1 - (txx + tyy)
    return m


def compute_rotation_matrix_from_ortho6d(poses: torch.Tensor) -> torch.Tensor:
    """lib3d/rotations.py:25-40."""
    x_raw, y_raw = poses[..., 0:3], poses[..., 3:6]
    x = x_raw / torch.norm(x_raw, p=2, dim=-1, keepdim=True)
    z = torch.cross(x, y_raw, dim=-1)
    z = z / torch.norm(z, p=2, dim=-1, keepdim=True)
    y = torch.cross(z, x, dim=-1)
    return torch.stack((x, y, z), -1)


def normalize_T(T: torch.Tensor) -> torch.Tensor:
    """lib3d/transform_ops.py:106-119 (compute_transform_from_pose9d o pose-9d packing)."""
    pose9 = torch.cat([T[..., :3, 0], T[..., :3, 1], T[..., :3, -1]], dim=-1)
    R = compute_rotation_matrix_from_ortho6d(pose9[..., :6])
    out = torch.zeros(*pose9.shape[:-1], 4, 4, dtype=T.dtype)
    out[..., :3, :3] = R
    out[..., :3, 3] = pose9[..., 6:]
    out[..., 3, 3] = 1
    return out


def invert_transform_matrices(T: torch.Tensor) -> torch.Tensor:
    """lib3d/transform_ops.py:49-57."""
    R = T[..., :3, :3]
    t = T[..., :3, [-1]]
    R_inv = R.transpose(-2, -1)
    out = T.clone()
    out[..., :3, :3] = R_inv
    out[..., :3, [-1]] = -R_inv @ t
    return out


def transform_pts(T: torch.Tensor, pts: torch.Tensor) -> torch.Tensor:
    """lib3d/transform_ops.py:31-46 (3-D T only)."""
    return (T.unsqueeze(-3)[..., :3, :3] @ pts.unsqueeze(-1) + T.unsqueeze(-3)[..., :3, [-1]]).squeeze(-1)


# ---------------------------------------------------------------------------------------------
# camera geometry
# ---------------------------------------------------------------------------------------------
def project_points_robust(points_3d: torch.Tensor, K: torch.Tensor, TCO: torch.Tensor, z_min: float = 0.1):
    """lib3d/camera_geometry.py:40-53."""
    b, n = points_3d.shape[:2]
    pts = torch.cat((points_3d, torch.ones(b, n, 1)), dim=-1)
    P = K @ TCO[:, :3]
    suv = (P.unsqueeze(1) @ pts.unsqueeze(-1)).squeeze(-1)
    z = suv[..., -1]
    suv[..., -1] = torch.max(torch.ones_like(z) * z_min, z)
    suv = suv / suv[..., [-1]]
    return suv[..., :2]


def boxes_from_uv(uv: torch.Tensor) -> torch.Tensor:
    """lib3d/camera_geometry.py:56-64."""
    x1 = uv[..., [0]].min(dim=1)[0]
    y1 = uv[..., [1]].min(dim=1)[0]
    x2 = uv[..., [0]].max(dim=1)[0]
    y2 = uv[..., [1]].max(dim=1)[0]
    return torch.cat((x1, y1, x2, y2), dim=1)


def get_K_crop_resize(K: torch.Tensor, boxes: torch.Tensor, crop_resize: Tuple[int, int]) -> torch.Tensor:
    """lib3d/camera_geometry.py:67-115 (orig_size is unused by the reference's arithmetic)."""
    K = K.float()
    boxes = boxes.float()
    new_K = K.clone()
    cr = torch.tensor(crop_resize, dtype=torch.float)
    final_width, final_height = max(cr), min(cr)
    crop_width = boxes[:, 2] - boxes[:, 0]
    crop_height = boxes[:, 3] - boxes[:, 1]
    crop_cj = (boxes[:, 0] + boxes[:, 2]) / 2
    crop_ci = (boxes[:, 1] + boxes[:, 3]) / 2
    cx = K[:, 0, 2] + (crop_width - 1) / 2 - crop_cj
    cy = K[:, 1, 2] + (crop_height - 1) / 2 - crop_ci
    center_x = (crop_width - 1) / 2
    center_y = (crop_height - 1) / 2
    orig_cx_diff = cx - center_x
    orig_cy_diff = cy - center_y
    scale_x = final_width / crop_width
    scale_y = final_height / crop_height
    new_K[:, 0, 0] = scale_x * K[:, 0, 0]
    new_K[:, 1, 1] = scale_y * K[:, 1, 1]
    new_K[:, 0, 2] = (final_width - 1) / 2 + scale_x * orig_cx_diff
    new_K[:, 1, 2] = (final_height - 1) / 2 + scale_y * orig_cy_diff
    return new_K


# ---------------------------------------------------------------------------------------------
# cropping
# ---------------------------------------------------------------------------------------------
def deepim_boxes(rend_center_uv, obs_boxes, rend_boxes, lamb=1.4, im_size=(240, 320)):
    """lib3d/cropping.py:30-67."""
    lobs, robs, uobs, dobs = obs_boxes[:, [0, 2, 1, 3]].t()
    lrend, rrend, urend, drend = rend_boxes[:, [0, 2, 1, 3]].t()
    xc = rend_center_uv[..., 0, 0]
    yc = rend_center_uv[..., 0, 1]
    w, h = max(im_size), min(im_size)
    r = w / h
    xdist = torch.stack(((lobs - xc).abs(), (lrend - xc).abs(), (robs - xc).abs(), (rrend - xc).abs()), 1).max(1)[0]
    ydist = torch.stack(((uobs - yc).abs(), (urend - yc).abs(), (dobs - yc).abs(), (drend - yc).abs()), 1).max(1)[0]
    width = torch.max(xdist, ydist * r) * 2 * lamb
    height = torch.max(xdist / r, ydist) * 2 * lamb
    return torch.stack((xc - width / 2, yc - height / 2, xc + width / 2, yc + height / 2), dim=1)


def crop_images(images: torch.Tensor, boxes5: torch.Tensor, output_size, sampling_ratio: int = 4):
    """lib3d/cropping.py:113-144; roi_align itself is torchvision's (importable third party)."""
    import torchvision

    nchannels = images.shape[1]
    assert nchannels in (3, 4)
    crops = torchvision.ops.roi_align(images, boxes5, output_size=output_size, sampling_ratio=sampling_ratio)
    if nchannels == 4:
        depth = images[:, [3]]
        valid = torch.zeros_like(depth)
        valid[depth > 0] = 1
        valid_crops = torchvision.ops.roi_align(valid, boxes5, output_size=output_size, sampling_ratio=4)
        mask = torch.ones_like(valid_crops)
        mask[valid_crops < 0.99] = 0
        crops[:, [3]] *= mask
    return crops


def deepim_crops_robust(images, obs_boxes, K, TCO_pred, tCR_in, O_vertices, output_size, lamb=1.4,
                        return_crops=True):
    """lib3d/cropping.py:84-110."""
    h, w = images.shape[-2:]
    bsz = TCO_pred.shape[0]
    uv = project_points_robust(O_vertices, K, TCO_pred)
    rend_boxes = boxes_from_uv(uv)
    TCR = TCO_pred.clone()
    TCR[:, :3, -1] = tCR_in
    center_uv = project_points_robust(torch.zeros(bsz, 1, 3), K, TCR)
    boxes = deepim_boxes(center_uv, obs_boxes, rend_boxes, im_size=(h, w), lamb=lamb)
    crops = None
    if return_crops:
        boxes5 = torch.cat((torch.arange(bsz).unsqueeze(1).float(), boxes), dim=1)
        crops = crop_images(images, boxes5, output_size=output_size, sampling_ratio=4)
    return boxes, crops


def roi_align_scalar(image: np.ndarray, box, out_h: int, out_w: int, sampling: int = 4) -> np.ndarray:
    """Scalar restatement of torchvision roi_align (aligned=False, spatial_scale=1) for one ROI.

    Independent check of the arithmetic the CUDA crop kernel implements; image [C,H,W].
    Small sizes only (pure Python loops).
    """
    c, h, w = image.shape
    x1, y1, x2, y2 = [np.float32(v) for v in box]
    roi_w = max(x2 - x1, np.float32(1.0))
    roi_h = max(y2 - y1, np.float32(1.0))
    bin_w, bin_h = np.float32(roi_w / out_w), np.float32(roi_h / out_h)
    out = np.zeros((c, out_h, out_w), dtype=np.float32)
    for ph in range(out_h):
        for pw in range(out_w):
            acc = np.zeros(c, dtype=np.float32)
            for iy in range(sampling):
                y = np.float32(y1 + ph * bin_h + np.float32(iy + 0.5) * bin_h / np.float32(sampling))
                for ix in range(sampling):
                    x = np.float32(x1 + pw * bin_w + np.float32(ix + 0.5) * bin_w / np.float32(sampling))
                    if y < -1.0 or y > h or x < -1.0 or x > w:
                        continue
                    yy, xx = max(y, np.float32(0)), max(x, np.float32(0))
                    y_low, x_low = int(yy), int(xx)
                    if y_low >= h - 1:
                        y_high = y_low = h - 1
                        yy = np.float32(y_low)
                    else:
                        y_high = y_low + 1
                    if x_low >= w - 1:
                        x_high = x_low = w - 1
                        xx = np.float32(x_low)
                    else:
                        x_high = x_low + 1
                    ly, lx = np.float32(yy - y_low), np.float32(xx - x_low)
                    hy, hx = np.float32(1) - ly, np.float32(1) - lx
                    acc += (hy * hx * image[:, y_low, x_low] + hy * lx * image[:, y_low, x_high]
                            + ly * hx * image[:, y_high, x_low] + ly * lx * image[:, y_high, x_high])
            out[:, ph, pw] = acc / np.float32(sampling * sampling)
    return out


# ---------------------------------------------------------------------------------------------
# pose init / update
# ---------------------------------------------------------------------------------------------
def TCO_init_from_boxes_autodepth_with_R(boxes_2d, model_points_3d, K, R):
    """lib3d/cosypose_ops.py:169-218."""
    bsz = boxes_2d.shape[0]
    z_guess = 1.0
    fxfy = K[:, [0, 1], [0, 1]]
    cxcy = K[:, [0, 1], [2, 2]]
    TCO = torch.eye(4).unsqueeze(0).repeat(bsz, 1, 1)
    TCO[:, 2, 3] = z_guess
    TCO[:, :3, :3] = R
    centers = (boxes_2d[:, [0, 1]] + boxes_2d[:, [2, 3]]) / 2
    TCO[:, :2, 3] = ((centers - cxcy) * z_guess) / fxfy
    C_pts = transform_pts(TCO, model_points_3d)
    dx = C_pts[:, :, 0].max(dim=1).values - C_pts[:, :, 0].min(dim=1).values
    dy = C_pts[:, :, 1].max(dim=1).values - C_pts[:, :, 1].min(dim=1).values
    bb_dx = (boxes_2d[:, 2] - boxes_2d[:, 0]) + 1
    bb_dy = (boxes_2d[:, 3] - boxes_2d[:, 1]) + 1
    z_from_dx = fxfy[:, 0] * dx / bb_dx
    z_from_dy = fxfy[:, 1] * dy / bb_dy
    z = (z_from_dy.unsqueeze(1) + z_from_dx.unsqueeze(1)) / 2
    TCO[:, :2, 3] = ((centers - cxcy) * z) / fxfy
    TCO[:, 2, 3] = z.flatten()
    return TCO


def pose_update_with_reference_point(TCO, K, vxvyvz, dRCO, tCR):
    """lib3d/cosypose_ops.py:33-58."""
    zsrc = tCR[:, [2]]
    vz = vxvyvz[:, [2]]
    ztgt = vz * zsrc
    vxvy = vxvyvz[:, :2]
    fxfy = K[:, [0, 1], [0, 1]]
    xsrcysrc = tCR[:, :2]
    tCR_out = tCR.clone()
    tCR_out[:, 2] = ztgt.flatten()
    tCR_out[:, :2] = ((vxvy / fxfy) + (xsrcysrc / zsrc.repeat(1, 2))) * ztgt.repeat(1, 2)
    tCO_out = (dRCO @ (TCO[:, :3, 3] - tCR).unsqueeze(-1) + tCR_out.unsqueeze(-1)).squeeze(-1)
    out = TCO.clone()
    out[:, :3, 3] = tCO_out
    out[:, :3, :3] = dRCO @ TCO[:, :3, :3]
    return out


def update_pose(TCO, K_crop, pose_outputs, tCR):
    """models/pose_rigid.py:305-312."""
    dR = compute_rotation_matrix_from_ortho6d(pose_outputs[:, 0:6])
    return pose_update_with_reference_point(TCO, K_crop, pose_outputs[:, 6:9], dR, tCR)


# ---------------------------------------------------------------------------------------------
# multi-view cameras (float64 numpy, closed form of the Panda3D scene graph)
# ---------------------------------------------------------------------------------------------
_TCCGL = np.array([[1, 0, 0, 0], [0, 0, -1, 0], [0, 1, 0, 0], [0, 0, 0, 1]], dtype=np.float64)

def _sphere_26():
    """get_26_views_TCO_pos_sphere (lib3d/multiview.py:149-160)."""
    return np.array([[x, y, z] for y in (0, 1, 2) for x in (0, -1, 1) for z in (0, 1, -1) if not (x == 0 and y == 1 and z == 0)],
                    dtype=np.float64)


VIEW_OFFSETS = {
    # lib3d/multiview.py:95-160
    "sphere_26views": _sphere_26(),
    "TCO+front_1view": np.array([[0, 0, 0]], dtype=np.float64),
    "TCO+front_3views": np.array([[0, 0, 0], [1, 0, 0], [-1, 0, 0]], dtype=np.float64),
    "TCO+front_5views": np.array([[0, 0, 0], [1, 0, 0], [-1, 0, 0], [0, 0, 1], [0, 0, -1]], dtype=np.float64),
}


def _look_at(fwd: np.ndarray, up: np.ndarray) -> np.ndarray:
    """Panda3D NodePath.lookAt in its default Z-up right-handed frame (+Y forward): PARITY UNPINNED,
    Panda3D (third party, absent) cannot be run here; SURVEY.md A.5."""
    y = fwd / np.linalg.norm(fwd)
    x = np.cross(y, up)
    x = x / np.linalg.norm(x)
    z = np.cross(x, y)
    return np.stack([x, y, z], axis=1)


def views_TC0_CV(TCO: np.ndarray, tCR: np.ndarray, offsets: np.ndarray):
    """lib3d/multiview.py:31-92 for one sample; returns the list of TC0_CV (4x4 float64)."""
    TCO = np.asarray(TCO, dtype=np.float64)
    tCR = np.asarray(tCR, dtype=np.float64)
    R, t = TCO[:3, :3], TCO[:3, 3]
    TOC = np.eye(4)
    TOC[:3, :3] = R.T
    TOC[:3, 3] = -R.T @ t
    if not np.isfinite(TOC).all() or not np.isfinite(TCO).all():
        TOC = np.eye(4)
        tCR = np.zeros(3)
    T_W_c0 = TOC @ _TCCGL
    c0 = T_W_c0[:3, 3]
    ref = TOC[:3, :3] @ tCR + TOC[:3, 3]
    radius = np.linalg.norm(tCR)
    up = T_W_c0[:3, 2]
    R_P = _look_at(ref - c0, up)
    inv_c0 = np.linalg.inv(T_W_c0)
    out = []
    for o in offsets:
        p = c0 + R_P @ (o * radius)
        T_W_n = np.eye(4)
        T_W_n[:3, :3] = _look_at(ref - p, up)
        T_W_n[:3, 3] = p
        out.append(_TCCGL @ (inv_c0 @ T_W_n) @ np.linalg.inv(_TCCGL))
    return out


def _inplane_rotations(TCV_O: torch.Tensor, remove_TCO_rendering: bool) -> torch.Tensor:
    """lib3d/multiview.py:236-246 (transforms3d.euler.euler2mat(0, 0, a) = rotation about z by a)."""
    assert remove_TCO_rendering
    TCV_O = TCV_O.unsqueeze(2).repeat(1, 1, 4, 1, 1)
    for idx, angle in enumerate([np.pi / 2, np.pi, 3 * np.pi / 2]):
        c, s = np.cos(angle), np.sin(angle)
        dR = torch.as_tensor(np.array([[c, -s, 0.0], [s, c, 0.0], [0.0, 0.0, 1.0]]), dtype=TCV_O.dtype)
        TCV_O[:, :, idx + 1, :3, :3] = dR @ TCV_O[:, :, idx + 1, :3, :3]
    return TCV_O.flatten(1, 2)


def make_TCO_multiview(TCO: torch.Tensor, tCR: torch.Tensor, multiview_type: str = "TCO+front_3views",
                       n_views: int = 4, remove_TCO_rendering: bool = False,
                       views_inplane_rotations: bool = False) -> torch.Tensor:
    """lib3d/multiview.py:165-246."""
    bsz = TCO.shape[0]
    if n_views == 1:
        out = TCO.unsqueeze(1).clone()
        return _inplane_rotations(out, remove_TCO_rendering) if views_inplane_rotations else out
    offsets = VIEW_OFFSETS[multiview_type]
    TCO_np, tCR_np = TCO.cpu().numpy(), tCR.cpu().numpy()
    all_views = []
    for b in range(bsz):
        views = [] if remove_TCO_rendering else [np.eye(4)]
        views += views_TC0_CV(TCO_np[b], tCR_np[b], offsets)
        all_views.append(views)
    TC0_CV = torch.as_tensor(np.stack(all_views), dtype=TCO.dtype)
    out = invert_transform_matrices(TC0_CV) @ TCO.unsqueeze(1)
    return _inplane_rotations(out, remove_TCO_rendering) if views_inplane_rotations else out


# ---------------------------------------------------------------------------------------------
# meshes / point sets / SO(3) grid
# ---------------------------------------------------------------------------------------------
def sample_point_ids(n_total: int, n_points: int) -> np.ndarray:
    """lib3d/mesh_ops.py:77-87 with deterministic=True."""
    assert n_points <= n_total
    return np.random.RandomState(0).choice(n_total, size=n_points, replace=False)


def pad_stack_points(point_list):
    """lib3d/rigid_mesh_database.py:171-200 (fill='select_random', deterministic)."""
    n_max = max(p.shape[0] for p in point_list)
    rs = np.random.RandomState(0)
    out = []
    for p in point_list:
        n_pad = n_max - len(p)
        if n_pad > 0:
            ids = rs.choice(np.arange(len(p)), size=n_pad)
            p = torch.cat((p, p[ids]), dim=0)
        out.append(p)
    return torch.stack(out)


def normalize_depth(depth: torch.Tensor, tCR: torch.Tensor, kind: Optional[str]) -> torch.Tensor:
    """models/pose_rigid.py:466-496."""
    z = tCR[:, 2][(...,) + (None,) * (depth.ndim - 1)]
    if kind == "tCR_scale":
        return depth / z
    if kind == "tCR_scale_clamp_center":
        return torch.clamp(depth / z, 0, 2) - 1
    if kind == "tCR_center_clamp":
        return torch.clamp(depth - z, -2, 2)
    if kind == "none":
        return depth
    raise ValueError(kind)
