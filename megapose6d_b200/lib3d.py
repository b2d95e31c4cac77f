"""Host-side mirror of the reference's lib3d functions on the hot path, executed by libmpx.so.

Same names, argument meaning and shapes as the reference (paths relative to
/root/reference/src/megapose): lib3d/cosypose_ops.py (TCO_init_from_boxes_autodepth_with_R,
pose_update_with_reference_point), lib3d/transform_ops.py (normalize_T), lib3d/rotations.py
(compute_rotation_matrix_from_ortho6d via update_pose), lib3d/multiview.py (make_TCO_multiview),
lib3d/cropping.py (crop_images / roi_align), lib3d/camera_geometry.py (boxes, K_crop).
All tensors must be CUDA float32; there is no CPU fallback.
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np
import torch

from . import _abi

def _sphere_26():
    """get_26_views_TCO_pos_sphere (lib3d/multiview.py:149-160)."""
    out = []
    for y in (0, 1, 2):
        for x in (0, -1, 1):
            for z in (0, 1, -1):
                if not (x == 0 and y == 1 and z == 0):
                    out.append([x, y, z])
    return out


VIEW_OFFSETS = {
    # camera positions wrt camera 0 in units of |tCR| (lib3d/multiview.py:95-160); make_TCO_multiview of the reference
    # accepts "TCO+front_1view", "TCO+front_3views" and "sphere_26views" (:197-232)
    "TCO+front_1view": [[0, 0, 0]],
    "TCO+front_3views": [[0, 0, 0], [1, 0, 0], [-1, 0, 0]],
    "TCO+front_5views": [[0, 0, 0], [1, 0, 0], [-1, 0, 0], [0, 0, 1], [0, 0, -1]],
    "sphere_26views": _sphere_26(),
}


def _f32(t: torch.Tensor) -> torch.Tensor:
    assert t.is_cuda, "libmpx operates on CUDA tensors only"
    return t.detach().to(torch.float32).contiguous()


def TCO_init_from_boxes_autodepth_with_R(boxes_2d: torch.Tensor, points: torch.Tensor, label_idx: torch.Tensor,
                                         K: torch.Tensor, R: torch.Tensor) -> torch.Tensor:
    """cosypose_ops.py:169-218.  `points` is the [L, Nv, 3] database indexed by `label_idx` [n] (int32)
    instead of a pre-gathered [n, Nv, 3] copy."""
    n = boxes_2d.shape[0]
    boxes_2d, points, K, R = _f32(boxes_2d), _f32(points), _f32(K), _f32(R)
    assert K.shape == (n, 3, 3) and R.shape == (n, 3, 3) and label_idx.dtype == torch.int32
    TCO = torch.empty(n, 4, 4, device=K.device, dtype=torch.float32)
    _abi.check(_abi.lib().mpx_pose_init_autodepth(_abi.ptr(points), points.shape[1], _abi.ptr(label_idx),
                                                  _abi.ptr(boxes_2d), _abi.ptr(K), _abi.ptr(R), n,
                                                  _abi.ptr(TCO), _abi.stream_ptr()))
    return TCO


def normalize_T(T: torch.Tensor) -> torch.Tensor:
    """transform_ops.py:117-119."""
    T = _f32(T)
    out = torch.empty_like(T)
    _abi.check(_abi.lib().mpx_normalize_T(_abi.ptr(T), T.shape[0], _abi.ptr(out), _abi.stream_ptr()))
    return out


def crop_geometry(points: torch.Tensor, label_idx: torch.Tensor, TCO: torch.Tensor, K: torch.Tensor,
                  tCR: torch.Tensor, im_size: Tuple[int, int], out_size: Tuple[int, int],
                  lamb: float = 1.4) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Boxes and crop intrinsics of PosePredictor.crop_inputs / compute_crops_multiview
    (models/pose_rigid.py:180-303).  Returns (boxes_rend, boxes_crop, K_crop)."""
    n = TCO.shape[0]
    points, TCO, K, tCR = _f32(points), _f32(TCO), _f32(K), _f32(tCR)
    boxes_rend = torch.empty(n, 4, device=TCO.device, dtype=torch.float32)
    boxes_crop = torch.empty_like(boxes_rend)
    K_crop = torch.empty(n, 3, 3, device=TCO.device, dtype=torch.float32)
    _abi.check(_abi.lib().mpx_crop_geometry(_abi.ptr(points), points.shape[1], _abi.ptr(label_idx), _abi.ptr(TCO),
                                            _abi.ptr(K), _abi.ptr(tCR), n, float(lamb), im_size[0], im_size[1],
                                            out_size[0], out_size[1], _abi.ptr(boxes_rend), _abi.ptr(boxes_crop),
                                            _abi.ptr(K_crop), _abi.stream_ptr()))
    return boxes_rend, boxes_crop, K_crop


def make_TCO_multiview(TCO: torch.Tensor, tCR: torch.Tensor, multiview_type: str = "TCO+front_3views",
                       n_views: int = 4, remove_TCO_rendering: bool = False,
                       views_inplane_rotations: bool = False) -> torch.Tensor:
    """multiview.py:165-246 -> TCV_O [bsz, n_views, 4, 4]."""
    TCO, tCR = _f32(TCO), _f32(tCR)
    n = TCO.shape[0]
    if n_views == 1:
        return _inplane(TCO.unsqueeze(1).clone(), remove_TCO_rendering) if views_inplane_rotations else TCO.unsqueeze(1).clone()
    if multiview_type not in VIEW_OFFSETS:
        raise ValueError(multiview_type)
    offs = np.ascontiguousarray(np.asarray(VIEW_OFFSETS[multiview_type], dtype=np.float32))
    n_extra = offs.shape[0]
    out = torch.empty(n, 1 + n_extra, 4, 4, device=TCO.device, dtype=torch.float32)
    _abi.check(_abi.lib().mpx_multiview_cameras(_abi.ptr(TCO), _abi.ptr(tCR), n, offs.ctypes.data, n_extra,
                                                _abi.ptr(out), _abi.stream_ptr()))
    if remove_TCO_rendering:
        out = out[:, 1:].contiguous()
    if views_inplane_rotations:
        return _inplane(out, remove_TCO_rendering)
    assert out.shape[1] == n_views, (out.shape, n_views)
    return out


def _inplane(TCV_O: torch.Tensor, remove_TCO_rendering: bool) -> torch.Tensor:
    """lib3d/multiview.py:236-246: every view also rotated in the image plane by 90, 180 and 270 degrees (the rotation
    block only, as in the reference); 4x the views."""
    assert remove_TCO_rendering
    out = TCV_O.unsqueeze(2).repeat(1, 1, 4, 1, 1)
    for idx, angle in enumerate((np.pi / 2, np.pi, 3 * np.pi / 2)):
        c, s = float(np.cos(angle)), float(np.sin(angle))
        # transforms3d.euler.euler2mat(0, 0, angle): rotation about z
        dR = torch.tensor([[c, -s, 0.0], [s, c, 0.0], [0.0, 0.0, 1.0]], device=TCV_O.device, dtype=TCV_O.dtype)
        out[:, :, idx + 1, :3, :3] = dR @ out[:, :, idx + 1, :3, :3]
    return out.flatten(1, 2).contiguous()


def update_pose(TCO: torch.Tensor, K_crop: torch.Tensor, pose_outputs: torch.Tensor, tCR: torch.Tensor) -> torch.Tensor:
    """PosePredictor.update_pose (models/pose_rigid.py:305-312)."""
    assert pose_outputs.shape[-1] == 9
    TCO, K_crop, pose_outputs, tCR = _f32(TCO), _f32(K_crop), _f32(pose_outputs), _f32(tCR)
    out = torch.empty_like(TCO)
    _abi.check(_abi.lib().mpx_pose_update(_abi.ptr(TCO), _abi.ptr(K_crop), _abi.ptr(pose_outputs), _abi.ptr(tCR),
                                          TCO.shape[0], _abi.ptr(out), _abi.stream_ptr()))
    return out


def topk_per_group(logits: torch.Tensor, k: int) -> torch.Tensor:
    """[B, M] logits -> [B, k] int32 indices (descending, ties to the lower index)."""
    logits = _f32(logits)
    b, m = logits.shape
    idx = torch.empty(b, k, device=logits.device, dtype=torch.int32)
    _abi.check(_abi.lib().mpx_topk_per_group(_abi.ptr(logits), b, m, k, _abi.ptr(idx), _abi.stream_ptr()))
    return idx


def image_to_nhwc4(images: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """[B, 3|4, H, W] float32 -> [B, H, W, 4] float32 (depth or 0 in channel 3)."""
    images = _f32(images)
    b, c, h, w = images.shape
    if out is None:
        out = torch.empty(b, h, w, 4, device=images.device, dtype=torch.float32)
    assert out.shape == (b, h, w, 4) and out.dtype == torch.float32 and out.is_contiguous()
    _abi.check(_abi.lib().mpx_image_to_nhwc4(_abi.ptr(images), b, c, h, w, _abi.ptr(out), _abi.stream_ptr()))
    return out


def crop_images(images_nhwc4: torch.Tensor, boxes: torch.Tensor, im_idx: Optional[torch.Tensor], n_channels: int,
                output_size: Tuple[int, int]) -> torch.Tensor:
    """crop_images (lib3d/cropping.py:113-144) -> [n, C, oh, ow] float32."""
    boxes = _f32(boxes)
    b, h, w, _ = images_nhwc4.shape
    n = boxes.shape[0]
    out = torch.empty(n, n_channels, output_size[0], output_size[1], device=boxes.device, dtype=torch.float32)
    _abi.check(_abi.lib().mpx_roi_align(_abi.ptr(images_nhwc4), b, h, w, _abi.ptr(im_idx), _abi.ptr(boxes), n,
                                        n_channels, output_size[0], output_size[1], _abi.ptr(out),
                                        _abi.stream_ptr()))
    return out
