"""Observation / detection / estimate containers of the inference API.

Mirrors src/megapose/inference/types.py:33-235 of the reference (`ObservationTensor`,
`InferenceConfig`, the `DetectionsType` / `PoseEstimatesType` aliases and
`assert_detections_valid`) so that existing callers keep working.

PoseEstimatesType.infos columns: label, batch_im_id, instance_id, hypothesis_id, coarse_logit,
coarse_score, pose_logit, pose_score, refiner_batch_idx, refiner_instance_idx; tensor `poses`
[B,4,4].  DetectionsType.infos columns: label, batch_im_id, instance_id[, score]; tensor `bboxes`
[B,4] (xmin, ymin, xmax, ymax).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch

from .tensor_collection import PandasTensorCollection

PoseEstimatesType = PandasTensorCollection
DetectionsType = PandasTensorCollection


def assert_detections_valid(detections: DetectionsType) -> None:
    for field in ("batch_im_id", "label", "instance_id"):
        assert field in detections.infos, f"detections.infos missing column {field}"
    assert "bboxes" in detections.tensors, "detections missing tensor bboxes."


@dataclass
class InferenceConfig:
    detection_type: str = "detector"  # ['detector', 'gt']
    coarse_estimation_type: str = "SO3_grid"
    SO3_grid_size: int = 576
    n_refiner_iterations: int = 5
    n_pose_hypotheses: int = 5
    run_depth_refiner: bool = False
    depth_refiner: Optional[str] = None  # ['icp', 'teaserpp']
    bsz_objects: int = 16
    bsz_images: int = 576


@dataclass
class ObservationTensor:
    """images: [B,C,H,W] float32, C=3 (rgb in [0,1]) or 4 (rgb + depth in metres); K: [B,3,3]."""

    images: torch.Tensor
    K: Optional[torch.Tensor] = None

    def cuda(self) -> "ObservationTensor":
        self.images = self.images.cuda()
        if self.K is not None:
            self.K = self.K.cuda()
        return self

    @property
    def batch_size(self) -> int:
        return self.images.shape[0]

    @property
    def channel_dim(self) -> int:
        return self.images.shape[1]

    @property
    def depth(self) -> torch.Tensor:
        assert self.channel_dim == 4
        return self.images[:, 3]

    def is_valid(self) -> bool:
        if self.images.ndim != 4 or self.channel_dim not in (3, 4):
            return False
        if self.K is not None and self.K.shape != torch.Size([self.batch_size, 3, 3]):
            return False
        if self.images.dtype != torch.float:
            return False
        return not bool(torch.max(self.images[:, :3]) > 1)

    @staticmethod
    def from_numpy(rgb: np.ndarray, depth: Optional[np.ndarray] = None,
                   K: Optional[np.ndarray] = None) -> "ObservationTensor":
        """rgb [H,W,3] uint8, depth [H,W] float (metres), K [3,3]."""
        assert rgb.dtype == np.uint8
        rgb_tensor = torch.as_tensor(np.array(rgb, copy=True)).float() / 255
        if rgb_tensor.shape[-1] == 3:
            rgb_tensor = rgb_tensor.permute(2, 0, 1)
        if depth is not None:
            img = torch.cat((rgb_tensor, torch.as_tensor(depth).unsqueeze(0)), dim=0)
        else:
            img = rgb_tensor
        return ObservationTensor(img.unsqueeze(0), torch.as_tensor(K).float().unsqueeze(0))

    @staticmethod
    def from_torch_batched(rgb: torch.Tensor, depth: Optional[torch.Tensor], K: torch.Tensor) -> "ObservationTensor":
        """rgb [B,3,H,W] uint8, depth [B,1,H,W] | [B,H,W] float, K [B,3,3]."""
        assert rgb.dtype == torch.uint8
        rgb = torch.as_tensor(rgb).float() / 255
        if depth is not None:
            if depth.ndim == 3:
                depth = depth.unsqueeze(1)
            img = torch.cat((rgb, depth), dim=1)
        else:
            img = rgb
        return ObservationTensor(img, torch.as_tensor(K).float())
