"""Seeded random network weights in the reference's checkpoint layout (no checkpoint is available offline).

Workload generation for bench.py, tools/ and tests/ -- neither the product (megapose6d_b200/) nor the oracle.  The
state dicts follow `resnet34(num_classes=512, n_input_channels=C)` + the single linear head
(models/torchvision_resnet.py:181-316, models/pose_rigid.py:120-130); the head is conditioned on a calibration batch
(plain torch fp32 on the host, run once per configuration) so that random weights give O(1) logits and small pose
updates.
"""
from __future__ import annotations

from typing import Dict

import torch
import torch.nn.functional as F

LAYERS = [3, 4, 6, 3]
WIDTHS = [64, 128, 256, 512]
BN_EPS = 1e-5

COARSE_CFG = dict(n_rendered_views=1, multiview_type="TCO", render_normals=True, render_depth=False, input_depth=False,
                  predict_rendered_views_logits=True, predict_pose_update=False, remove_TCO_rendering=False,
                  depth_normalization_type="tCR_scale_clamp_center")
REFINER_CFG = dict(n_rendered_views=4, multiview_type="TCO+front_3views", render_normals=True, render_depth=False,
                   input_depth=False, predict_rendered_views_logits=False, predict_pose_update=True,
                   remove_TCO_rendering=False, depth_normalization_type="tCR_scale_clamp_center")
REFINER_RGBD_CFG = dict(REFINER_CFG, render_depth=True, input_depth=True)


def n_inputs(cfg) -> int:
    per_view = 3 + (3 if cfg.get("render_normals", True) else 0) + int(cfg["render_depth"])  # pose_models_cfg.py:95-103
    return (3 + int(cfg["input_depth"])) + per_view * cfg["n_rendered_views"]


def init_state_dict(n_inputs: int, head: str, head_dim: int, seed: int = 0) -> Dict[str, torch.Tensor]:
    """Seeded random weights in the reference checkpoint layout (kaiming fan_out convs as
    torchvision_resnet.py:232-237, non-trivial BN statistics so that folding is exercised)."""
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}

    def conv(name, co, ci, k):
        std = (2.0 / (co * k * k)) ** 0.5
        sd[name + ".weight"] = torch.randn(co, ci, k, k, generator=g) * std

    def bn(name, c):
        sd[name + ".weight"] = 0.5 + torch.rand(c, generator=g)
        sd[name + ".bias"] = 0.2 * torch.randn(c, generator=g)
        sd[name + ".running_mean"] = 0.2 * torch.randn(c, generator=g)
        sd[name + ".running_var"] = 0.5 + torch.rand(c, generator=g)
        sd[name + ".num_batches_tracked"] = torch.tensor(1)

    conv("backbone.conv1", 64, n_inputs, 7)
    bn("backbone.bn1", 64)
    inplanes = 64
    for li, (nb, width) in enumerate(zip(LAYERS, WIDTHS)):
        for b in range(nb):
            p = f"backbone.layer{li + 1}.{b}"
            stride = 2 if (b == 0 and li > 0) else 1
            conv(p + ".conv1", width, inplanes, 3)
            bn(p + ".bn1", width)
            conv(p + ".conv2", width, width, 3)
            bn(p + ".bn2", width)
            if stride != 1 or inplanes != width:
                conv(p + ".downsample.0", width, inplanes, 1)
                bn(p + ".downsample.1", width)
            inplanes = width
    sd["backbone.fc.weight"] = torch.randn(512, 512, generator=g) * (1.0 / 512) ** 0.5
    sd["backbone.fc.bias"] = 0.1 * torch.randn(512, generator=g)
    sd[head + ".weight"] = torch.randn(head_dim, 512, generator=g) * (1.0 / 512) ** 0.5
    sd[head + ".bias"] = 0.1 * torch.randn(head_dim, generator=g)
    return sd



def _bn(x, sd, name):
    return F.batch_norm(x, sd[name + ".running_mean"], sd[name + ".running_var"], sd[name + ".weight"],
                        sd[name + ".bias"], training=False, eps=BN_EPS)


def _pooled_features(sd: Dict[str, torch.Tensor], x: torch.Tensor) -> torch.Tensor:
    """fp32 backbone up to the global average pool [b, 512] (calibration only)."""
    x = F.conv2d(x, sd["backbone.conv1.weight"], stride=2, padding=3)
    x = F.relu(_bn(x, sd, "backbone.bn1"))
    x = F.max_pool2d(x, kernel_size=3, stride=2, padding=1)
    for li, nb in enumerate(LAYERS):
        for b in range(nb):
            p = f"backbone.layer{li + 1}.{b}"
            stride = 2 if (b == 0 and li > 0) else 1
            identity = x
            out = F.relu(_bn(F.conv2d(x, sd[p + ".conv1.weight"], stride=stride, padding=1), sd, p + ".bn1"))
            out = _bn(F.conv2d(out, sd[p + ".conv2.weight"], stride=1, padding=1), sd, p + ".bn2")
            if (p + ".downsample.0.weight") in sd:
                identity = _bn(F.conv2d(x, sd[p + ".downsample.0.weight"], stride=stride), sd, p + ".downsample.1")
            x = F.relu(out + identity)
    return torch.flatten(F.adaptive_avg_pool2d(x, (1, 1)), 1)


WIDE_LAYERS = {"resnet34": [3, 4, 6, 3], "resnet18": [2, 2, 2, 2]}


def init_state_dict_wide(n_inputs: int, head: str, head_dim: int, seed: int = 0, backbone_str: str = "resnet34"):
    """Seeded random weights in the checkpoint layout of a WideResNet backbone (models/wide_resnet.py:59-126) + head."""
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}

    def conv(name, co, ci, k):
        sd[name + ".weight"] = torch.randn(co, ci, k, k, generator=g) * (2.0 / (co * k * k)) ** 0.5

    def bn(name, c):
        sd[name + ".weight"] = 0.5 + torch.rand(c, generator=g)
        sd[name + ".bias"] = 0.2 * torch.randn(c, generator=g)
        sd[name + ".running_mean"] = 0.2 * torch.randn(c, generator=g)
        sd[name + ".running_var"] = 0.5 + torch.rand(c, generator=g)
        sd[name + ".num_batches_tracked"] = torch.tensor(1)

    conv("backbone.conv1", 64, n_inputs, 5)
    bn("backbone.bn1", 64)
    inplanes = 64
    for li, (nb, width) in enumerate(zip(WIDE_LAYERS[backbone_str], WIDTHS)):
        for b in range(nb):
            p = f"backbone.layer{li + 1}.{b}"
            stride = 2 if (b == 0 and li > 0) else 1
            bn(p + ".bn1", inplanes)
            conv(p + ".conv1", width, inplanes, 3)
            bn(p + ".bn2", width)
            conv(p + ".conv2", width, width, 3)
            if stride != 1 or inplanes != width:
                conv(p + ".downsample", width, inplanes, 1)
            inplanes = width
    sd[head + ".weight"] = torch.randn(head_dim, 512, generator=g) * (1.0 / 512) ** 0.5
    sd[head + ".bias"] = 0.1 * torch.randn(head_dim, generator=g)
    return sd


def _pooled_features_wide(sd: Dict[str, torch.Tensor], x: torch.Tensor) -> torch.Tensor:
    """fp32 pre-activation backbone up to the spatial mean [b, 512] (calibration only)."""
    x = F.relu(_bn(F.conv2d(x, sd["backbone.conv1.weight"], stride=2, padding=2), sd, "backbone.bn1"))
    x = F.max_pool2d(x, kernel_size=3, stride=2, padding=1)
    li = 0
    while f"backbone.layer{li + 1}.0.conv1.weight" in sd:
        b = 0
        while f"backbone.layer{li + 1}.{b}.conv1.weight" in sd:
            p = f"backbone.layer{li + 1}.{b}"
            stride = 2 if (b == 0 and li > 0) else 1
            a = F.relu(_bn(x, sd, p + ".bn1"))
            res = F.conv2d(a, sd[p + ".downsample.weight"], stride=stride) if (p + ".downsample.weight") in sd else x
            y = F.relu(_bn(F.conv2d(a, sd[p + ".conv1.weight"], stride=stride, padding=1), sd, p + ".bn2"))
            x = F.conv2d(y, sd[p + ".conv2.weight"], stride=1, padding=1) + res
            b += 1
        li += 1
    return x.flatten(2).mean(dim=-1)


def calibration_batch(c, seed, n=4, h=240, w=320):
    """Smooth images in [0,1]; half of them with the render channels masked to a blob on black, like real inputs."""
    g = torch.Generator().manual_seed(1000 + seed)
    x = torch.rand(n, c, h // 8, w // 8, generator=g)
    x = torch.nn.functional.interpolate(x, size=(h, w), mode="bilinear", align_corners=False)
    yy, xx = torch.meshgrid(torch.linspace(-1, 1, h), torch.linspace(-1, 1, w), indexing="ij")
    blob = ((xx ** 2 + yy ** 2) < 0.4).float()
    x[n // 2:, 3:] *= blob
    return x.clamp(0, 1)


_SD_CACHE = {}


def make_state_dict(cfg, seed=0):
    """Seeded random weights in the checkpoint layout with a conditioned head: the head is made orthogonal to the
    dominant feature direction of a calibration batch and scaled so that coarse logits are O(1) and pose updates are
    small (R ~ I, v_z ~ 1) -- random heads otherwise produce |logit| ~ 300 and 20x depth jumps."""
    key = (tuple(sorted(cfg.items())), seed)
    if key in _SD_CACHE:
        return dict(_SD_CACHE[key])
    head = "pose_fc" if cfg["predict_pose_update"] else "views_logits_head"
    dim = 9 if cfg["predict_pose_update"] else cfg["n_rendered_views"]
    c = n_inputs(cfg)
    wide = cfg.get("backbone_str", "vanilla_resnet34") in ("resnet34", "resnet18")
    if wide:
        sd = init_state_dict_wide(c, head, dim, seed=seed, backbone_str=cfg["backbone_str"])
        with torch.no_grad():
            feats = _pooled_features_wide(sd, calibration_batch(c, seed))
    else:
        sd = init_state_dict(c, head, dim, seed=seed)
        with torch.no_grad():
            pooled = _pooled_features(sd, calibration_batch(c, seed))
            feats = torch.nn.functional.linear(pooled, sd["backbone.fc.weight"], sd["backbone.fc.bias"])
    v = torch.linalg.svd(feats, full_matrices=False)[2][0]
    W = sd[head + ".weight"]
    W = W - (W @ v).unsqueeze(1) * v.unsqueeze(0)
    raw = feats @ W.t()
    W = W * ((0.02 if cfg["predict_pose_update"] else 1.5) / (raw - raw.mean(0)).std().clamp_min(1e-12))
    sd[head + ".weight"] = W
    offset = (feats @ W.t()).mean(0)
    if cfg["predict_pose_update"]:
        sd[head + ".bias"] = torch.tensor([1.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0, 1.0]) - offset
    else:
        sd[head + ".bias"] = -offset
    _SD_CACHE[key] = dict(sd)
    return sd
