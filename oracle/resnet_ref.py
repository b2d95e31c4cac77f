"""oracle/resnet_ref.py -- torch fp32 restatement of the reference backbone + heads.

TEST INFRASTRUCTURE ONLY (see oracle/lib3d_ref.py header for the import rule).

Restates ResNet-34 as built by `resnet34(num_classes=512, n_input_channels=C)`
(models/torchvision_resnet.py:181-316,345-353; BasicBlock :74-120) followed by the single linear
head of PosePredictor (models/pose_rigid.py:120-130, 314-334), evaluated functionally from a
reference-format state dict (keys `backbone.*`, `pose_fc.*` | `views_logits_head.*`).
Validated against the reference's own nn.Module in tests/test_oracle_vs_reference.py.

`forward_act16_emulated` mirrors the engine's quantisation points (BN folded into 16-bit weights, 16-bit
activations between layers -- fp16 by default, bf16 for a library built with -DMPX_ACT_BF16 --, fp32
accumulation) so that kernel tests can use a tight tolerance; the fp32 `forward` is the parity target with
the tolerance `act16_forward_error_bound` states (ACT16_EPS per number format).
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import torch
import torch.nn.functional as F

LAYERS = [3, 4, 6, 3]
WIDTHS = [64, 128, 256, 512]
BN_EPS = 1e-5


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


def head_name(sd) -> str:
    return "pose_fc" if "pose_fc.weight" in sd else "views_logits_head"


def forward(sd: Dict[str, torch.Tensor], x: torch.Tensor) -> torch.Tensor:
    """fp32 forward: x [b,C,H,W] -> head output [b, 1|9]."""
    x = F.conv2d(x, sd["backbone.conv1.weight"], stride=2, padding=3)
    x = F.relu(_bn(x, sd, "backbone.bn1"))
    x = F.max_pool2d(x, kernel_size=3, stride=2, padding=1)
    for li, nb in enumerate(LAYERS):
        for b in range(nb):
            p = f"backbone.layer{li + 1}.{b}"
            stride = 2 if (b == 0 and li > 0) else 1
            identity = x
            out = F.conv2d(x, sd[p + ".conv1.weight"], stride=stride, padding=1)
            out = F.relu(_bn(out, sd, p + ".bn1"))
            out = F.conv2d(out, sd[p + ".conv2.weight"], stride=1, padding=1)
            out = _bn(out, sd, p + ".bn2")
            if (p + ".downsample.0.weight") in sd:
                identity = _bn(F.conv2d(x, sd[p + ".downsample.0.weight"], stride=stride), sd, p + ".downsample.1")
            x = F.relu(out + identity)
    x = torch.flatten(F.adaptive_avg_pool2d(x, (1, 1)), 1)
    x = F.linear(x, sd["backbone.fc.weight"], sd["backbone.fc.bias"])
    h = head_name(sd)
    return F.linear(x, sd[h + ".weight"], sd[h + ".bias"])


def pooled_features(sd: Dict[str, torch.Tensor], x: torch.Tensor) -> torch.Tensor:
    """fp32 backbone up to the global average pool: [b, 512] (the input of the folded fc o head map)."""
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


# Stated tolerance of the 16-bit engine against the fp32 network, as a fraction of the folded head's absolute-value
# condition bound.  Observed on the seeded test networks (tests/test_oracle_golden.py measures the emulation on the CPU,
# the GPU tests print the engine's own error): fp16 max |err| ~ 0.055 on logits of std 1.3 = 2^-14.7 of the bound, bf16
# 0.28 = 2^-12.3 of it; the stated figures leave a factor ~3.  (Round 1 used 2^-8, i.e. five logit standard deviations.)
ACT16_EPS = {torch.float16: 2.0 ** -13, torch.bfloat16: 2.0 ** -10}


def act16_forward_error_bound(sd: Dict[str, torch.Tensor], x: torch.Tensor, eps: float = None,
                              dtype: torch.dtype = torch.float16) -> torch.Tensor:
    """Stated tolerance for a 16-bit-activation network against this fp32 oracle: eps (default ACT16_EPS[dtype])
    times the absolute-value condition bound of the folded head, sum_i |W_ji| |pooled_i| (+|b_j|), per output
    [b, out_dim]: outputs that are small only through cancellation of large terms (random-weight networks) cannot
    be expected to agree to a fraction of their own size."""
    if eps is None:
        # pre-activation (WideResNet) nets round three tensors per block instead of two and carry an unnormalised
        # residual stream into the pooled features: observed error / condition bound is 2-3x that of the vanilla net
        eps = ACT16_EPS[dtype] * (3.0 if is_wide(sd) else 1.0)
    W, b = folded_head(sd)
    pooled = (pooled_features_wide(sd, x) if is_wide(sd) else pooled_features(sd, x)).double()
    return (eps * (pooled.abs() @ W.abs().t() + b.abs())).float()


def fold_bn(sd, conv: str, bn: str) -> Tuple[torch.Tensor, torch.Tensor]:
    """w' = w * gamma / sqrt(var + eps), b' = beta - mean * gamma / sqrt(var + eps) (float64)."""
    w = sd[conv + ".weight"].double()
    scale = sd[bn + ".weight"].double() / torch.sqrt(sd[bn + ".running_var"].double() + BN_EPS)
    return w * scale.view(-1, 1, 1, 1), sd[bn + ".bias"].double() - sd[bn + ".running_mean"].double() * scale


def conv_plan(sd) -> List[Tuple[str, str]]:
    """(conv, bn) names in the engine's execution order: stem, then per block conv1, conv2, [downsample]."""
    plan = [("backbone.conv1", "backbone.bn1")]
    for li, nb in enumerate(LAYERS):
        for b in range(nb):
            p = f"backbone.layer{li + 1}.{b}"
            plan.append((p + ".conv1", p + ".bn1"))
            plan.append((p + ".conv2", p + ".bn2"))
            if (p + ".downsample.0.weight") in sd:
                plan.append((p + ".downsample.0", p + ".downsample.1"))
    return plan


def forward_act16_emulated(sd: Dict[str, torch.Tensor], x: torch.Tensor,
                           dtype: torch.dtype = torch.float16) -> torch.Tensor:
    """Same network with the engine's quantisation points (runs on x.device, fp32 math); conversions saturate
    like the engine's (fp16: +-65504)."""
    dev = x.device
    lim = float(torch.finfo(dtype).max)

    def _q(t: torch.Tensor) -> torch.Tensor:
        return t.clamp(-lim, lim).to(dtype).float()

    def cw(conv, bn):
        w, b = fold_bn(sd, conv, bn)
        return _q(w.float()).to(dev), b.float().to(dev)

    if dev.type == "cuda" and torch.backends.cudnn.allow_tf32:
        # fp32 means fp32: with cuDNN's default TF32 convolutions the "exact" side of this comparison carries 2^-11 relative
        # errors of its own -- as large as the 16-bit roundings it emulates -- and which algorithm cuDNN picks (TF32 or not)
        # depends on the process' memory state, which made the comparison flaky at the stated half-bound
        with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
            return forward_act16_emulated(sd, x, dtype)
    x = _q(x)
    w, b = cw("backbone.conv1", "backbone.bn1")
    x = _q(F.relu(F.conv2d(x, w, b, stride=2, padding=3)))
    x = F.max_pool2d(x, kernel_size=3, stride=2, padding=1)
    for li, nb in enumerate(LAYERS):
        for bi in range(nb):
            p = f"backbone.layer{li + 1}.{bi}"
            stride = 2 if (bi == 0 and li > 0) else 1
            identity = x
            w, b = cw(p + ".conv1", p + ".bn1")
            out = _q(F.relu(F.conv2d(x, w, b, stride=stride, padding=1)))
            if (p + ".downsample.0.weight") in sd:
                wd, bd = cw(p + ".downsample.0", p + ".downsample.1")
                identity = _q(F.conv2d(x, wd, bd, stride=stride))
            w, b = cw(p + ".conv2", p + ".bn2")
            x = _q(F.relu(F.conv2d(out, w, b, stride=1, padding=1) + identity))
    pooled = x.flatten(2).mean(dim=-1)
    hw, hb = folded_head(sd)
    return F.linear(pooled, hw.float().to(dev), hb.float().to(dev))


def folded_head(sd) -> Tuple[torch.Tensor, torch.Tensor]:
    """head o fc as one linear map (float64): W = Wh Wfc, b = Wh bfc + bh."""
    h = head_name(sd)
    Wh, bh = sd[h + ".weight"].double(), sd[h + ".bias"].double()
    if "backbone.fc.weight" not in sd:  # WideResNet: the head sits directly on the pooled features
        return Wh, bh
    Wf, bf = sd["backbone.fc.weight"].double(), sd["backbone.fc.bias"].double()
    return Wh @ Wf, Wh @ bf + bh


# ---------------------------------------------------------------------------------------------
# pre-activation backbone: WideResNet34 / WideResNet18 of models/wide_resnet.py:29-126 (backbone_str "resnet34" /
# "resnet18", training/pose_models_cfg.py:110-116), width 1.  5x5 / stride-2 stem, BasicBlockV2
# (relu(bn1(x)) -> [bare 1x1 downsample on the activated tensor] -> conv1 -> relu(bn2) -> conv2 -> + residual), the 4-D
# output averaged in PosePredictor.net_forward (models/pose_rigid.py:323-328), no fc in front of the head.
# ---------------------------------------------------------------------------------------------
WIDE_LAYERS = {"resnet34": [3, 4, 6, 3], "resnet18": [2, 2, 2, 2]}


def is_wide(sd) -> bool:
    return "backbone.layer1.0.bn1.weight" in sd and "backbone.layer1.0.conv1.weight" in sd and "backbone.fc.weight" not in sd


def wide_layers(sd) -> List[int]:
    return [sum(1 for k in sd if k.startswith(f"backbone.layer{li}.") and k.endswith(".conv1.weight")) for li in (1, 2, 3, 4)]


def init_state_dict_wide(n_inputs: int, head: str, head_dim: int, seed: int = 0, backbone_str: str = "resnet34"):
    """Seeded random weights in the checkpoint layout of a WideResNet backbone + head."""
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


def _wide_trunk(sd, x, q=lambda t: t, fold: bool = False, dev=None):
    """Shared structure of the fp32 and the emulated forward: returns the last block's output [b, 512, h, w]."""
    def cw(conv, bn=None):
        if fold:
            if bn is None:
                w, b = sd[conv + ".weight"].double(), torch.zeros(sd[conv + ".weight"].shape[0], dtype=torch.float64)
            else:
                w, b = fold_bn(sd, conv, bn)
            return q(w.float()).to(dev), b.float().to(dev)
        return sd[conv + ".weight"], None

    def affine(name):
        scale = sd[name + ".weight"].double() / torch.sqrt(sd[name + ".running_var"].double() + BN_EPS)
        shift = sd[name + ".bias"].double() - sd[name + ".running_mean"].double() * scale
        return scale.float().view(1, -1, 1, 1).to(x.device), shift.float().view(1, -1, 1, 1).to(x.device)

    if fold:
        w, b = cw("backbone.conv1", "backbone.bn1")
        x = q(F.relu(F.conv2d(q(x), w, b, stride=2, padding=2)))
    else:
        x = F.relu(_bn(F.conv2d(x, sd["backbone.conv1.weight"], stride=2, padding=2), sd, "backbone.bn1"))
    x = F.max_pool2d(x, kernel_size=3, stride=2, padding=1)
    for li, nb in enumerate(wide_layers(sd)):
        for bi in range(nb):
            p = f"backbone.layer{li + 1}.{bi}"
            stride = 2 if (bi == 0 and li > 0) else 1
            if fold:
                sc, sh = affine(p + ".bn1")
                a = q(F.relu(x * sc + sh))  # the engine's elementwise pass: fp32 affine of the 16-bit tensor, one rounding
                res = x
                if (p + ".downsample.weight") in sd:
                    wd, bd = cw(p + ".downsample")
                    res = q(F.conv2d(a, wd, bd, stride=stride))
                w1, b1 = cw(p + ".conv1", p + ".bn2")
                y = q(F.relu(F.conv2d(a, w1, b1, stride=stride, padding=1)))
                w2, b2 = cw(p + ".conv2")
                x = q(F.conv2d(y, w2, b2, stride=1, padding=1) + res)
            else:
                a = F.relu(_bn(x, sd, p + ".bn1"))
                res = F.conv2d(a, sd[p + ".downsample.weight"], stride=stride) if (p + ".downsample.weight") in sd else x
                y = F.relu(_bn(F.conv2d(a, sd[p + ".conv1.weight"], stride=stride, padding=1), sd, p + ".bn2"))
                x = F.conv2d(y, sd[p + ".conv2.weight"], stride=1, padding=1) + res
    return x


def forward_wide(sd: Dict[str, torch.Tensor], x: torch.Tensor) -> torch.Tensor:
    """fp32 forward of a WideResNet backbone + head: x [b,C,H,W] -> [b, 1|9]."""
    feat = _wide_trunk(sd, x).flatten(2).mean(dim=-1)
    h = head_name(sd)
    return F.linear(feat, sd[h + ".weight"], sd[h + ".bias"])


def pooled_features_wide(sd, x):
    return _wide_trunk(sd, x).flatten(2).mean(dim=-1)


def forward_wide_act16_emulated(sd, x, dtype: torch.dtype = torch.float16) -> torch.Tensor:
    """The engine's quantisation points for the pre-activation backbone (16-bit weights and stored tensors, fp32 math)."""
    if x.device.type == "cuda" and torch.backends.cudnn.allow_tf32:  # see forward_act16_emulated
        with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
            return forward_wide_act16_emulated(sd, x, dtype)
    lim = float(torch.finfo(dtype).max)

    def q(t):
        return t.clamp(-lim, lim).to(dtype).float()

    feat = _wide_trunk(sd, x, q=q, fold=True, dev=x.device).flatten(2).mean(dim=-1)
    h = head_name(sd)
    return F.linear(feat, sd[h + ".weight"].float().to(x.device), sd[h + ".bias"].float().to(x.device))


def forward_any(sd, x):
    """fp32 forward of either backbone family."""
    return forward_wide(sd, x) if is_wide(sd) else forward(sd, x)
