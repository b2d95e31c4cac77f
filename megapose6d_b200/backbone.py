"""ResNet-34 backbone + head as a libmpx network handle.

Takes the parameters of the reference's `vanilla_resnet34` backbone
(src/megapose/models/torchvision_resnet.py:181-316, created by
training/pose_models_cfg.py:106-109 with `num_classes=512, n_input_channels=C`) and of the head
(`pose_fc` or `views_logits_head`, models/pose_rigid.py:120-130) in the reference's state-dict
layout and repacks them once for the tcgen05 kernels:
  * eval-mode BatchNorm folded into the preceding conv (w' = w*g/sqrt(var+eps), b' = beta - mean*g/sqrt(var+eps));
  * conv weights OIHW fp32 -> [C_out, R*S*C_in] in the library's 16-bit type (`_abi.act_dtype()`: fp16 unless the
    library was built for bf16), K ordered (r, s, c);
  * the 7x7/s2 stem rewritten as a 4x4/s1 conv over the space-to-depth input (channels padded to
    c_pad = a multiple of 16 (16 | 32 for the released models), four sub-pixels -> 4 * c_pad input channels);
  * avgpool -> fc(512x512) -> head(512 x 1|9) folded into one linear map (there is no
    non-linearity between fc and the head, models/pose_rigid.py:323-334).
"""
from __future__ import annotations

import ctypes
from typing import Dict, List, Optional, Tuple

import torch

from . import _abi

LAYERS = [3, 4, 6, 3]
BN_EPS = 1e-5


def _fold(sd: Dict[str, torch.Tensor], conv: str, bn: str) -> Tuple[torch.Tensor, torch.Tensor]:
    w = sd[conv + ".weight"].detach().double().cpu()
    g = sd[bn + ".weight"].detach().double().cpu()
    beta = sd[bn + ".bias"].detach().double().cpu()
    mean = sd[bn + ".running_mean"].detach().double().cpu()
    var = sd[bn + ".running_var"].detach().double().cpu()
    scale = g / torch.sqrt(var + BN_EPS)
    return w * scale.view(-1, 1, 1, 1), beta - mean * scale


def _pack(w: torch.Tensor) -> torch.Tensor:
    """[co, ci, r, s] -> [co, r*s*ci] with k = (r, s, c)."""
    co = w.shape[0]
    return w.permute(0, 2, 3, 1).reshape(co, -1).contiguous()


def _stem_s2d(w: torch.Tensor, c_pad: int) -> torch.Tensor:
    """7x7/s2/p3 weights [64, C, 7, 7] -> 4x4/s1 (pad 2 low, 1 high) weights over the s2d input:
    w2[co, by, bx, (dy*2+dx)*c_pad + c] = w[co, c, 2*by+dy-1, 2*bx+dx-1] (zero outside 0..6)."""
    co, c, _, _ = w.shape
    w2 = torch.zeros(co, 4, 4, 4 * c_pad, dtype=w.dtype)
    for by in range(4):
        for dy in range(2):
            kh = 2 * by + dy - 1
            if not 0 <= kh <= 6:
                continue
            for bx in range(4):
                for dx in range(2):
                    kw = 2 * bx + dx - 1
                    if not 0 <= kw <= 6:
                        continue
                    base = (dy * 2 + dx) * c_pad
                    w2[:, by, bx, base:base + c] = w[:, :, kh, kw]
    return w2.reshape(co, -1).contiguous()


class ResNet34Engine:
    """Owns the repacked device weights and the mpx_net handle; `forward(x)` runs the whole network."""

    def __init__(self, state_dict: Dict[str, torch.Tensor], n_inputs: int, head: str, device="cuda"):
        sd = state_dict
        self.n_inputs = n_inputs
        self.n_features = 512
        self.c_pad = 16 * ((n_inputs + 15) // 16)  # 16 (coarse), 32 (refiner) for the released models
        assert n_inputs <= 256, f"n_inputs={n_inputs} > 256 is not supported"
        assert sd["backbone.conv1.weight"].shape[1] == n_inputs, "checkpoint / config channel mismatch"
        self.device = torch.device(device)
        self.act_dtype = _abi.act_dtype()
        self._weights: List[torch.Tensor] = []
        self._biases: List[torch.Tensor] = []

        def add(wmat: torch.Tensor, bias: torch.Tensor) -> None:
            self._weights.append(wmat.to(torch.float32).to(self.device).to(self.act_dtype).contiguous())
            self._biases.append(bias.to(torch.float32).to(self.device).contiguous())

        w, b = _fold(sd, "backbone.conv1", "backbone.bn1")
        add(_stem_s2d(w, self.c_pad), b)
        for li, nb in enumerate(LAYERS):
            for bi in range(nb):
                p = f"backbone.layer{li + 1}.{bi}"
                w, b = _fold(sd, p + ".conv1", p + ".bn1")
                add(_pack(w), b)
                w, b = _fold(sd, p + ".conv2", p + ".bn2")
                add(_pack(w), b)
                if (p + ".downsample.0.weight") in sd:
                    w, b = _fold(sd, p + ".downsample.0", p + ".downsample.1")
                    add(_pack(w), b)
        Wh, bh = sd[head + ".weight"].detach().double().cpu(), sd[head + ".bias"].detach().double().cpu()
        Wf, bf = sd["backbone.fc.weight"].detach().double().cpu(), sd["backbone.fc.bias"].detach().double().cpu()
        self.out_dim = Wh.shape[0]
        self.head_w = (Wh @ Wf).float().to(self.device).contiguous()
        self.head_b = (Wh @ bf + bh).float().to(self.device).contiguous()

        n = len(self._weights)
        wp = (ctypes.c_void_p * n)(*[t.data_ptr() for t in self._weights])
        bp = (ctypes.c_void_p * n)(*[t.data_ptr() for t in self._biases])
        handle = ctypes.c_void_p()
        _abi.check(_abi.lib().mpx_net_create(self.c_pad, self.out_dim, wp, bp, n, _abi.ptr(self.head_w),
                                             _abi.ptr(self.head_b), ctypes.byref(handle)))
        self._handle = handle
        self._workspace: Optional[torch.Tensor] = None
        # workspaces that were outgrown: CUDA graphs captured by the callers (PosePredictor._iterate_graphed,
        # PoseEstimator._coarse_stage_graphed) have their addresses baked in, so they are kept alive, never freed;
        # growth is geometric, the retired ones therefore sum to less than the live one
        self._retired_workspaces: List[torch.Tensor] = []
        self._out_cache: Dict[int, torch.Tensor] = {}

    def __del__(self):
        try:
            if getattr(self, "_handle", None) is not None:
                _abi.lib().mpx_net_destroy(self._handle)
        except Exception:  # noqa: BLE001
            pass

    def alloc_input(self, n: int, h: int, w: int) -> torch.Tensor:
        """Zero-initialised network input tensor [n, h/2, w/2, 4*c_pad] fp16|bf16 (pad channels stay 0)."""
        return torch.zeros(n, h // 2, w // 2, 4 * self.c_pad, device=self.device, dtype=self.act_dtype)

    def pack_input(self, x_nchw: torch.Tensor) -> torch.Tensor:
        """[n, C, h, w] float -> space-to-depth 16-bit input (for tests and the non-fused API path)."""
        n, c, h, w = x_nchw.shape
        assert c == self.n_inputs
        x = torch.zeros(n, self.c_pad, h, w, device=self.device, dtype=torch.float32)
        x[:, :c] = x_nchw.to(self.device).float()
        x = x.view(n, self.c_pad, h // 2, 2, w // 2, 2).permute(0, 2, 4, 3, 5, 1)  # n, h/2, w/2, dy, dx, c
        return x.reshape(n, h // 2, w // 2, 4 * self.c_pad).clamp(-65504.0, 65504.0).to(self.act_dtype).contiguous()

    def forward(self, x: torch.Tensor, h: int, w: int) -> torch.Tensor:
        """x: network input tensor for n samples of size h x w -> [n, out_dim] float32."""
        n = x.shape[0]
        assert x.dtype == self.act_dtype and x.is_contiguous() and x.shape == (n, h // 2, w // 2, 4 * self.c_pad)
        need = _abi.lib().mpx_net_workspace_bytes(self._handle, n, h, w)
        if self._workspace is None or self._workspace.numel() < need:
            if self._workspace is not None:
                self._retired_workspaces.append(self._workspace)
                need = max(need, 2 * self._workspace.numel())
            self._workspace = torch.empty(need, dtype=torch.uint8, device=self.device)
        # persistent output buffer per batch size: (x, out, workspace, shape) identify the cached CUDA graph
        out = self._out_cache.get(n)
        if out is None:
            out = self._out_cache[n] = torch.empty(n, self.out_dim, device=self.device, dtype=torch.float32)
        _abi.check(_abi.lib().mpx_net_forward(self._handle, _abi.ptr(x), n, h, w, _abi.ptr(out),
                                              _abi.ptr(self._workspace), self._workspace.numel(), _abi.stream_ptr()))
        return out.clone()

    def __call__(self, x_nchw: torch.Tensor) -> torch.Tensor:
        n, c, h, w = x_nchw.shape
        return self.forward(self.pack_input(x_nchw), h, w)
