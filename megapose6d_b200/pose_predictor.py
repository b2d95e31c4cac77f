"""Render-and-compare model: crop -> render -> ResNet-34 -> head -> pose update, all on the GPU.

Drop-in for the reference's PosePredictor (src/megapose/models/pose_rigid.py:81-708): same
constructor keywords, `forward` / `forward_coarse` / `forward_coarse_tensor` / `crop_inputs` /
`compute_crops_multiview` / `render_images_multiview` / `net_forward` / `update_pose` /
`normalize_images` with the same argument meaning and output structure (PosePredictorOutput).

What differs underneath: one fused path.  The crop kernel and the rasteriser write 16-bit channels
straight into the network input tensor (no fp32 NCHW intermediates, no torch.cat, no host round
trip for the multi-view cameras); the fp32 `images_crop` / `renders` tensors of the reference's
outputs are only materialised on request (`keep_images=True` or `return_debug_data=True`).
`forward` / `forward_coarse` additionally accept `batch_im_ids` so that callers can pass the
un-replicated observation images (the reference replicates the frame once per hypothesis,
inference/pose_estimator.py:389).
"""
from __future__ import annotations

import time
from collections import defaultdict
from dataclasses import dataclass
from typing import Any, Dict, List, Optional, Tuple

import torch
from torch import nn

from . import lib3d
from .backbone import ResNet34Engine
from .meshes import BatchedMeshes
from .renderer import (DEPTH_NORM_KINDS, DEPTH_NORM_SHIFT, RASTER_POINT_LIGHTS, BatchRenderer, Panda3dLightData,
                       make_scene_lights)


@dataclass
class PosePredictorOutput:
    TCO_output: torch.Tensor
    TCO_input: torch.Tensor
    renders: Optional[torch.Tensor]
    images_crop: Optional[torch.Tensor]
    TCV_O_input: torch.Tensor
    KV_crop: torch.Tensor
    tCR: torch.Tensor
    labels: List[str]
    K: torch.Tensor
    K_crop: torch.Tensor
    network_outputs: Dict[str, torch.Tensor]
    boxes_rend: torch.Tensor
    boxes_crop: torch.Tensor
    renderings_logits: torch.Tensor
    timing_dict: Dict[str, float]


class _Timer:
    """CUDA-event timer when enabled, wall clock otherwise (reference: training/utils.py:224-264)."""

    def __init__(self, enabled: bool):
        self.enabled = enabled
        self.t0 = 0.0
        self.elapsed_s = 0.0

    def start(self):
        if self.enabled:
            self.e0, self.e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            self.e0.record()
        else:
            self.t0 = time.time()

    def stop(self):
        if self.enabled:
            self.e1.record()
            torch.cuda.synchronize()
            self.elapsed_s = self.e0.elapsed_time(self.e1) / 1000.0
        else:
            self.elapsed_s = time.time() - self.t0
        return self.elapsed_s


class PosePredictor(nn.Module):
    def __init__(
        self,
        backbone: ResNet34Engine,
        renderer: BatchRenderer,
        mesh_db: BatchedMeshes,
        render_size: Tuple[int, int] = (240, 320),
        multiview_type: str = "front_3views",
        views_inplane_rotations: bool = False,
        remove_TCO_rendering: bool = False,
        predict_pose_update: bool = True,
        predict_rendered_views_logits: bool = False,
        render_normals: bool = True,
        n_rendered_views: int = 1,
        input_depth: bool = False,
        render_depth: bool = False,
        depth_normalization_type: Optional[str] = None,
    ):
        super().__init__()
        self.backbone = backbone
        self.renderer = renderer
        self.mesh_db = mesh_db
        self.render_size = tuple(render_size)
        self.n_rendered_views = n_rendered_views
        self.input_depth = input_depth
        self.multiview_type = multiview_type
        self.views_inplane_rotations = views_inplane_rotations
        self.render_normals = render_normals
        self.render_depth = render_depth
        self.depth_normalization_type = depth_normalization_type
        self.predict_rendered_views_logits = predict_rendered_views_logits
        self.remove_TCO_rendering = remove_TCO_rendering
        self.predict_pose_update = predict_pose_update
        if views_inplane_rotations:
            assert remove_TCO_rendering, "views_inplane_rotations needs remove_TCO_rendering (lib3d/multiview.py:237)"
        if predict_pose_update:
            assert backbone.out_dim == 9 and not predict_rendered_views_logits
        if predict_rendered_views_logits:
            assert backbone.out_dim == n_rendered_views
        self._input_rgb_dims = [0, 1, 2]
        self._input_depth_dims = [3] if input_depth else []
        self._n_single_render_channels = 3 + (3 if render_normals else 0) + (1 if render_depth else 0)
        n_inputs = (3 + (1 if input_depth else 0)) + self._n_single_render_channels * n_rendered_views
        assert n_inputs == backbone.n_inputs, (n_inputs, backbone.n_inputs)
        if (input_depth or render_depth) and depth_normalization_type not in DEPTH_NORM_KINDS:
            raise ValueError(f"Unknown depth_normalization_type = {depth_normalization_type}")
        # what the fused crop / raster kernels need to know about this configuration (include/mpx.h)
        self._depth_norm_kind = DEPTH_NORM_KINDS.get(depth_normalization_type, 3)
        self._raster_flags = (0 if render_normals else RASTER_POINT_LIGHTS) | (self._depth_norm_kind << DEPTH_NORM_SHIFT)
        self.debug = False
        self.keep_images = False  # materialise fp32 images_crop / renders in the outputs
        self.max_batch = 1152     # hypotheses per fused launch (memory: ~9 MB each at 240x320)
        self._nhwc4_bufs: Dict[Tuple[int, ...], torch.Tensor] = {}
        self._graphs: Dict[Any, Dict[str, Any]] = {}
        self.use_cuda_graphs = True   # replay the refinement loop as one CUDA graph for small batches
        # stream the small-batch graphs are captured on (None: torch's side stream).  Kernel nodes inherit its priority: a
        # high-priority stream here makes the latency-bound launches schedule ahead of another frame's queued CTAs
        self.graph_capture_stream = None
        self.graph_max_batch = 1024
        self._x_cache: Dict[Tuple[int, int, int], torch.Tensor] = {}
        self.graph_epoch = 0  # bumped whenever buffers that captured graphs point into are released

    # ------------------------------------------------------------------------------------------
    # helpers
    # ------------------------------------------------------------------------------------------
    @property
    def input_rgb_dims(self) -> List[int]:
        return self._input_rgb_dims

    @property
    def input_depth_dims(self) -> List[int]:
        return self._input_depth_dims

    def _nhwc4(self, images: torch.Tensor, refresh: bool = False) -> torch.Tensor:
        """NHWC4 copy of the frame(s) in a persistent buffer per shape (stable address for the captured graphs).
        Every public entry point refreshes it once (`refresh=True`); the inner steps only look it up."""
        key = tuple(images.shape)
        buf = self._nhwc4_bufs.get(key)
        if buf is None:
            if len(self._nhwc4_bufs) >= 4:
                self._nhwc4_bufs.clear()
            b, _, h, w = images.shape
            buf = self._nhwc4_bufs[key] = torch.empty(b, h, w, 4, device=images.device, dtype=torch.float32)
            refresh = True
        if refresh:
            lib3d.image_to_nhwc4(images, out=buf)
        return buf

    def _input_buffer(self, n: int, h: int, w: int) -> torch.Tensor:
        """Persistent network input per (batch size, render size).  Captured graphs (here and in
        PoseEstimator._coarse_stage_graphed) have these addresses baked in, so a buffer is only ever released together
        with every graph that may point into it: `graph_epoch` is part of the owners' graph keys."""
        x = self._x_cache.get((n, h, w))
        if x is None:
            if len(self._x_cache) >= 32:
                assert not torch.cuda.is_current_stream_capturing(), "input buffers must exist before capture"
                self._graphs.clear()
                self._x_cache.clear()
                self.graph_epoch += 1
            x = self._x_cache[(n, h, w)] = self.backbone.alloc_input(n, h, w)
        return x

    def _label_idx(self, labels: List[str], device) -> torch.Tensor:
        return self.mesh_db.label_ids(labels, device)

    # ------------------------------------------------------------------------------------------
    # reference API: pieces
    # ------------------------------------------------------------------------------------------
    def crop_inputs(self, images: torch.Tensor, K: torch.Tensor, TCO: torch.Tensor, tCR: torch.Tensor,
                    labels: List[str]) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
        """pose_rigid.py:180-247 -> (images_cropped, K_crop, boxes_rend, boxes_crop)."""
        bsz = images.shape[0]
        assert K.shape == (bsz, 3, 3) and tCR.shape == (bsz, 3) and TCO.shape == (bsz, 4, 4) and len(labels) == bsz
        label_idx = self._label_idx(labels, TCO.device)
        boxes_rend, boxes_crop, K_crop = lib3d.crop_geometry(
            self.mesh_db.point_subset(2000), label_idx, TCO, K, tCR, images.shape[-2:], self.render_size)
        crops = lib3d.crop_images(self._nhwc4(images, refresh=True), boxes_crop, None, images.shape[1], self.render_size)
        return crops, K_crop, boxes_rend, boxes_crop

    def compute_crops_multiview(self, images: torch.Tensor, K: torch.Tensor, TCV_O: torch.Tensor,
                                tCR: torch.Tensor, labels: List[str]) -> torch.Tensor:
        """pose_rigid.py:249-303 -> KV_crop [bsz, n_views, 3, 3]."""
        bsz, n_views = TCV_O.shape[:2]
        label_idx = self._label_idx(labels, TCV_O.device).repeat_interleave(n_views)
        Kr = K.unsqueeze(1).repeat(1, n_views, 1, 1).flatten(0, 1)
        _, _, K_crop = lib3d.crop_geometry(self.mesh_db.point_subset(200), label_idx, TCV_O.flatten(0, 1), Kr,
                                           tCR.flatten(0, 1), images.shape[-2:], self.render_size)
        return K_crop.view(bsz, n_views, 3, 3)

    def update_pose(self, TCO, K_crop, pose_outputs, tCR) -> torch.Tensor:
        return lib3d.update_pose(TCO, K_crop, pose_outputs, tCR)

    def net_forward(self, x: torch.Tensor) -> Dict[str, torch.Tensor]:
        """pose_rigid.py:314-334; x is the concatenated [b, C, h, w] float input."""
        out = self.backbone(x)
        return {"pose": out} if self.predict_pose_update else {"renderings_logits": out}

    def render_images_multiview(self, labels: List[str], TCV_O: torch.Tensor, KV: torch.Tensor,
                                random_ambient_light: bool = False) -> torch.Tensor:
        """pose_rigid.py:336-408 -> renders [bsz, n_views*n_channels, H, W] float32."""
        if random_ambient_light:
            raise NotImplementedError("random_ambient_light is a training-time augmentation")
        bsz, n_views = TCV_O.shape[:2]
        labels_mv = [labels[n] for n in range(bsz) for _ in range(n_views)]
        if self.render_normals:  # pose_rigid.py:374-378
            lights = [[Panda3dLightData("ambient", (1.0, 1.0, 1.0, 1.0))] for _ in labels_mv]
        else:
            lights = [make_scene_lights() for _ in labels_mv]
        data = self.renderer.render(labels=labels_mv, TCO=TCV_O.flatten(0, 1), K=KV.flatten(0, 1), light_datas=lights,
                                    resolution=self.render_size, render_normals=self.render_normals,
                                    render_depth=self.render_depth, render_mask=False)
        cat = [data.rgbs] + ([data.normals] if self.render_normals else []) + ([data.depths] if self.render_depth else [])
        renders = torch.cat(cat, dim=1)
        return renders.view(bsz, n_views, renders.shape[1], *renders.shape[-2:]).flatten(1, 2)

    def normalize_depth(self, depth: torch.Tensor, tCR: torch.Tensor) -> torch.Tensor:
        """pose_rigid.py:466-496."""
        z = tCR[:, 2][(...,) + (None,) * (depth.ndim - 1)]
        kind = self.depth_normalization_type
        if kind == "tCR_scale":
            return depth / z
        if kind == "tCR_scale_clamp_center":
            return torch.clamp(depth / z, 0, 2) - 1
        if kind == "tCR_center_clamp":
            return torch.clamp(depth - z, -2, 2)
        if kind == "none":
            return depth
        raise ValueError(f"Unknown depth_normalization_type = {kind}")

    def normalize_images(self, images: torch.Tensor, renders: torch.Tensor, tCR: torch.Tensor,
                         images_inplace: bool = False, renders_inplace: bool = False):
        """pose_rigid.py:410-464."""
        if not images_inplace:
            images = images.clone()
        if not renders_inplace:
            renders = renders.clone()
        if self.input_depth:
            images[:, self._input_depth_dims] = self.normalize_depth(images[:, self._input_depth_dims], tCR)
        if self.render_depth:
            dims = (self._n_single_render_channels - 1) + self._n_single_render_channels * torch.arange(0, self.n_rendered_views)
            renders[:, dims] = self.normalize_depth(renders[:, dims], tCR)
        return images, renders

    # ------------------------------------------------------------------------------------------
    # fused step: one render-and-compare evaluation for n hypotheses
    # ------------------------------------------------------------------------------------------
    def _step(self, images: torch.Tensor, im_idx: torch.Tensor, K: torch.Tensor, label_idx: torch.Tensor,
              TCO_input: torch.Tensor, tCR: torch.Tensor, TCV_O: torch.Tensor, timing: Dict[str, float],
              cuda_timer: bool = False) -> Dict[str, torch.Tensor]:
        n = TCO_input.shape[0]
        n_views = TCV_O.shape[1]
        h, w = self.render_size
        dev = TCO_input.device
        im_size = images.shape[-2:]
        c_in = 4 if self.input_depth else 3
        boxes_rend, boxes_crop, K_crop = lib3d.crop_geometry(self.mesh_db.point_subset(2000), label_idx, TCO_input, K,
                                                             tCR, im_size, self.render_size)
        # coarse / scoring models render the single view with the crop intrinsics themselves (pose_rigid.py:684-693)
        if (n_views > 1 or self.remove_TCO_rendering) and self.predict_pose_update:
            lab_mv = label_idx.repeat_interleave(n_views)
            K_mv = K.unsqueeze(1).expand(n, n_views, 3, 3).reshape(-1, 3, 3).contiguous()
            # tOR = 0  =>  the reference point seen from each view is that view's translation
            tCV_R = TCV_O[:, :, :3, 3].reshape(-1, 3).contiguous()
            _, _, KV = lib3d.crop_geometry(self.mesh_db.point_subset(200), lab_mv, TCV_O.reshape(-1, 4, 4), K_mv, tCV_R,
                                           im_size, self.render_size)
            KV_crop = KV.view(n, n_views, 3, 3)
            if not self.remove_TCO_rendering:
                KV_crop[:, 0] = K_crop
        else:
            lab_mv = label_idx
            KV_crop = K_crop.unsqueeze(1)
        depth_z = tCR[:, 2].contiguous() if (self.input_depth or self.render_depth) and self._depth_norm_kind != 3 else None

        # persistent network input per batch size: the pad channels are zeroed once, every real channel of every
        # pixel is rewritten by the crop and raster kernels on each call (background pixels included)
        x = self._input_buffer(n, h, w)
        from . import _abi  # local import keeps the module import light

        nhwc4 = self._nhwc4(images)
        t_r = _Timer(cuda_timer)
        t_r.start()
        if getattr(self.renderer, "msaa4", False):
            # anti-aliased renders (BatchRenderer(msaa4=True)): the un-fused form of the reference -- fp32 crops and renders,
            # normalisation, concatenation (models/pose_rigid.py:546-567) -- then packed for the network
            crops = lib3d.crop_images(nhwc4, boxes_crop, im_idx, c_in, self.render_size)
            lights = [[Panda3dLightData("ambient", (1.0, 1.0, 1.0, 1.0))] if self.render_normals else make_scene_lights()]
            data = self.renderer.render(None, TCV_O.reshape(-1, 4, 4).contiguous(), KV_crop.reshape(-1, 3, 3).contiguous(),
                                        lights * lab_mv.shape[0], self.render_size, render_depth=self.render_depth,
                                        render_normals=self.render_normals, label_idx=lab_mv)
            cat = [data.rgbs] + ([data.normals] if self.render_normals else []) + ([data.depths] if self.render_depth else [])
            renders = torch.cat(cat, dim=1)
            renders = renders.view(n, n_views, renders.shape[1], h, w).flatten(1, 2)
            crops, renders = self.normalize_images(crops, renders, tCR, images_inplace=True, renders_inplace=True)
            x.copy_(self.backbone.pack_input(torch.cat((crops, renders), dim=1)))
        elif n_views == 1:
            # one kernel: the rasteriser's resolve pass also crops the observation and stores whole pixel vectors
            self.renderer.render_crop_fused(lab_mv, TCV_O.reshape(-1, 4, 4).contiguous(),
                                            KV_crop.reshape(-1, 3, 3).contiguous(), self.render_size, nhwc4, im_idx,
                                            boxes_crop, c_in, x, self.backbone.c_pad, self._n_single_render_channels,
                                            depth_z, self._raster_flags)
        else:
            _abi.check(_abi.lib().mpx_roi_align_fused(
                _abi.ptr(nhwc4), nhwc4.shape[0], nhwc4.shape[1], nhwc4.shape[2], _abi.ptr(im_idx), _abi.ptr(boxes_crop),
                n, c_in, h, w, _abi.ptr(x), self.backbone.c_pad, _abi.ptr(depth_z if self.input_depth else None),
                self._depth_norm_kind, _abi.stream_ptr()))
            self.renderer.render_fused(lab_mv, TCV_O.reshape(-1, 4, 4).contiguous(),
                                       KV_crop.reshape(-1, 3, 3).contiguous(), n_views, self.render_size, x,
                                       self.backbone.c_pad, c_in, self._n_single_render_channels,
                                       depth_z if self.render_depth else None, self._raster_flags)
        timing["render"] += t_r.stop()
        t_m = _Timer(cuda_timer)
        t_m.start()
        out = self.backbone.forward(x, h, w)
        timing["model"] += t_m.stop()
        return dict(out=out, K_crop=K_crop, KV_crop=KV_crop, boxes_rend=boxes_rend, boxes_crop=boxes_crop)

    def _materialize(self, images, im_idx, boxes_crop, labels, TCV_O, KV_crop, tCR):
        crops = lib3d.crop_images(self._nhwc4(images), boxes_crop, im_idx, 4 if self.input_depth else 3, self.render_size)
        renders = self.render_images_multiview(labels, TCV_O, KV_crop)
        return self.normalize_images(crops, renders, tCR, images_inplace=True, renders_inplace=True)

    # ------------------------------------------------------------------------------------------
    # reference API: refiner forward
    # ------------------------------------------------------------------------------------------
    @torch.no_grad()
    def forward(self, images: torch.Tensor, K: torch.Tensor, labels: List[str], TCO: torch.Tensor,
                n_iterations: int = 1, random_ambient_light: bool = False,
                batch_im_ids: Optional[torch.Tensor] = None, cuda_timer: bool = False) -> Dict[str, PosePredictorOutput]:
        """pose_rigid.py:498-604."""
        if random_ambient_light:
            raise NotImplementedError("random_ambient_light is a training-time augmentation")
        bsz = TCO.shape[0]
        assert TCO.shape == (bsz, 4, 4) and K.shape == (bsz, 3, 3) and len(labels) == bsz
        dev = TCO.device
        if batch_im_ids is None:
            assert images.shape[0] == bsz
            im_idx = torch.arange(bsz, device=dev, dtype=torch.int32)
        else:
            im_idx = batch_im_ids.to(device=dev, dtype=torch.int32).contiguous()
        K = K.float().contiguous()
        label_idx = self._label_idx(labels, dev)
        timing: Dict[str, float] = defaultdict(float)
        TCO0 = TCO.float().contiguous()
        self._nhwc4(images, refresh=True)
        eager = self.keep_images or self.debug or cuda_timer or not self.use_cuda_graphs or bsz > self.graph_max_batch
        if eager:
            iters = self._iterate(images, im_idx, K, label_idx, TCO0, n_iterations, timing, cuda_timer)
        else:
            iters = self._iterate_graphed(images, im_idx, K, label_idx, TCO0, n_iterations, timing)
        outputs: Dict[str, PosePredictorOutput] = dict()
        for n, it in enumerate(iters):
            images_crop = renders = None
            if self.keep_images or self.debug:
                images_crop, renders = self._materialize(images, im_idx, it["boxes_crop"], labels, it["TCV_O"],
                                                         it["KV_crop"], it["tCR"])
            if self.predict_pose_update:
                network_outputs = {"pose": it["out"]}
                renderings_logits = torch.empty(bsz, self.n_rendered_views, dtype=TCO0.dtype, device=dev)
            else:
                network_outputs = {"renderings_logits": it["out"]}
                renderings_logits = it["out"]
            outputs[f"iteration={n + 1}"] = PosePredictorOutput(
                renders=renders, images_crop=images_crop, TCO_input=it["TCO_input"], TCO_output=it["TCO_output"],
                TCV_O_input=it["TCV_O"], tCR=it["tCR"], labels=labels, K=K, K_crop=it["K_crop"], KV_crop=it["KV_crop"],
                network_outputs=network_outputs, boxes_rend=it["boxes_rend"], boxes_crop=it["boxes_crop"],
                renderings_logits=renderings_logits, timing_dict=timing)
        return outputs

    @torch.no_grad()
    def refine_tensors(self, images: torch.Tensor, im_idx: torch.Tensor, K: torch.Tensor, label_idx: torch.Tensor,
                       TCO: torch.Tensor, n_iterations: int) -> List[Dict[str, torch.Tensor]]:
        """`forward` on device tensors only (no label strings, no output dataclasses): used by the pipeline's
        sync-free path.  Returns one dict per iteration (TCO_input, TCO_output, K_crop, boxes_rend, boxes_crop, ...)."""
        if TCO.shape[0] == 0:
            return []
        self._nhwc4(images, refresh=True)
        timing: Dict[str, float] = defaultdict(float)
        im_idx = im_idx.to(torch.int32).contiguous()
        K, TCO = K.float().contiguous(), TCO.float().contiguous()
        label_idx = label_idx.to(torch.int32).contiguous()
        if self.use_cuda_graphs and TCO.shape[0] <= self.graph_max_batch and not (self.keep_images or self.debug):
            return self._iterate_graphed(images, im_idx, K, label_idx, TCO, n_iterations, timing)
        return self._iterate(images, im_idx, K, label_idx, TCO, n_iterations, timing)

    def _iterate(self, images, im_idx, K, label_idx, TCO_input, n_iterations, timing, cuda_timer=False):
        """The refinement loop on tensors only (pose_rigid.py:523-603); returns one dict of tensors per iteration."""
        iters = []
        for _ in range(n_iterations):
            TCO_input = lib3d.normalize_T(TCO_input)
            tCR = TCO_input[:, :3, 3].contiguous()  # tOR = 0 (pose_rigid.py:527-529)
            TCV_O = lib3d.make_TCO_multiview(TCO_input, tCR, multiview_type=self.multiview_type,
                                             n_views=self.n_rendered_views,
                                             remove_TCO_rendering=self.remove_TCO_rendering,
                                             views_inplane_rotations=self.views_inplane_rotations)
            step = self._step(images, im_idx, K, label_idx, TCO_input, tCR, TCV_O, timing, cuda_timer)
            if self.predict_pose_update:
                TCO_output = self.update_pose(TCO_input, step["K_crop"], step["out"], tCR)
            else:
                TCO_output = TCO_input.detach().clone()
            iters.append(dict(TCO_input=TCO_input, TCO_output=TCO_output, tCR=tCR, TCV_O=TCV_O, out=step["out"],
                              K_crop=step["K_crop"], KV_crop=step["KV_crop"], boxes_rend=step["boxes_rend"],
                              boxes_crop=step["boxes_crop"]))
            TCO_input = TCO_output
        return iters

    def _iterate_graphed(self, images, im_idx, K, label_idx, TCO0, n_iterations, timing):
        """Replay the whole refinement loop as one CUDA graph: with a handful of hypotheses the loop is bound by the
        ~60 launches per iteration, not by the GPU work.  One graph per (batch size, iterations, frame buffer)."""
        bsz = TCO0.shape[0]
        # the graph reads the frame from the persistent NHWC4 buffer of this shape (refreshed by forward())
        self._input_buffer(bsz, *self.render_size)  # may retire older buffers and the graphs over them (bumps the epoch)
        key = (bsz, n_iterations, tuple(images.shape), self._nhwc4(images).data_ptr(), self.graph_epoch)
        entry = self._graphs.get(key)
        if entry is None:
            # first sight: run eagerly (allocates the persistent buffers, one-time CUDA set-up); capture next time
            static = dict(im_idx=im_idx.clone(), K=K.clone(), label_idx=label_idx.clone(), TCO=TCO0.clone())
            if len(self._graphs) >= 16:
                self._graphs.clear()
            self._graphs[key] = dict(graph=None, static=static)
            return self._iterate(images, im_idx, K, label_idx, TCO0, n_iterations, timing)
        if entry["graph"] is None:
            static = entry["static"]
            for name, src in (("im_idx", im_idx), ("K", K), ("label_idx", label_idx), ("TCO", TCO0)):
                static[name].copy_(src)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            try:
                with torch.cuda.graph(graph, stream=self.graph_capture_stream):
                    iters = self._iterate(images, static["im_idx"], static["K"], static["label_idx"], static["TCO"],
                                          n_iterations, defaultdict(float))
            except Exception:  # noqa: BLE001 -- capture not possible here: stay eager for this predictor
                self.use_cuda_graphs = False
                torch.cuda.synchronize()
                return self._iterate(images, im_idx, K, label_idx, TCO0, n_iterations, timing)
            entry["graph"], entry["iters"] = graph, iters
        static = entry["static"]
        for name, src in (("im_idx", im_idx), ("K", K), ("label_idx", label_idx), ("TCO", TCO0)):
            static[name].copy_(src)
        t0 = time.time()
        entry["graph"].replay()
        timing["model"] += time.time() - t0
        return [{k: v.clone() for k, v in it.items()} for it in entry["iters"]]

    # ------------------------------------------------------------------------------------------
    # reference API: coarse / scoring forward
    # ------------------------------------------------------------------------------------------
    @torch.no_grad()
    def forward_coarse_tensor(self, x: torch.Tensor, cuda_timer: bool = False) -> Dict[str, Any]:
        """pose_rigid.py:606-632; x [B, C, H, W] already concatenated and normalised."""
        assert self.predict_rendered_views_logits, "Method only valid if coarse classification model"
        t = _Timer(cuda_timer)
        t.start()
        logits = self.net_forward(x)["renderings_logits"]
        scores = torch.sigmoid(logits)
        return {"logits": logits, "scores": scores, "time": t.stop()}

    @torch.no_grad()
    def forward_coarse(self, images: torch.Tensor, K: torch.Tensor, labels: List[str], TCO_input: torch.Tensor,
                       cuda_timer: bool = False, return_debug_data: bool = False,
                       batch_im_ids: Optional[torch.Tensor] = None,
                       label_idx: Optional[torch.Tensor] = None) -> Dict[str, Any]:
        """pose_rigid.py:634-708 -> dict(logits [B,1], scores [B,1], time, render_time, model_time)."""
        assert self.predict_rendered_views_logits, "Method only valid if coarse classification model"
        bsz = TCO_input.shape[0]
        assert TCO_input.shape == (bsz, 4, 4) and K.shape == (bsz, 3, 3) and len(labels) == bsz
        dev = TCO_input.device
        if batch_im_ids is None:
            assert images.shape[0] == bsz
            im_idx = torch.arange(bsz, device=dev, dtype=torch.int32)
        else:
            im_idx = batch_im_ids.to(device=dev, dtype=torch.int32).contiguous()
        K = K.float().contiguous()
        if label_idx is None:
            label_idx = self._label_idx(labels, dev)
        else:
            label_idx = label_idx.to(device=dev, dtype=torch.int32).contiguous()
        self._nhwc4(images, refresh=True)
        timing: Dict[str, float] = defaultdict(float)
        if (0 < bsz <= self.graph_max_batch and self.use_cuda_graphs and not (return_debug_data or cuda_timer or self.debug)
                and self.n_rendered_views == 1 and not self.predict_pose_update):
            # a handful of rows (the scoring pass): one graph replay instead of ~45 launches
            it = self._iterate_graphed(images, im_idx, K, label_idx, TCO_input.float().contiguous(), 1, timing)[0]
            logits = it["out"]
            return {"logits": logits, "scores": torch.sigmoid(logits), "time": timing["model"],
                    "render_time": timing["render"], "model_time": timing["model"]}
        TCO_n = lib3d.normalize_T(TCO_input.float())
        logits_chunks, extra = [], defaultdict(list)
        for s in range(0, bsz, self.max_batch):
            e = min(bsz, s + self.max_batch)
            T = TCO_n[s:e].contiguous()
            tCR = T[:, :3, 3].contiguous()
            step = self._step(images, im_idx[s:e].contiguous(), K[s:e].contiguous(), label_idx[s:e].contiguous(), T,
                              tCR, T.unsqueeze(1), timing, cuda_timer)
            logits_chunks.append(step["out"])
            if return_debug_data:
                crops, renders = self._materialize(images, im_idx[s:e].contiguous(), step["boxes_crop"], labels[s:e],
                                                   T.unsqueeze(1), step["KV_crop"], tCR)
                extra["images_crop"].append(crops)
                extra["renders"].append(renders)
        logits = torch.cat(logits_chunks) if logits_chunks else torch.empty(0, 1, device=dev)
        out: Dict[str, Any] = {"logits": logits, "scores": torch.sigmoid(logits), "time": timing["model"],
                               "render_time": timing["render"], "model_time": timing["model"]}
        if return_debug_data:
            out["images_crop"] = torch.cat(extra["images_crop"])
            out["renders"] = torch.cat(extra["renders"])
        return out
