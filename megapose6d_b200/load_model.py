"""Model zoo, checkpoint and config loading (reference on-disk formats kept byte-compatible).

Mirrors src/megapose/utils/load_model.py:8-89 (`NAMED_MODELS`, `load_named_model`),
src/megapose/inference/utils.py:73-148 (`load_cfg`, `load_pose_models`),
src/megapose/training/pose_models_cfg.py:36-138 (`check_update_config`, `create_model_pose`) and
src/megapose/utils/models_compat.py:17-27 (`change_keys_of_older_models`).

Layout: $MEGAPOSE_DATA_DIR/megapose-models/<run_id>/{config.yaml, checkpoint.pth.tar}, the checkpoint
being {"state_dict": OrderedDict, "epoch": int} with keys `backbone.*` (torchvision ResNet names),
`pose_fc.*` (refiner) or `views_logits_head.*` (coarse).  config.yaml is parsed with PyYAML (omegaconf is
not needed); pickled-object YAML written by old runs is read as a plain mapping.
"""
from __future__ import annotations

import os
from pathlib import Path
from typing import Any, Dict, Optional, Tuple

import torch
import yaml

from .backbone import ResNet34Engine
from .meshes import BatchedMeshes, MeshDataBase
from .object_dataset import RigidObjectDataset
from .pose_estimator import PoseEstimator
from .pose_predictor import PosePredictor
from .renderer import BatchRenderer

LOCAL_DATA_DIR = Path(os.environ.get("MEGAPOSE_DATA_DIR", Path.cwd() / "local_data"))

NAMED_MODELS = {
    "megapose-1.0-RGB": {
        "coarse_run_id": "coarse-rgb-906902141",
        "refiner_run_id": "refiner-rgb-653307694",
        "requires_depth": False,
        "inference_parameters": {"n_refiner_iterations": 5, "n_pose_hypotheses": 1},
    },
    "megapose-1.0-RGBD": {
        "coarse_run_id": "coarse-rgb-906902141",
        "refiner_run_id": "refiner-rgbd-288182519",
        "requires_depth": True,
        "inference_parameters": {"n_refiner_iterations": 5, "n_pose_hypotheses": 1},
    },
    "megapose-1.0-RGB-multi-hypothesis": {
        "coarse_run_id": "coarse-rgb-906902141",
        "refiner_run_id": "refiner-rgb-653307694",
        "requires_depth": False,
        "inference_parameters": {"n_refiner_iterations": 5, "n_pose_hypotheses": 5},
    },
    "megapose-1.0-RGB-multi-hypothesis-icp": {
        "coarse_run_id": "coarse-rgb-906902141",
        "refiner_run_id": "refiner-rgb-653307694",
        "requires_depth": True,
        "depth_refiner": "ICP",
        "inference_parameters": {"n_refiner_iterations": 5, "n_pose_hypotheses": 5, "run_depth_refiner": True},
    },
}


class Cfg(dict):
    """Mapping with attribute access; `"key" in cfg` and hasattr(cfg, "key") both work like OmegaConf."""

    def __getattr__(self, name: str) -> Any:
        try:
            return self[name]
        except KeyError as e:
            raise AttributeError(name) from e

    def __setattr__(self, name: str, value: Any) -> None:
        self[name] = value

    def __delattr__(self, name: str) -> None:
        del self[name]


class _PermissiveLoader(yaml.SafeLoader):
    pass


def _construct_python_object(loader, suffix, node):
    if isinstance(node, yaml.MappingNode):
        return loader.construct_mapping(node, deep=True)
    if isinstance(node, yaml.SequenceNode):
        return loader.construct_sequence(node, deep=True)
    return loader.construct_scalar(node)


_PermissiveLoader.add_multi_constructor("tag:yaml.org,2002:python/", _construct_python_object)


def load_cfg(path) -> Cfg:
    data = yaml.load(Path(path).read_text(), Loader=_PermissiveLoader)
    if isinstance(data, dict) and "dictitems" in data and isinstance(data["dictitems"], dict):
        data = data["dictitems"]
    assert isinstance(data, dict), f"unsupported config format: {path}"
    return Cfg(data)


def check_update_config(cfg: Cfg) -> Cfg:
    """training/pose_models_cfg.py:36-87 (back-compat rules for older runs)."""
    cfg.is_coarse_compat = False
    if cfg.get("input_strategy") == "input=obs+one_render":
        cfg.is_coarse_compat = True
        cfg.n_rendered_views = 1
        cfg.multiview_type = "1view_TCO"
        cfg.predict_rendered_views_logits = True
        cfg.remove_TCO_rendering = True
        cfg.predict_pose_update = False
    renames = {"front_3views": "TCO+front_3views", "front_5views": "TCO+front_5views", "front_1view": "TCO+front_1view"}
    if cfg.get("multiview_type") in renames:
        cfg.multiview_type = renames[cfg.multiview_type]
    cfg.setdefault("predict_pose_update", True)
    cfg.setdefault("remove_TCO_rendering", False)
    cfg.setdefault("predict_rendered_views_logits", False)
    if "n_rendered_views" not in cfg:
        cfg.n_rendered_views = cfg.pop("n_views") if "n_views" in cfg else 1
    cfg.setdefault("render_normals", False)
    cfg.setdefault("render_depth", False)
    cfg.setdefault("input_depth", False)
    if "multiview_type" not in cfg:
        cfg.multiview_type = "TCO"
        assert not cfg.remove_TCO_rendering
    cfg.views_inplane_rotations = cfg.get("views_inplane_rotations", False)
    if "depth_augmentation" not in cfg:  # pose_models_cfg.py:81-82
        cfg.depth_normalization_type = "tCR_scale"
    cfg.setdefault("depth_normalization_type", "tCR_scale")
    cfg.setdefault("renderer", "panda3d")
    cfg.setdefault("backbone_str", "vanilla_resnet34")
    return cfg


def change_keys_of_older_models(state_dict: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """utils/models_compat.py:17-27."""
    out = dict()
    for k, v in state_dict.items():
        if k.startswith("backbone.backbone"):
            k = "backbone." + k[len("backbone.backbone."):]
        elif k.startswith("backbone.head.0."):
            k = "views_logits_head." + k[len("backbone.head.0."):]
        out[k] = v
    return out


def n_input_channels(cfg: Cfg) -> int:
    """training/pose_models_cfg.py:95-103."""
    return (3 + (1 if cfg.input_depth else 0)) + (3 + (3 if cfg.render_normals else 0) + (1 if cfg.render_depth else 0)) * cfg.n_rendered_views


def create_model_pose(cfg: Cfg, renderer: BatchRenderer, mesh_db: BatchedMeshes,
                      state_dict: Dict[str, torch.Tensor], render_size: Tuple[int, int] = (240, 320)) -> PosePredictor:
    """training/pose_models_cfg.py:90-138 + load_state_dict; builds the engine from the checkpoint tensors."""
    if cfg.backbone_str not in ("vanilla_resnet34", "resnet34", "resnet18", "resnet34_width=1"):
        # training/pose_models_cfg.py:106-118; wider WideResNets (resnet34_width=k, k > 1) have up to 2048 channels
        raise NotImplementedError(f"backbone '{cfg.backbone_str}': vanilla_resnet34 (all released models), resnet34 and "
                                  "resnet18 (WideResNet, width 1) are implemented")
    head = "pose_fc" if cfg.predict_pose_update else "views_logits_head"
    expected = {head + ".weight", head + ".bias", "backbone.conv1.weight"}
    expected.add("backbone.fc.weight" if cfg.backbone_str == "vanilla_resnet34" else "backbone.layer1.0.bn1.weight")
    missing = expected - set(state_dict.keys())
    if missing:
        raise RuntimeError(f"checkpoint is missing keys {sorted(missing)}")
    backbone = ResNet34Engine(state_dict, n_inputs=n_input_channels(cfg), head=head)
    model = PosePredictor(
        backbone=backbone, renderer=renderer, mesh_db=mesh_db, render_size=tuple(render_size),
        n_rendered_views=cfg.n_rendered_views, views_inplane_rotations=cfg.views_inplane_rotations,
        multiview_type=cfg.multiview_type, render_normals=cfg.render_normals, render_depth=cfg.render_depth,
        input_depth=cfg.input_depth, predict_rendered_views_logits=cfg.predict_rendered_views_logits,
        remove_TCO_rendering=cfg.remove_TCO_rendering, predict_pose_update=cfg.predict_pose_update,
        depth_normalization_type=cfg.depth_normalization_type)
    return model


def load_pose_models(coarse_run_id: str, refiner_run_id: str, object_dataset: RigidObjectDataset,
                     force_panda3d_renderer: bool = False, renderer_kwargs: Optional[dict] = None,
                     models_root: Optional[Path] = None,
                     render_size: Tuple[int, int] = (240, 320)) -> Tuple[PosePredictor, PosePredictor, MeshDataBase]:
    """inference/utils.py:80-148 -> (coarse_model, refiner_model, mesh_db).  `render_size` is the crop / render
    resolution of both models: the reference fixes it at (240, 320) in training/pose_models_cfg.py:105 and passes it
    to PosePredictor(render_size=...) (models/pose_rigid.py:87); any even size works here."""
    models_root = Path(models_root) if models_root is not None else LOCAL_DATA_DIR / "megapose-models"
    mesh_db = MeshDataBase.from_object_ds(object_dataset)
    mesh_db_batched = mesh_db.batched().cuda()
    kwargs = dict(renderer_kwargs or {})
    kwargs.pop("split_objects", None)
    kwargs.pop("preload_cache", None)
    kwargs.pop("n_workers", None)
    renderer = BatchRenderer(object_dataset=object_dataset, mesh_db=mesh_db_batched, **kwargs)

    def load_model(run_id: Optional[str]) -> Optional[PosePredictor]:
        if run_id is None:
            return None
        run_dir = models_root / run_id
        cfg = check_update_config(load_cfg(run_dir / "config.yaml"))
        ckpt = torch.load(run_dir / "checkpoint.pth.tar", map_location="cpu", weights_only=False)
        state_dict = change_keys_of_older_models(ckpt["state_dict"])
        model = create_model_pose(cfg, renderer=renderer, mesh_db=mesh_db_batched, state_dict=state_dict,
                                  render_size=render_size)
        model = model.eval()
        model.cfg = cfg
        model.config = cfg
        return model

    return load_model(coarse_run_id), load_model(refiner_run_id), mesh_db


def load_named_model(model_name: str, object_dataset: RigidObjectDataset, n_workers: int = 4,
                     bsz_images: int = 128, models_root: Optional[Path] = None,
                     render_size: Tuple[int, int] = (240, 320)) -> PoseEstimator:
    """utils/load_model.py:50-89."""
    model = NAMED_MODELS[model_name]
    coarse_model, refiner_model, mesh_db = load_pose_models(
        coarse_run_id=model["coarse_run_id"], refiner_run_id=model["refiner_run_id"], object_dataset=object_dataset,
        force_panda3d_renderer=True, renderer_kwargs={"preload_cache": False, "split_objects": False,
                                                      "n_workers": n_workers},
        models_root=models_root, render_size=render_size)
    depth_refiner = None
    if model.get("depth_refiner") == "ICP":  # utils/load_model.py:75-79
        from .icp_refiner import ICPRefiner

        depth_refiner = ICPRefiner(refiner_model.mesh_db, refiner_model.renderer)
    return PoseEstimator(refiner_model=refiner_model, coarse_model=coarse_model, detector_model=None,
                         depth_refiner=depth_refiner, bsz_objects=8, bsz_images=bsz_images)


# ---------------------------------------------------------------------------------------------
# synthetic zoo (no checkpoint can be downloaded offline): seeded random weights in the zoo layout
# ---------------------------------------------------------------------------------------------
ZOO_CONFIGS = {
    "coarse-rgb-906902141": dict(backbone_str="vanilla_resnet34", n_rendered_views=1, multiview_type="TCO",
                                 render_normals=True, render_depth=False, input_depth=False,
                                 predict_rendered_views_logits=True, predict_pose_update=False,
                                 remove_TCO_rendering=False, depth_normalization_type="tCR_scale_clamp_center",
                                 depth_augmentation=False, renderer="panda3d"),
    "refiner-rgb-653307694": dict(backbone_str="vanilla_resnet34", n_rendered_views=4, multiview_type="TCO+front_3views",
                                  render_normals=True, render_depth=False, input_depth=False,
                                  predict_rendered_views_logits=False, predict_pose_update=True,
                                  remove_TCO_rendering=False, depth_normalization_type="tCR_scale_clamp_center",
                                  depth_augmentation=False, renderer="panda3d"),
    "refiner-rgbd-288182519": dict(backbone_str="vanilla_resnet34", n_rendered_views=4, multiview_type="TCO+front_3views",
                                   render_normals=True, render_depth=True, input_depth=True,
                                   predict_rendered_views_logits=False, predict_pose_update=True,
                                   remove_TCO_rendering=False, depth_normalization_type="tCR_scale_clamp_center",
                                   depth_augmentation=True, renderer="panda3d"),
}


def write_run(models_root: Path, run_id: str, state_dict: Dict[str, torch.Tensor], cfg: Optional[dict] = None) -> Path:
    """Write <models_root>/<run_id>/{config.yaml, checkpoint.pth.tar} in the reference format
    (training/utils.py:156-172)."""
    run_dir = Path(models_root) / run_id
    run_dir.mkdir(parents=True, exist_ok=True)
    cfg = dict(cfg if cfg is not None else ZOO_CONFIGS[run_id])
    (run_dir / "config.yaml").write_text(yaml.safe_dump(cfg))
    torch.save({"state_dict": state_dict, "epoch": 0}, run_dir / "checkpoint.pth.tar")
    return run_dir
