"""Host-side containers and helpers of the product against the REAL reference classes (loaded by path from
/root/reference, marker `reference`: runs only where the reference tree exists)."""
import pickle

import numpy as np
import pandas as pd
import pytest
import torch

from megapose6d_b200 import pose_estimator as mine_pe
from megapose6d_b200 import tensor_collection as mine_tc
from megapose6d_b200.types import ObservationTensor as MineObs
from oracle import refload

pytestmark = pytest.mark.reference


@pytest.fixture(scope="module")
def ref():
    return refload.load()


def _frame(n, seed):
    rs = np.random.RandomState(seed)
    return pd.DataFrame(dict(label=[f"obj_{k}" for k in rs.randint(0, 3, n)], batch_im_id=rs.randint(0, 2, n),
                             score=rs.rand(n)), index=rs.permutation(n) + 10)


def _same(a, b):
    assert len(a) == len(b)
    pd.testing.assert_frame_equal(a.infos, b.infos)
    assert list(a.tensors.keys()) == list(b.tensors.keys())
    for k in a.tensors:
        assert torch.equal(a.tensors[k], b.tensors[k]), k


def test_pandas_tensor_collection_behaves_like_the_reference(ref):
    R = ref.tensor_collection
    n = 9
    g = torch.Generator().manual_seed(0)
    poses, boxes = torch.randn(n, 4, 4, generator=g), torch.randn(n, 4, generator=g)
    a = R.PandasTensorCollection(infos=_frame(n, 1), poses=poses.clone(), bboxes=boxes.clone())
    b = mine_tc.PandasTensorCollection(infos=_frame(n, 1), poses=poses.clone(), bboxes=boxes.clone())
    _same(a, b)                                           # index reset to 0..n-1 on construction
    assert torch.equal(a.poses, b.poses) and a.device == b.device
    for ids in ([2, 0, 5], torch.tensor([1, 1, 8]), np.array([3, 4]), [0]):
        _same(a[ids], b[ids])
    extra = pd.DataFrame(dict(label=["obj_0", "obj_1", "obj_2"], diameter=[0.1, 0.2, 0.3]))
    _same(a.merge_df(extra, on="label"), b.merge_df(extra, on="label"))
    _same(R.concatenate([a[[0, 1]], a[[5]]]), mine_tc.concatenate([b[[0, 1]], b[[5]]]))
    _same(a.clone(), b.clone())
    a.register_tensor("extra", torch.arange(n))
    b.register_tensor("extra", torch.arange(n))
    a.poses = a.poses * 2
    b.poses = b.poses * 2
    _same(a, b)
    a.delete_tensor("extra")
    b.delete_tensor("extra")
    _same(pickle.loads(pickle.dumps(a)), pickle.loads(pickle.dumps(b)))
    _same(a.double(), b.double())
    with pytest.raises(AttributeError):
        _ = b.not_a_tensor
    assert len(mine_tc.concatenate([])) == len(R.concatenate([])) == 0


def test_observation_tensor_behaves_like_the_reference(ref):
    RObs = ref.types.ObservationTensor
    rs = np.random.RandomState(3)
    rgb = rs.randint(0, 255, size=(12, 16, 3), dtype=np.uint8)
    depth = rs.rand(12, 16).astype(np.float32)
    K = np.array([[500.0, 0, 8], [0, 510, 6], [0, 0, 1]])
    for d in (None, depth):
        a, b = RObs.from_numpy(rgb, d, K), MineObs.from_numpy(rgb, d, K)
        assert torch.equal(a.images, b.images) and torch.equal(a.K, b.K)
        assert a.batch_size == b.batch_size and a.channel_dim == b.channel_dim and a.is_valid() == b.is_valid()
        if d is not None:
            assert torch.equal(a.depth, b.depth)
    rgb_b = torch.from_numpy(rs.randint(0, 255, size=(2, 3, 12, 16), dtype=np.uint8))
    dep_b = torch.from_numpy(rs.rand(2, 1, 12, 16).astype(np.float32))
    Kb = torch.from_numpy(np.stack([K, K])).float()
    a, b = RObs.from_torch_batched(rgb_b, dep_b, Kb), MineObs.from_torch_batched(rgb_b, dep_b, Kb)
    assert torch.equal(a.images, b.images) and torch.equal(a.K, b.K) and a.channel_dim == b.channel_dim == 4
    # [B,H,W] depth: the reference drops the result of `depth.unsqueeze(1)` (inference/types.py:224-229) and raises; the
    # product accepts it as documented there
    with pytest.raises(RuntimeError):
        RObs.from_torch_batched(rgb_b, dep_b[:, 0], Kb)
    assert torch.equal(MineObs.from_torch_batched(rgb_b, dep_b[:, 0], Kb).images, b.images)


def test_instance_ids_and_detection_filter_match_the_reference(ref):
    R, U = ref.tensor_collection, ref.inference_utils
    for seed in range(4):
        n = 11
        boxes = torch.randn(n, 4, generator=torch.Generator().manual_seed(seed))
        a = U.add_instance_id(R.PandasTensorCollection(infos=_frame(n, seed), bboxes=boxes.clone()))
        b = mine_pe.add_instance_id(mine_tc.PandasTensorCollection(infos=_frame(n, seed), bboxes=boxes.clone()))
        assert a.infos["instance_id"].tolist() == b.infos["instance_id"].tolist()
        # an existing column is left alone by both
        assert mine_pe.add_instance_id(b).infos["instance_id"].tolist() == a.infos["instance_id"].tolist()
        # (with this pandas the reference's groupby().apply() drops the grouping columns from `a.infos`; the product keeps
        # them, as the reference did with the pandas it was written for) -> give the reference filter the complete frame
        a = R.PandasTensorCollection(infos=b.infos.copy(), bboxes=boxes.clone())
        for kwargs in (dict(labels=["obj_0", "obj_2"]), dict(one_instance_per_class=True),
                       dict(labels=["obj_1"], one_instance_per_class=True)):
            fa, fb = U.filter_detections(a, **kwargs), mine_pe.filter_detections(b, **kwargs)
            assert fa.infos["label"].tolist() == fb.infos["label"].tolist()
            assert fa.infos["score"].tolist() == fb.infos["score"].tolist()
            assert torch.equal(fa.bboxes, fb.bboxes)


class _AttrDict(dict):
    """OmegaConf-like config for the reference function: attribute access, `in`, `del`, hasattr."""
    __getattr__ = lambda self, k: self[k] if k in self else (_ for _ in ()).throw(AttributeError(k))  # noqa: E731
    __setattr__ = dict.__setitem__
    __delattr__ = dict.__delitem__


def test_check_update_config_matches_the_reference():
    """training/pose_models_cfg.py:36-87 executed from the reference source (its module imports Panda3D, so only the
    function body is compiled) against load_model.check_update_config on legacy and current run configurations."""
    import logging
    import re

    from megapose6d_b200 import load_model

    src = (refload.REF_ROOT / "training/pose_models_cfg.py").read_text()
    body = src[src.index("def check_update_config"):src.index("def create_model_pose")]
    body = re.sub(r"\(cfg: TrainingConfig\) -> TrainingConfig", "(cfg)", body)
    ns = dict(logger=logging.getLogger("ref"))
    exec(compile(body, "pose_models_cfg.py", "exec"), ns)
    cases = [
        dict(load_model.ZOO_CONFIGS["coarse-rgb-906902141"]),
        dict(load_model.ZOO_CONFIGS["refiner-rgbd-288182519"]),
        dict(input_strategy="input=obs+one_render", multiview_type="TCO", backbone_str="vanilla_resnet34"),
        dict(multiview_type="front_3views", n_views=4, backbone_str="vanilla_resnet34"),
        dict(multiview_type="front_5views", n_rendered_views=6, render_normals=True, depth_augmentation=True,
             depth_normalization_type="tCR_scale_clamp_center"),
    ]
    for case in cases:
        want = ns["check_update_config"](_AttrDict(case))
        got = load_model.check_update_config(load_model.Cfg(dict(case)))
        for key in ("is_coarse_compat", "n_rendered_views", "multiview_type", "predict_rendered_views_logits",
                    "remove_TCO_rendering", "predict_pose_update", "render_normals", "render_depth", "input_depth",
                    "renderer"):
            assert got[key] == want[key], (case, key, got[key], want[key])
        if "depth_normalization_type" in want:
            assert got["depth_normalization_type"] == want["depth_normalization_type"], case
        assert "n_views" not in got
