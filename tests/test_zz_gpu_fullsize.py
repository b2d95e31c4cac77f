"""BASELINE configs[1..3] at FULL size on the GPU against what the REAL reference returned for the same inputs.

tests/golden/fullsize_*.npz were written by tools/make_golden.py: the reference's own
PoseEstimator.run_inference_pipeline (inference/pose_estimator.py:511-641; fp32 on the host, the C rasteriser of
oracle/raster_ref.c standing in for Panda3D) on the scenarios of workloads/scenes.py:

  fullsize_rgb      configs[1]: 1 object x 576 hypotheses + 5 refiner iterations + scoring, 240x320 (what bench.py times)
  fullsize_rgb_224  the same at 224x224 crops / renders (the size BASELINE.json's metric names)
  fullsize_rgbd32   configs[2]: RGB-D refiner, 32-object batch x 5 refiner iterations (72-rotation coarse grid)
  fullsize_ycbv21   configs[3]: 21 objects x 576 hypotheses in one frame (12 096 coarse rows)

Asserted, with the fp16-vs-fp32 tolerances stated here:
  * coarse logits: common offset within LOGIT_OFFSET_STD standard deviations of the reference's logits, every logit within
    LOGIT_TOL_STD (max) and LOGIT_RMS_STD (rms) of the reference's about that offset;
  * per detection the same surviving hypothesis as the reference -- or, where the reference's own margin between its
    best candidates is inside twice the observed noise, one of those near-tied candidates (counted and bounded);
  * the refiner, iteration by iteration, each started from the REFERENCE's input pose of that iteration: output pose within
    ROT_TOL_DEG / TRANS_TOL_MM of the reference's (0.5 deg, 1 mm).  Free-running, the five iterations of a
    random-weight refiner are not a contraction (a trained one converges): a 1e-4 difference after iteration 1 grows by
    3-5x per iteration, and in the RGB-D scenario the 0.99 validity threshold of the noisy depth crop flips pixels on top
    of that -- the free-running poses are therefore printed and bounded loosely (FREE_ROT_TOL_DEG / FREE_TRANS_TOL_MM on the
    RGB scenarios), the per-iteration comparison is the parity statement;
  * scoring logits of the detections with the reference's survivor within twice the logit tolerance on RGB.
Plus the size-independent properties of round 1 (idempotence over eager / capture / replay, host-resident inputs, rigid
poses, score = sigmoid(logit)).  The file sorts last on purpose: it allocates the full-size buffers and graphs."""
from pathlib import Path

import numpy as np
import pytest
import torch

from workloads import scenes

pytestmark = pytest.mark.gpu
GOLDEN = Path(__file__).resolve().parent / "golden"

LOGIT_OFFSET_STD = 0.10  # |mean logit error| / std of the reference's logits: a common offset (observed -0.06 .. -0.07 std
                         # on the RGB scenarios) moves no ranking and shifts every sigmoid score alike
LOGIT_TOL_STD = 0.08     # max |logit error - mean error| / std (observed with fp16: 0.03-0.06 over 576 .. 12096 rows)
LOGIT_RMS_STD = 0.03     # rms of the error about its mean (observed 0.010-0.015)
ROT_TOL_DEG = 0.5      # SURVEY 8c (iv), per refiner iteration from the reference's input pose
TRANS_TOL_MM = 1.0
FREE_ROT_TOL_DEG = 3.0   # free-running 5 iterations, RGB scenarios (see the docstring; observed 0.03-0.6 deg, 0.2-6 mm,
FREE_TRANS_TOL_MM = 15.0  # depending on which kernels round the first iteration)


def _run(est, sc, pinned=False):
    from megapose6d_b200.tensor_collection import PandasTensorCollection
    from megapose6d_b200.types import ObservationTensor

    if pinned:
        images, K, bboxes = (t.pin_memory().cuda(non_blocking=True) for t in (sc["images"], sc["K"], sc["bboxes"]))
    else:
        images, K, bboxes = sc["images"].cuda(), sc["K"].cuda(), sc["bboxes"].cuda()
    det = PandasTensorCollection(sc["det_df"].copy(), bboxes=bboxes)
    final, extra = est.run_inference_pipeline(ObservationTensor(images, K), detections=det,
                                              n_refiner_iterations=sc["n_refiner_iterations"],
                                              n_pose_hypotheses=sc["n_pose_hypotheses"])
    torch.cuda.synchronize()
    return final, extra


def _pose_err(a: torch.Tensor, b: torch.Tensor):
    """geodesic rotation error [deg] and translation error [mm] between batches of 4x4 poses."""
    a, b = a.double().cpu(), b.double().cpu()
    dR = a[:, :3, :3].transpose(1, 2) @ b[:, :3, :3]
    cos = ((dR.diagonal(dim1=1, dim2=2).sum(-1) - 1) / 2).clamp(-1, 1)
    return torch.rad2deg(torch.acos(cos)), (a[:, :3, 3] - b[:, :3, 3]).norm(dim=-1) * 1000.0


def _rigid(poses):
    R = poses[:, :3, :3].double().cpu()
    eye = torch.eye(3, dtype=torch.float64).expand_as(R)
    assert torch.allclose(R.transpose(1, 2) @ R, eye, atol=1e-4), "rotation block is not orthonormal"
    assert torch.allclose(torch.linalg.det(R), torch.ones(len(R), dtype=torch.float64), atol=1e-4)
    assert (poses[:, 2, 3] > 0).all() and torch.isfinite(poses).all()


@pytest.mark.parametrize("name", list(scenes.FULLSIZE))
def test_full_size_pipeline_matches_the_reference(name):
    path = GOLDEN / f"{name}.npz"
    if not path.exists():
        pytest.skip(f"{path.name} has not been generated (tools/make_golden.py {name})")
    g = np.load(path)
    sc = scenes.FULLSIZE[name]()
    est = scenes.build_estimator(sc)
    B, M = len(sc["labels"]), sc["grid"]
    runs = [_run(est, sc) for _ in range(3)]  # eager, graph capture, graph replay
    final, extra = runs[-1]

    # ---- size-independent properties
    for f, _ in runs[:-1]:
        assert torch.equal(f.poses, final.poses) and f.infos["pose_logit"].tolist() == final.infos["pose_logit"].tolist()
    fp, _ = _run(est, sc, pinned=True)
    assert torch.equal(fp.poses, final.poses)
    coarse = extra["coarse"]["preds"]
    assert len(coarse) == B * M and coarse.infos["hypothesis_id"].tolist() == list(range(M)) * B
    _rigid(coarse.poses), _rigid(final.poses)
    logits = coarse.infos["coarse_logit"].to_numpy().astype(np.float64)
    assert np.allclose(coarse.infos["coarse_score"].to_numpy(), 1.0 / (1.0 + np.exp(-logits)), atol=1e-6)

    # ---- coarse logits against the reference's
    want = g["coarse_logit"].astype(np.float64)
    assert np.array_equal(g["coarse_hypothesis"], coarse.infos["hypothesis_id"].to_numpy())
    err = np.abs(logits - want)
    std = want.std()
    d = logits - want
    rms_c = np.sqrt(((d - d.mean()) ** 2).mean())
    print(f"[{name}] coarse logits: max err {err.max():.4f} = {err.max() / std:.3f} std, mean offset {d.mean():+.4f}, rms about "
          f"the mean {rms_c:.4f} = {rms_c / std:.3f} std (reference std {std:.3f}, {B * M} rows)")
    max_c = np.abs(d - d.mean()).max()
    print(f"[{name}] coarse logits about the offset: max {max_c:.4f} = {max_c / std:.3f} std")
    assert abs(d.mean()) <= LOGIT_OFFSET_STD * std and max_c <= LOGIT_TOL_STD * std and rms_c <= LOGIT_RMS_STD * std

    # ---- survivors
    kept = extra["coarse_filter"]["preds"].infos
    noise = 2.0 * max_c  # ranking noise: the error about the common offset
    same, near_tie = [], 0
    for det in range(B):
        w = want[det * M:(det + 1) * M]
        got_h = int(kept[kept["bbox_id"] == det]["hypothesis_id"].iloc[0])
        want_h = int(g["kept_hypothesis"][g["kept_bbox_id"] == det][0])
        if got_h == want_h:
            same.append(det)
        else:
            assert w[want_h] - w[got_h] <= noise, (det, got_h, want_h, w[want_h] - w[got_h], noise)
            near_tie += 1
    print(f"[{name}] survivors: {len(same)}/{B} identical to the reference's, {near_tie} near-ties inside {noise:.3f}")
    assert near_tie <= max(1, B // 8)

    # ---- refiner, one iteration at a time from the reference's own input pose of that iteration
    n_it = sc["n_refiner_iterations"]
    order = [int(np.flatnonzero(g["kept_bbox_id"] == det)[0]) for det in range(B)]  # fixture rows in detection order
    images_dev, K_dev = sc["images"].cuda(), sc["K"].cuda()
    K_rows = K_dev.expand(B, 3, 3).contiguous()
    im_ids = torch.zeros(B, dtype=torch.long)
    worst = (0.0, 0.0)
    for it in range(n_it):
        pose_in = torch.from_numpy(g["kept_poses"][order] if it == 0 else g["refiner_poses"][it - 1][order]).cuda()
        out = est.refiner_model(images=images_dev, K=K_rows, labels=sc["labels"], TCO=pose_in, n_iterations=1,
                                batch_im_ids=im_ids)["iteration=1"].TCO_output
        rot, tr = _pose_err(out, torch.from_numpy(g["refiner_poses"][it][order]))
        print(f"[{name}] refiner iteration {it + 1} from the reference's input: max rotation error {rot.max():.4f} deg, "
              f"max translation error {tr.max():.4f} mm")
        worst = (max(worst[0], rot.max().item()), max(worst[1], tr.max().item()))
    assert worst[0] <= ROT_TOL_DEG and worst[1] <= TRANS_TOL_MM, worst

    # ---- free-running pipeline: iterations, scoring logit and final pose of the detections with the reference's survivor
    preds = extra["refiner_all_hypotheses"]["preds"]
    rows_g = [order[det] for det in same]
    rows_o = [int(np.flatnonzero(kept["bbox_id"].to_numpy() == det)[0]) for det in same]
    free = (0.0, 0.0)
    for it in range(n_it):
        p = preds[f"iteration={it + 1}"].poses[rows_o]
        rot, tr = _pose_err(p, torch.from_numpy(g["refiner_poses"][it][rows_g]))
        print(f"[{name}] free-running iteration {it + 1}: max rotation error {rot.max():.4f} deg, max translation error "
              f"{tr.max():.4f} mm (median {rot.median():.4f} deg, {tr.median():.4f} mm)")
        free = (max(free[0], rot.max().item()), max(free[1], tr.max().item()))
    rgb_only = not sc["cfg_refiner"]["input_depth"]
    if rgb_only:
        assert free[0] <= FREE_ROT_TOL_DEG and free[1] <= FREE_TRANS_TOL_MM, free
    scored = extra["scoring"]["preds"].infos
    sl = scored["pose_logit"].to_numpy().astype(np.float64)[rows_o]
    serr = np.abs(sl - g["scored_pose_logit"].astype(np.float64)[rows_g])
    print(f"[{name}] scoring logits (free-running poses): max err {serr.max():.4f}, median {np.median(serr):.4f}")
    if rgb_only:  # the refined poses differ on top of the network's own error; single worst cases follow the pose divergence
        assert np.median(serr) <= (LOGIT_OFFSET_STD + 2 * LOGIT_TOL_STD) * std
    labels_final = final.infos["label"].tolist()
    assert sorted(labels_final) == sorted(g["final_label"].tolist())
    for det in same:
        i_o = labels_final.index(sc["labels"][det])           # the scenarios use one detection per label
        i_g = g["final_label"].tolist().index(sc["labels"][det])
        assert int(final.infos["hypothesis_id"].iloc[i_o]) == int(g["final_hypothesis"][i_g])
        if rgb_only:
            rot, tr = _pose_err(final.poses[i_o:i_o + 1], torch.from_numpy(g["final_poses"][i_g:i_g + 1]))
            assert rot.item() <= FREE_ROT_TOL_DEG and tr.item() <= FREE_TRANS_TOL_MM, (det, rot.item(), tr.item())
