#!/usr/bin/env python
"""bench.py -- pose hypotheses / second through render + coarse + 5x refine (BASELINE.json metric).

    python bench.py --gpus 1 --steps 5 --warmup 3            # this framework (libmpx.so, sm_100a)
    python bench.py --impl reference --steps 2 --warmup 1    # CPU oracle port of the reference path
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (config.workload): BASELINE.json configs[1] "megapose-1.0-RGB: 1 object x 576 coarse hypotheses + 5
refiner iters" at the reference's 240x320 render size, synthetic 480x640 frame, procedural 10k-triangle mesh,
seeded random vanilla_resnet34 weights in the model-zoo checkpoint format.  One step = one call of
PoseEstimator.run_inference_pipeline for a frame with one detection per rank (weak scaling: N ranks score a
frame with N detections, rows sharded by detection, one all-gather per stage).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

M_GRID = 576
GFLOP_PER_HYP = 12.213  # coarse 12.068 + (5 * 14.236 + 12.068) / 576, 240x320, FLOP = 2*MAC
GFLOP_COARSE_PER_HYP = 12.068
# bf16 bytes the 36 convolutions of one coarse forward move at least once per hypothesis (each conv: input read +
# output write + residual read; weights excluded): stem 2 x 2.46 MB, layer1 6 convs of 60x80x64, ... = 28.26 MB
ALGO_CONV_BYTES_PER_HYP = 28.26e6
N_REFINER_ITERS = 5


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="mpx", choices=["mpx", "reference"])
    ap.add_argument("--cpu-sample", type=int, default=32, help="coarse hypotheses in the bounded CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------
# synthetic scene shared by both arms
# ---------------------------------------------------------------------------------------------
def build_scene(n_objects: int):
    import pandas as pd
    import torch

    from megapose6d_b200 import procedural
    from tests import helpers

    ds, images, K = helpers.make_scene(n_objects, seed=0)
    labels = [o.label for o in ds.list_objects]
    poses = torch.from_numpy(procedural.random_poses(n_objects, 5, z_range=(0.5, 0.9), xy_range=0.1)).float()
    bboxes = torch.stack([helpers.detection_for_pose(K[0], poses[i], torch.from_numpy(ds[i].mesh.vertices).float())
                          for i in range(n_objects)])
    det_df = pd.DataFrame(dict(label=labels, batch_im_id=0, instance_id=np.arange(n_objects)))
    sds = {"coarse-rgb-906902141": helpers.make_state_dict(helpers.COARSE_CFG, 1),
           "refiner-rgb-653307694": helpers.make_state_dict(helpers.REFINER_CFG, 2)}
    return ds, images, K, det_df, bboxes, sds


# ---------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference path on the host cores
# ---------------------------------------------------------------------------------------------
def cpu_sample_rate(scene, n_coarse: int):
    """Bounded sample of the workload on the host: n_coarse coarse hypotheses, then 5 refiner iterations and the
    scoring pass on the best of them; extrapolated to hypotheses/s of the 576-hypothesis unit."""
    import torch

    from oracle import pipeline_ref
    from tests import helpers

    ds, images, K, det_df, bboxes, sds = scene
    # torch's CPU convolutions stop scaling (and regress) past a few dozen threads at these batch sizes
    cores = min(os.cpu_count() or 1, 32)
    torch.set_num_threads(cores)
    meshes = helpers.ref_meshes_from_dataset(ds)
    rr = pipeline_ref.RefRenderer(meshes, n_threads=cores)
    oc = pipeline_ref.RefPosePredictor(sds["coarse-rgb-906902141"], helpers.COARSE_CFG, meshes, rr)
    orf = pipeline_ref.RefPosePredictor(sds["refiner-rgb-653307694"], helpers.REFINER_CFG, meshes, rr)
    est = pipeline_ref.RefPoseEstimator(oc, orf, bsz_images=n_coarse, bsz_objects=8, SO3_grid_size=M_GRID)
    df1 = det_df.iloc[:1].copy()
    with torch.no_grad():
        est.forward_coarse_model(images, K, df1, bboxes[:1], max_hypotheses=2)  # untimed: one-time op loading
        t0 = time.time()
        df_c, TCO_c = est.forward_coarse_model(images, K, df1, bboxes[:1], max_hypotheses=n_coarse)
        t_coarse = time.time() - t0
        keep = est.filter_pose_estimates(df_c, 1, "coarse_logit")
        t0 = time.time()
        ref = est.forward_refiner(images, K, df_c.iloc[keep].reset_index(drop=True), TCO_c[keep], N_REFINER_ITERS)
        est.forward_scoring_model(images, K, df_c.iloc[keep].reset_index(drop=True), ref[f"iteration={N_REFINER_ITERS}"]["TCO_output"])
        t_refine = time.time() - t0
    unit_time = M_GRID * (t_coarse / n_coarse) + t_refine
    sample = (f"{n_coarse} of 576 coarse hypotheses ({t_coarse:.2f} s) + 5 refiner iterations and scoring of the best one "
              f"({t_refine:.2f} s), extrapolated to the 576-hypothesis unit; torch fp32 + C rasteriser on {cores} threads")
    return M_GRID / unit_time, cores, sample, t_coarse + t_refine


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    scene = build_scene(1)
    for _ in range(max(0, args.warmup)):
        cpu_sample_rate(scene, max(4, args.cpu_sample // 4))
    vals, secs = [], []
    for _ in range(max(1, args.steps)):
        v, cores, sample, s = cpu_sample_rate(scene, args.cpu_sample)
        vals.append(v)
        secs.append(s)
    value = float(np.mean(vals))
    line = {
        "metric": "pose hypotheses/sec through render+coarse+5x refine", "value": value, "unit": "hypotheses/s",
        "impl": "reference", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1000.0 * float(np.mean(secs)), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": workload_config(1),
        "cpu_baseline": {"value": value, "unit": "hypotheses/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "hypotheses/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def workload_config(n_gpus: int):
    return {"workload": "megapose-1.0-RGB: 1 object x 576 coarse hypotheses + 5 refiner iters + scoring per GPU "
                        "(BASELINE configs[1]); 480x640 frame, 240x320 crops/renders, 10k-triangle procedural mesh, "
                        "vanilla_resnet34 random weights",
            "hypotheses_per_step": M_GRID * n_gpus, "parallelism": f"hypothesis-sharded x{n_gpus}",
            "l2": "network input tensor (1.4 GB per 576 hypotheses) >> 126 MB L2, no explicit flush"}


# ---------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi during the timed region)
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for l in self.lines:
            f = [t.strip() for t in l.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for name, val in zip(names, f[2:6]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------
def run_mpx_arm(args):
    import ctypes

    import torch
    import torch.distributed as dist

    from megapose6d_b200 import _abi, load_model
    from megapose6d_b200.parallel import HypothesisSharder
    from megapose6d_b200.tensor_collection import PandasTensorCollection
    from megapose6d_b200.types import ObservationTensor

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py --impl mpx needs a CUDA device (there is no CPU fallback)"
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    n_gpus = world
    scene = build_scene(n_gpus)
    ds, images, K, det_df, bboxes, sds = scene
    with tempfile.TemporaryDirectory() as tmp:
        for run_id, sd in sds.items():
            load_model.write_run(tmp, run_id, sd)
        est = load_model.load_named_model("megapose-1.0-RGB", ds, models_root=Path(tmp))
    est.sharder = HypothesisSharder(enabled=world > 1)
    lib = _abi.lib()

    images_dev, K_dev, bboxes_dev = images.cuda(), K.cuda(), bboxes.cuda()
    images_pin, K_pin, bboxes_pin = images.pin_memory(), K.pin_memory(), bboxes.pin_memory()

    def step_device():
        obs = ObservationTensor(images_dev, K_dev)
        det = PandasTensorCollection(det_df.copy(), bboxes=bboxes_dev)
        final, _ = est.run_inference_pipeline(obs, detections=det, n_refiner_iterations=N_REFINER_ITERS, n_pose_hypotheses=1)
        return final

    def step_e2e():
        obs = ObservationTensor(images_pin.cuda(non_blocking=True), K_pin.cuda(non_blocking=True))
        det = PandasTensorCollection(det_df.copy(), bboxes=bboxes_pin.cuda(non_blocking=True))
        final, _ = est.run_inference_pipeline(obs, detections=det, n_refiner_iterations=N_REFINER_ITERS, n_pose_hypotheses=1)
        poses = final.poses.cpu()
        scores = final.infos["pose_score"].values
        return poses, scores

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    for _ in range(max(3, args.warmup)):
        step_device()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    total_ms = timed(step_device, args.steps)
    for _ in range(2):
        step_e2e()
    e2e_ms = timed(step_e2e, args.steps)
    clocks = sampler.stop() if rank == 0 else None  # sampled over both timed regions (device-resident and end-to-end)

    # roofline of the dominant kernels (the convolutions of the 576-hypothesis coarse forward): CUDA events around every
    # conv launch, same workload.  The coarse forward normally replays a CUDA graph (its launches cannot be timed one by
    # one), so for this pass it is launched eagerly; the small-batch forwards (<= 64 rows) keep replaying graphs.
    saved_gmb = est.coarse_model.graph_max_batch
    est.coarse_model.graph_max_batch = 64
    lib.mpx_profile_enable(1)
    prof_steps = 2
    for _ in range(prof_steps):
        step_device()
    conv_ms, conv_fl, conv_n = ctypes.c_double(), ctypes.c_double(), ctypes.c_longlong()
    _abi.check(lib.mpx_profile_summary(ctypes.byref(conv_ms), ctypes.byref(conv_fl), ctypes.byref(conv_n)))
    lib.mpx_profile_enable(0)
    est.coarse_model.graph_max_batch = saved_gmb
    # kernels of libmpx.so per step: counted on one step launched eagerly (graph replays execute the same kernels but do
    # not pass through the library's launch counter)
    flags = (est.coarse_model.use_cuda_graphs, est.refiner_model.use_cuda_graphs)
    est.coarse_model.use_cuda_graphs = est.refiner_model.use_cuda_graphs = False
    lib.mpx_net_set_graphs(0)
    l0 = lib.mpx_launch_count()
    step_device()
    torch.cuda.synchronize()
    launches = torch.tensor([(lib.mpx_launch_count() - l0) * args.steps], device="cuda", dtype=torch.float64)
    lib.mpx_net_set_graphs(1)
    est.coarse_model.use_cuda_graphs, est.refiner_model.use_cuda_graphs = flags
    if world > 1:
        dist.all_reduce(launches)
    barrier()

    hyp_per_step = M_GRID * n_gpus
    ms_per_step = total_ms / args.steps
    value = hyp_per_step / (ms_per_step / 1000.0)
    e2e_value = hyp_per_step / (e2e_ms / args.steps / 1000.0)
    if rank == 0:
        peaks_path = ROOT / "MEASURED_PEAKS.json"
        if peaks_path.exists():
            peaks = json.loads(peaks_path.read_text())
            peak, peak_src = float(peaks.get("bf16_tflops_sustained", 1400.0)), "MEASURED_PEAKS.json bf16_tflops_sustained (of measured)"
        else:
            peak, peak_src = 1400.0, "fallback sustained figure of B200_PROFILING.md (of fallback)"
        conv_ms_per_step = conv_ms.value / prof_steps
        # The event-timed launches are the 36 convolutions of the 576-hypothesis coarse forward (the refiner iterations and
        # the final scoring replay CUDA graphs, whose kernels are not individually timed): algorithmic work of exactly
        # those launches = 576 x 12.068 GFLOP (coarse model, FLOP = 2*MAC)
        launches_per_step = conv_n.value / prof_steps
        gflop_per_hyp = GFLOP_COARSE_PER_HYP if abs(launches_per_step - 36.0) < 0.5 else GFLOP_PER_HYP
        algo_tflop_per_step = gflop_per_hyp * M_GRID / 1000.0  # this rank's unit: one object x 576
        achieved = algo_tflop_per_step / (conv_ms_per_step / 1000.0)
        traffic = None
        tpath = Path(__file__).resolve().parent / "profiles" / "conv_traffic.json"
        if tpath.exists():  # dram__bytes_read.sum + dram__bytes_write.sum of the same 36 launches, one ncu pass (see file)
            tj = json.loads(tpath.read_text())
            traffic = float(tj["dram_read_bytes"]) + float(tj["dram_write_bytes"])
        line = {
            "metric": "pose hypotheses/sec through render+coarse+5x refine", "value": value, "unit": "hypotheses/s",
            "impl": "mpx", "n_gpus": n_gpus, "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": workload_config(n_gpus),
            "e2e": {"value": e2e_value, "unit": "hypotheses/s",
                    "h2d_bytes_per_step": int(images.numel() * 4 + K.numel() * 4 + bboxes.numel() * 4),
                    "d2h_bytes_per_step": int(n_gpus * 16 * 4 + n_gpus * 4 * 2 + M_GRID * n_gpus * 4 * 2)},
            "gpu_launches": int(launches.item()),
            "clocks": clocks,
            "roofline": {"bound": "tensor",
                         "kernel": "tcgen05 implicit-GEMM convolutions (conv_window / conv_igemm / conv_igemm2): the 36 "
                                   "launches of the 576-hypothesis coarse forward, summed",
                         "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak, "traffic": traffic,
                         "traffic_unit": "bytes of DRAM traffic of the same 36 launches (ncu), vs "
                                         "algorithmic_activation_bytes",
                         "algorithmic_activation_bytes": ALGO_CONV_BYTES_PER_HYP * M_GRID,
                         "peak_source": peak_src, "algorithmic_tflop_per_step": algo_tflop_per_step,
                         "executed_tflop_per_step": conv_fl.value / prof_steps / 1e12,
                         "conv_ms_per_step": conv_ms_per_step, "conv_launches_per_step": conv_n.value / prof_steps,
                         "conv_share_of_step": conv_ms_per_step / ms_per_step},
        }
        if n_gpus == 1 and not args.no_cpu_baseline:
            v, cores, sample, _ = cpu_sample_rate(build_scene(1), args.cpu_sample)
            line["cpu_baseline"] = {"value": v, "unit": "hypotheses/s", "cores": cores, "kind": "port", "sample": sample}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse_args()
    if a.impl == "reference":
        run_reference_arm(a)
    else:
        run_mpx_arm(a)
