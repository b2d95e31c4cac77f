"""A/B of the kernel-selection bits on the bench workload (1 object x 576 hypotheses, 5 refiner iterations), one fresh
process per setting (the bits are read when libmpx.so is loaded; graphs captured under one setting must not leak into the
next).  Interleaves the settings over `--rounds` rounds so that clock / thermal drift hits all of them alike, and checks
that every setting returns the same survivor and a pose within tolerance of the default's.

    python tools/gpu_ab.py --conv 11,2059,4107 --raster 3 --steps 20 --rounds 3 --out gpurun_out/ab_modes.json

Mode bits: include/mpx.h (`mpx_conv_set_mode`, `mpx_raster_set_mode`); 2048 = arrival-gated window refills, 4096 / 8192 = pair-window
kernel for layer3/4 / layer2 (all written without a GPU at hand: run `MPX_EXPERIMENTAL=1 pytest tests/test_gpu_net.py -k experimental`
first).  A setting that traps or hangs is killed by the per-process timeout and reported as failed."""
import argparse
import json
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]


def child(steps: int, warmup: int) -> None:
    import torch

    sys.path.insert(0, str(ROOT))
    from workloads import scenes
    from megapose6d_b200 import _abi
    from megapose6d_b200.tensor_collection import PandasTensorCollection
    from megapose6d_b200.types import ObservationTensor

    sc = scenes.bench_scene(1)
    images, K, det_df, bboxes = sc["images"], sc["K"], sc["det_df"], sc["bboxes"]
    est = scenes.build_estimator(sc)
    images_dev, K_dev, bboxes_dev = images.cuda(), K.cuda(), bboxes.cuda()

    def step():
        det = PandasTensorCollection(det_df.copy(), bboxes=bboxes_dev)
        return est.run_inference_pipeline(ObservationTensor(images_dev, K_dev), detections=det,
                                          n_refiner_iterations=5, n_pose_hypotheses=1)

    for _ in range(max(3, warmup)):
        final, _ = step()
    torch.cuda.synchronize()
    lib = _abi.lib()
    lib.mpx_profile_enable(0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n0 = lib.mpx_launch_count()
    e0.record()
    for _ in range(steps):
        final, extra = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    print("AB_RESULT " + json.dumps(dict(
        ms_per_step=ms, hyp_per_s=576 / ms * 1e3, launches_per_step=(lib.mpx_launch_count() - n0) / steps,
        survivor=int(final.infos["hypothesis_id"].iloc[0]), pose_logit=float(final.infos["pose_logit"].iloc[0]),
        pose=final.poses[0].cpu().flatten().tolist())))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--conv", default="60866571", help="comma-separated mpx_conv_set_mode values")
    ap.add_argument("--raster", default="7", help="comma-separated mpx_raster_set_mode values")
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--rounds", type=int, default=3)
    ap.add_argument("--timeout", type=int, default=240, help="seconds per process")
    ap.add_argument("--out", type=Path, default=None)
    ap.add_argument("--child", action="store_true")
    args = ap.parse_args()
    if args.child:
        child(args.steps, args.warmup)
        return
    settings = [(int(c), int(r)) for c in args.conv.split(",") for r in args.raster.split(",")]
    results = {s: [] for s in settings}
    for rnd in range(args.rounds):
        for s in settings:
            env = dict(os.environ, MPX_CONV_MODE=str(s[0]), MPX_RASTER_MODE=str(s[1]))
            cmd = [sys.executable, str(Path(__file__).resolve()), "--child", "--steps", str(args.steps), "--warmup", str(args.warmup)]
            try:
                res = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=args.timeout)
                line = next((l for l in res.stdout.splitlines() if l.startswith("AB_RESULT ")), None)
                rec = json.loads(line[len("AB_RESULT "):]) if line else dict(failed=f"exit {res.returncode}: {res.stderr[-400:]}")
            except subprocess.TimeoutExpired:
                rec = dict(failed=f"timeout after {args.timeout} s")
            results[s].append(rec)
            print(f"round {rnd} conv={s[0]} raster={s[1]}: " + (f"{rec['ms_per_step']:.3f} ms" if "ms_per_step" in rec else rec["failed"]),
                  flush=True)
    base = next((r for r in results[settings[0]] if "pose" in r), None)
    summary = []
    for s in settings:
        ok = [r for r in results[s] if "ms_per_step" in r]
        rec = dict(conv_mode=s[0], raster_mode=s[1], runs=len(results[s]), failed=len(results[s]) - len(ok))
        if ok:
            ms = sorted(r["ms_per_step"] for r in ok)
            rec.update(ms_median=ms[len(ms) // 2], ms_min=ms[0], ms_max=ms[-1], launches_per_step=ok[0]["launches_per_step"],
                       survivor=ok[0]["survivor"])
            if base is not None:
                rec["same_survivor_as_first"] = ok[0]["survivor"] == base["survivor"]
                rec["max_pose_diff_vs_first"] = max(abs(a - b) for a, b in zip(ok[0]["pose"], base["pose"]))
        summary.append(rec)
    print(json.dumps(summary, indent=1))
    if args.out is not None:
        args.out.parent.mkdir(parents=True, exist_ok=True)
        args.out.write_text(json.dumps(dict(summary=summary, runs={f"{k[0]}/{k[1]}": v for k, v in results.items()}), indent=1))


if __name__ == "__main__":
    main()
