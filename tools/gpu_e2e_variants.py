"""Where does the end-to-end (host buffers in, host results out) frame pipeline lose time against device-resident inputs?
Times the in-flight loop of bench.py under variants of the host<->device copies."""
import json
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from megapose6d_b200.frame_pipeline import FramePipeline  # noqa: E402
from megapose6d_b200.tensor_collection import PandasTensorCollection  # noqa: E402
from megapose6d_b200.types import ObservationTensor  # noqa: E402
from workloads import scenes  # noqa: E402

STEPS = 24


def main():
    sc = scenes.bench_scene(1)
    images, K, det_df = sc["images"], sc["K"], sc["det_df"]
    kw = dict(n_refiner_iterations=5, n_pose_hypotheses=1)
    ests = [scenes.build_estimator(sc) for _ in range(2)]
    pipe = FramePipeline(None, estimators=ests)
    dev = dict(images=images.cuda(), K=K.cuda(), bboxes=sc["bboxes"].cuda())
    pin = dict(images=images.pin_memory(), K=K.pin_memory(), bboxes=sc["bboxes"].pin_memory())
    io = torch.cuda.Stream()
    for e in ests:
        for _ in range(5):
            e.run_inference_pipeline(ObservationTensor(dev["images"], dev["K"]),
                                     detections=PandasTensorCollection(det_df.copy(), bboxes=dev["bboxes"]), **kw)

    def frame(variant):
        if variant["h2d"] == "none":
            t = dev
        elif variant["h2d"] == "default":
            t = {k: v.cuda(non_blocking=True) for k, v in pin.items()}
        else:
            with torch.cuda.stream(io):
                t = {k: v.cuda(non_blocking=True) for k, v in pin.items()}
            torch.cuda.current_stream().wait_stream(io)
        return ObservationTensor(t["images"], t["K"]), PandasTensorCollection(det_df.copy(), bboxes=t["bboxes"])

    host_buf = torch.empty(1, 4, 4).pin_memory()

    def read(final, variant):
        if variant["d2h"] == "cpu":
            return final.poses.cpu(), final.infos["pose_score"].values
        if variant["d2h"] == "pinned":
            host_buf.copy_(final.poses, non_blocking=True)
            torch.cuda.current_stream().synchronize()
            return host_buf, final.infos["pose_score"].values
        return None

    def run(variant):
        host = 0.0
        for _ in range(STEPS):
            t0 = time.perf_counter()
            obs, det = frame(variant)
            done = pipe.submit(obs, det, **kw)
            if done is not None:
                read(done[0], variant)
            host += time.perf_counter() - t0
        for done in pipe.drain():
            read(done[0], variant)
        pipe.join()
        return host

    variants = [dict(h2d="none", d2h="none"), dict(h2d="none", d2h="cpu"), dict(h2d="default", d2h="none"),
                dict(h2d="default", d2h="cpu"), dict(h2d="io", d2h="cpu"), dict(h2d="io", d2h="pinned"),
                dict(h2d="default", d2h="pinned")]
    out = []
    for v in variants:
        best = 1e9
        for _ in range(3):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            t0 = time.perf_counter()
            run(v)
            e1.record()
            torch.cuda.synchronize()
            wall = (time.perf_counter() - t0) * 1e3 / STEPS
            best = min(best, e0.elapsed_time(e1) / STEPS)
        rec = dict(v, ms_per_frame=best, wall_ms_per_frame_last=wall)
        print(json.dumps(rec), flush=True)
        out.append(rec)
    Path("gpurun_out").mkdir(exist_ok=True)
    Path("gpurun_out/r02p_e2e_variants.json").write_text(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
