"""BASELINE.json configs[4]: rasteriser microbench -- one 10k-triangle mesh x 4096 views @224x224, rgb + normals.

Reports the contract bytes written (fp32 NCHW planes, 6 channels: 1 204 224 B/view) per second against the
measured HBM copy bandwidth, and the same for the fused 16-bit output path, for the tiled kernel (default, raster mode 7)
and the untiled one (mode 3).  Inputs are resident in HBM; CUDA events.
"""
import json
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from megapose6d_b200 import _abi, procedural  # noqa: E402
from megapose6d_b200.renderer import BatchRenderer  # noqa: E402
from megapose6d_b200.so3 import load_SO3_grid  # noqa: E402


def main(n_views=4096, h=224, w=224, iters=5):
    ds = procedural.make_object_dataset(1)  # 10 000 triangles / 5 002 vertices
    r = BatchRenderer(object_dataset=ds)
    rng = np.random.RandomState(0)
    R = load_SO3_grid(576)[torch.arange(n_views) % 576]
    TCO = torch.eye(4).repeat(n_views, 1, 1)
    TCO[:, :3, :3] = R
    TCO[:, 2, 3] = torch.from_numpy(rng.uniform(0.4, 1.2, n_views)).float()
    # focal length so that the ~10 cm object fills about half of the frame at the mean distance
    f = 0.5 * w * 0.8 / 0.05 / 2
    K = torch.tensor([[f, 0, w / 2], [0, f, h / 2], [0, 0, 1]]).repeat(n_views, 1, 1)
    labels = [ds[0].label] * n_views
    TCO, K = TCO.cuda(), K.cuda()
    out = r.render(labels, TCO, K, None, (h, w), render_normals=True)
    torch.cuda.synchronize()
    cover = (out.rgbs.sum(1) > 0).float().mean().item()
    lab = r.mesh_db.label_ids(labels, "cuda")
    rgbs = torch.empty(n_views, 3, h, w, device="cuda")
    nrms = torch.empty(n_views, 3, h, w, device="cuda")
    ws = r.workspace(h, w, "cuda")
    lib = _abi.lib()

    def contract():
        _abi.check(lib.mpx_raster_render(r.mesh_db.handle, _abi.ptr(lab), _abi.ptr(TCO), _abi.ptr(K), n_views, h, w, r.flags,
                                         _abi.ptr(rgbs), _abi.ptr(nrms), None, _abi.ptr(ws), ws.numel(), _abi.stream_ptr()))

    x = torch.zeros(n_views, h // 2, w // 2, 64, device="cuda", dtype=_abi.act_dtype())

    def fused():
        r.render_fused(lab, TCO, K, 1, (h, w), x, 16, 3, 6)

    res = {}
    for mode, name, fn, nbytes in ((7, "contract_fp32_nchw", contract, 4 * h * w * 6 + 100), (7, "fused_16bit", fused, 2 * h * w * 6 + 100),
                                   (3, "untiled_contract_fp32_nchw", contract, 4 * h * w * 6 + 100),
                                   (3, "untiled_fused_16bit", fused, 2 * h * w * 6 + 100)):
        lib.mpx_raster_set_mode(mode)
        for _ in range(2):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        res[name] = {"ms": ms, "views_per_s": n_views / ms * 1e3, "GBps": nbytes * n_views / ms / 1e6, "bytes_per_view": nbytes}
    lib.mpx_raster_set_mode(7)
    peaks = ROOT / "MEASURED_PEAKS.json"
    peak = json.loads(peaks.read_text())["hbm_gbs"] if peaks.exists() else 6650.0
    for v in res.values():
        v["frac_of_hbm_peak"] = v["GBps"] / peak
    print(json.dumps({"workload": f"10k-triangle mesh x {n_views} views @{h}x{w}, rgb+normals", "coverage": cover,
                      "hbm_peak_GBps": peak, "results": res}))


if __name__ == "__main__":
    main()
