"""Window-conv bring-up: correctness vs torch for the 64->64 stride-1 shapes and timing vs the im2col kernel."""
import sys
from pathlib import Path

import torch
import torch.nn.functional as F

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from megapose6d_b200 import _abi  # noqa: E402

lib = _abi.lib()
torch.backends.cudnn.allow_tf32 = False


def run(n, h, w, r, s, pads, relu, use_res, mode, check=True, iters=0):
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn(n, h, w, 64, device="cuda", generator=g).to(torch.bfloat16)
    wt = (torch.randn(64, r, s, 64, device="cuda", generator=g) / (r * s * 64) ** 0.5).to(torch.bfloat16)
    bias = torch.randn(64, device="cuda", generator=g)
    res = torch.randn(n, h, w, 64, device="cuda", generator=g).to(torch.bfloat16) if use_res else None
    out = torch.full((n, h, w, 64), float("nan"), device="cuda", dtype=torch.bfloat16)
    lib.mpx_conv_set_mode(mode)

    def call():
        return lib.mpx_conv2d_bf16(_abi.ptr(x), n, h, w, 64, _abi.ptr(wt.view(64, -1)), _abi.ptr(bias), 64, r, s, 1, pads[0],
                                   pads[1], pads[2], pads[3], int(relu), _abi.ptr(res), _abi.ptr(out), 0, 0, _abi.stream_ptr())

    rc = call()
    torch.cuda.synchronize()
    msg = f"mode={mode} n={n} {h}x{w} {r}x{s} pads={pads} res={use_res} rc={rc}"
    if rc != 0:
        print(msg, lib.mpx_last_error().decode())
        return
    if check:
        xf = F.pad(x.float().permute(0, 3, 1, 2), (pads[1], pads[3], pads[0], pads[2]))
        ref = F.conv2d(xf, wt.float().permute(0, 3, 1, 2), bias=bias).permute(0, 2, 3, 1)
        if res is not None:
            ref = ref + res.float()
        if relu:
            ref = torch.relu(ref)
        o = out.float()
        err = (o - ref).abs()
        nan = int(torch.isnan(o).sum())
        ok = nan == 0 and err.max().item() <= 2 ** -7 * ref.abs().max().item() + 1e-2
        msg += f" max_err={err.max().item():.4g} nan={nan} {'OK' if ok else 'MISMATCH'}"
        if not ok:
            bad = ((err > 0.05) | torch.isnan(o)).any(dim=-1)
            idx = bad.nonzero()
            msg += f" bad_pixels={idx.shape[0]} first={idx[:6].tolist()}"
    if iters:
        for _ in range(3):
            call()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            call()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        fl = 2.0 * n * h * w * 64 * r * s * 64
        msg += f" | {ms:.3f} ms {fl / ms / 1e9:.0f} TFLOP/s"
    print(msg)


for mode in (1, 0):
    run(1, 12, 20, 3, 3, (1, 1, 1, 1), False, False, mode)
    run(2, 12, 20, 3, 3, (1, 1, 1, 1), True, True, mode)
    run(3, 60, 80, 3, 3, (1, 1, 1, 1), True, True, mode)
    run(2, 24, 32, 4, 4, (2, 2, 1, 1), True, False, mode)
    run(2, 120, 160, 4, 4, (2, 2, 1, 1), True, False, mode)
    run(5, 7, 9, 3, 3, (1, 1, 1, 1), False, False, mode)
for mode in (1, 0):
    run(576, 60, 80, 3, 3, (1, 1, 1, 1), True, True, mode, check=False, iters=10)
    run(576, 120, 160, 4, 4, (2, 2, 1, 1), True, False, mode, check=False, iters=5)
