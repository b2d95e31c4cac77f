"""Layer2 window kernel bring-up: correctness vs torch, timing vs the im2col kernel (mode bit 256 disables window2)."""
import sys
from pathlib import Path

import torch
import torch.nn.functional as F

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from megapose6d_b200 import _abi  # noqa: E402

lib = _abi.lib()
torch.backends.cudnn.allow_tf32 = False


def run(n, h, w, relu, use_res, mode, max_ctas=0, check=True, iters=0):
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn(n, h, w, 128, device="cuda", generator=g).to(torch.bfloat16)
    wt = (torch.randn(128, 3, 3, 128, device="cuda", generator=g) / (9 * 128) ** 0.5).to(torch.bfloat16)
    bias = torch.randn(128, device="cuda", generator=g)
    res = torch.randn(n, h, w, 128, device="cuda", generator=g).to(torch.bfloat16) if use_res else None
    out = torch.full((n, h, w, 128), float("nan"), device="cuda", dtype=torch.bfloat16)
    lib.mpx_conv_set_mode(mode)

    def call():
        return lib.mpx_conv2d_bf16(_abi.ptr(x), n, h, w, 128, _abi.ptr(wt.view(128, -1)), _abi.ptr(bias), 128, 3, 3, 1, 1, 1, 1, 1,
                                   int(relu), _abi.ptr(res), _abi.ptr(out), 0, max_ctas, _abi.stream_ptr())

    rc = call()
    try:
        torch.cuda.synchronize()
    except Exception as e:  # noqa: BLE001
        print(f"mode={mode} n={n} {h}x{w} LAUNCH FAILED: {str(e)[:120]}")
        raise
    msg = f"mode={mode} n={n} {h}x{w} res={use_res} ctas={max_ctas} rc={rc}"
    if rc != 0:
        print(msg, lib.mpx_last_error().decode())
        return
    if check:
        ref = F.conv2d(x.float().permute(0, 3, 1, 2), wt.float().permute(0, 3, 1, 2), bias=bias, padding=1).permute(0, 2, 3, 1)
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
            msg += f" bad_pixels={idx.shape[0]}/{bad.numel()} first={idx[:6].tolist()}"
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
        fl = 2.0 * n * h * w * 128 * 9 * 128
        msg += f" | {ms:.3f} ms {fl / ms / 1e9:.0f} TFLOP/s"
    print(msg, flush=True)


if __name__ == "__main__":
    run(1, 12, 16, True, False, 11, max_ctas=1)
    run(4, 30, 40, True, True, 11, max_ctas=4)
    run(9, 30, 40, True, False, 11, max_ctas=3)
    run(3, 17, 23, False, True, 11, max_ctas=2)
    run(11, 5, 7, True, True, 11, max_ctas=2)
    for rep in range(2):
        for mode in (11, 11 | 256):
            run(576, 30, 40, True, False, mode, check=False, iters=10)
            run(576, 30, 40, True, True, mode, check=False, iters=10)
    lib.mpx_conv_set_mode(11)
