"""CTA-pair (cta_group::2) conv bring-up: correctness vs torch on the wide layers and timing vs the single-CTA kernel."""
import sys
from pathlib import Path

import torch
import torch.nn.functional as F

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from megapose6d_b200 import _abi  # noqa: E402

lib = _abi.lib()
torch.backends.cudnn.allow_tf32 = False


def run(name, n, h, w, cin, cout, r, s, stride, pads, relu, use_res, mode, check=True, iters=0):
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn(n, h, w, cin, device="cuda", generator=g).to(torch.bfloat16)
    wt = (torch.randn(cout, r, s, cin, device="cuda", generator=g) / (r * s * cin) ** 0.5).to(torch.bfloat16)
    bias = torch.randn(cout, device="cuda", generator=g)
    p = (h + pads[0] + pads[2] - r) // stride + 1
    q = (w + pads[1] + pads[3] - s) // stride + 1
    res = torch.randn(n, p, q, cout, device="cuda", generator=g).to(torch.bfloat16) if use_res else None
    out = torch.full((n, p, q, cout), float("nan"), device="cuda", dtype=torch.bfloat16)
    lib.mpx_conv_set_mode(mode)

    def call():
        return lib.mpx_conv2d_bf16(_abi.ptr(x), n, h, w, cin, _abi.ptr(wt.view(cout, -1)), _abi.ptr(bias), cout, r, s, stride,
                                   pads[0], pads[1], pads[2], pads[3], int(relu), _abi.ptr(res), _abi.ptr(out), 0, 0,
                                   _abi.stream_ptr())

    rc = call()
    torch.cuda.synchronize()
    msg = f"[{name}] mode={mode} n={n} {h}x{w} {cin}->{cout} {r}x{s}/s{stride} rc={rc}"
    if rc != 0:
        print(msg, lib.mpx_last_error().decode(), flush=True)
        return
    if check:
        xf = F.pad(x.float().permute(0, 3, 1, 2), (pads[1], pads[3], pads[0], pads[2]))
        ref = F.conv2d(xf, wt.float().permute(0, 3, 1, 2), bias=bias, stride=stride).permute(0, 2, 3, 1)
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
            bad = ((err > 0.05) | torch.isnan(o)).view(-1, cout)
            rows = bad.any(1).nonzero().flatten()
            cols = bad.any(0).nonzero().flatten()
            msg += f" bad_rows={rows.numel()}/{bad.shape[0]} first={rows[:8].tolist()} bad_cols={cols.numel()} first={cols[:8].tolist()}"
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
        fl = 2.0 * n * p * q * cout * r * s * cin
        msg += f" | {ms:.3f} ms {fl / ms / 1e9:.0f} TFLOP/s"
    print(msg, flush=True)


CASES = [
    ("one_pair_tile", 1, 16, 16, 64, 128, 1, 1, 1, (0, 0, 0, 0), False, False),
    ("odd_tiles", 1, 12, 32, 64, 128, 3, 3, 1, (1, 1, 1, 1), True, False),
    ("l2", 4, 30, 40, 128, 128, 3, 3, 1, (1, 1, 1, 1), True, True),
    ("l2_s2", 4, 60, 80, 64, 128, 3, 3, 2, (1, 1, 1, 1), True, False),
    ("l2_ds", 4, 60, 80, 64, 128, 1, 1, 2, (0, 0, 0, 0), False, False),
    ("l3", 5, 15, 20, 256, 256, 3, 3, 1, (1, 1, 1, 1), True, True),
    ("l4", 3, 8, 10, 512, 512, 3, 3, 1, (1, 1, 1, 1), True, True),
    ("l4_many", 40, 8, 10, 512, 512, 3, 3, 1, (1, 1, 1, 1), True, True),
]
for c in CASES:
    run(*c, mode=3)
for mode in (3, 1):
    run("t_l2", 576, 30, 40, 128, 128, 3, 3, 1, (1, 1, 1, 1), True, True, mode, check=False, iters=10)
    run("t_l3", 576, 15, 20, 256, 256, 3, 3, 1, (1, 1, 1, 1), True, True, mode, check=False, iters=10)
    run("t_l4", 576, 8, 10, 512, 512, 3, 3, 1, (1, 1, 1, 1), True, True, mode, check=False, iters=10)
