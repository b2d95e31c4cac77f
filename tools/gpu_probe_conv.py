"""Bring-up probe for the tcgen05 convolution (run on the GPU box):
    python tools/gpu_probe_conv.py
Prints one line per case; never raises on numerical mismatch so that one run reports everything.
"""
import sys
import time
from pathlib import Path

import torch
import torch.nn.functional as F

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from megapose6d_b200 import _abi  # noqa: E402


def conv_ref(x, w, bias, stride, pads, relu, residual):
    # x [n,h,w,c] bf16; w [co, r, s, c] bf16
    xf = x.float().permute(0, 3, 1, 2)
    wf = w.float().permute(0, 3, 1, 2)
    plh, plw, phh, phw = pads
    xf = F.pad(xf, (plw, phw, plh, phh))
    y = F.conv2d(xf, wf, bias=bias, stride=stride)
    y = y.permute(0, 2, 3, 1)
    if residual is not None:
        y = y + residual.float()
    if relu:
        y = torch.relu(y)
    return y


def run_case(name, n, h, w, cin, cout, r, s, stride, pads, relu, use_res, block_n=0, max_ctas=0, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    x = (torch.randn(n, h, w, cin, device="cuda", generator=g)).to(torch.bfloat16).contiguous()
    wt = (torch.randn(cout, r, s, cin, device="cuda", generator=g) / (r * s * cin) ** 0.5).to(torch.bfloat16).contiguous()
    bias = torch.randn(cout, device="cuda", generator=g).float().contiguous()
    plh, plw, phh, phw = pads
    p = (h + plh + phh - r) // stride + 1
    q = (w + plw + phw - s) // stride + 1
    res = torch.randn(n, p, q, cout, device="cuda", generator=g).to(torch.bfloat16).contiguous() if use_res else None
    out = torch.full((n, p, q, cout), float("nan"), device="cuda", dtype=torch.bfloat16)
    lib = _abi.lib()
    rc = lib.mpx_conv2d_bf16(_abi.ptr(x), n, h, w, cin, _abi.ptr(wt.view(cout, -1)), _abi.ptr(bias), cout, r, s,
                             stride, plh, plw, phh, phw, int(relu), _abi.ptr(res), _abi.ptr(out), block_n,
                             max_ctas, _abi.stream_ptr())
    if rc != 0:
        print(f"[{name}] CALL FAILED rc={rc}: {lib.mpx_last_error().decode()}")
        return False
    try:
        torch.cuda.synchronize()
    except Exception as e:  # noqa: BLE001
        print(f"[{name}] KERNEL ERROR: {e}")
        raise
    ref = conv_ref(x, wt, bias, stride, pads, relu, res)
    o = out.float()
    nan = int(torch.isnan(o).sum())
    err = (o - ref).abs()
    tol = 2e-2 * ref.abs().max().item() + 1e-2
    ok = nan == 0 and err.max().item() <= tol
    print(f"[{name}] n={n} {h}x{w} cin={cin} cout={cout} {r}x{s}/s{stride} pads={pads} bn={block_n} "
          f"-> max_err={err.max().item():.4g} mean_err={err.mean().item():.4g} ref_max={ref.abs().max().item():.3g} "
          f"nan={nan} {'OK' if ok else 'MISMATCH'}")
    if not ok:
        # localise: which rows (pixels) / columns are wrong
        bad = (err > tol) | torch.isnan(o)
        rows = bad.view(-1, cout).any(dim=1).nonzero().flatten()
        cols = bad.view(-1, cout).any(dim=0).nonzero().flatten()
        print(f"    bad rows: {rows.numel()} of {bad.view(-1, cout).shape[0]} (first {rows[:12].tolist()}); "
              f"bad cols: {cols.numel()} (first {cols[:12].tolist()})")
        print("    out[0,0,0,:8]", o[0, 0, 0, :8].tolist())
        print("    ref[0,0,0,:8]", ref[0, 0, 0, :8].tolist())
    return ok


def main():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    print(torch.cuda.get_device_name(0), torch.version.cuda)
    cases = [
        ("gemm1x1", 1, 8, 16, 64, 64, 1, 1, 1, (0, 0, 0, 0), False, False),
        ("gemm1x1_k128", 1, 8, 16, 128, 64, 1, 1, 1, (0, 0, 0, 0), False, False),
        ("gemm1x1_n128", 1, 8, 16, 64, 128, 1, 1, 1, (0, 0, 0, 0), False, False),
        ("c3x3", 2, 12, 20, 64, 64, 3, 3, 1, (1, 1, 1, 1), False, False),
        ("c3x3_relu_res", 2, 12, 20, 64, 64, 3, 3, 1, (1, 1, 1, 1), True, True),
        ("c3x3_s2", 2, 30, 40, 64, 128, 3, 3, 2, (1, 1, 1, 1), True, False),
        ("c1x1_s2", 2, 30, 40, 64, 128, 1, 1, 2, (0, 0, 0, 0), False, False),
        ("odd_s2", 3, 15, 20, 128, 256, 3, 3, 2, (1, 1, 1, 1), True, False),
        ("odd_1x1_s2", 3, 15, 20, 128, 256, 1, 1, 2, (0, 0, 0, 0), False, False),
        ("stem4x4", 2, 24, 32, 64, 64, 4, 4, 1, (2, 2, 1, 1), True, False),
        ("stem4x4_c128", 2, 24, 32, 128, 64, 4, 4, 1, (2, 2, 1, 1), True, False),
        ("l4_bn256", 3, 8, 10, 512, 512, 3, 3, 1, (1, 1, 1, 1), True, True),
        ("l4_bn128", 3, 8, 10, 512, 512, 3, 3, 1, (1, 1, 1, 1), True, True, 128),
        ("l3_bn64", 3, 15, 20, 256, 256, 3, 3, 1, (1, 1, 1, 1), True, True, 64),
        ("persist", 16, 60, 80, 64, 64, 3, 3, 1, (1, 1, 1, 1), True, True),
        ("persist_fewctas", 4, 60, 80, 64, 64, 3, 3, 1, (1, 1, 1, 1), True, True, 0, 7),
    ]
    n_ok = 0
    for c in cases:
        try:
            n_ok += bool(run_case(*c))
        except Exception as e:  # noqa: BLE001
            print(f"[{c[0]}] EXCEPTION {type(e).__name__}: {e}")
            break
    print(f"{n_ok}/{len(cases)} cases OK")

    # quick timing of a layer1-like conv at batch 128
    try:
        n, h, w, c = 128, 60, 80, 64
        x = torch.randn(n, h, w, c, device="cuda").to(torch.bfloat16)
        wt = torch.randn(c, 9 * c, device="cuda").to(torch.bfloat16)
        bias = torch.zeros(c, device="cuda")
        out = torch.empty(n, h, w, c, device="cuda", dtype=torch.bfloat16)
        lib = _abi.lib()
        for cfg in [(64, 64), (128, 128), (256, 256), (512, 512)]:
            cin, cout = cfg
            hh, ww = {64: (60, 80), 128: (30, 40), 256: (15, 20), 512: (8, 10)}[cin]
            x = torch.randn(n, hh, ww, cin, device="cuda").to(torch.bfloat16)
            wt = (torch.randn(cout, 9 * cin, device="cuda") * 0.02).to(torch.bfloat16)
            bias = torch.zeros(cout, device="cuda")
            out = torch.empty(n, hh, ww, cout, device="cuda", dtype=torch.bfloat16)
            for it in range(3):
                lib.mpx_conv2d_bf16(_abi.ptr(x), n, hh, ww, cin, _abi.ptr(wt), _abi.ptr(bias), cout, 3, 3, 1, 1, 1,
                                    1, 1, 1, None, _abi.ptr(out), 0, 0, _abi.stream_ptr())
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            iters = 10
            for it in range(iters):
                lib.mpx_conv2d_bf16(_abi.ptr(x), n, hh, ww, cin, _abi.ptr(wt), _abi.ptr(bias), cout, 3, 3, 1, 1, 1,
                                    1, 1, 1, None, _abi.ptr(out), 0, 0, _abi.stream_ptr())
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / iters
            fl = 2.0 * n * hh * ww * cout * 9 * cin
            print(f"timing 3x3 {cin}->{cout} @{hh}x{ww} b{n}: {ms:.3f} ms  {fl / ms / 1e9:.1f} TFLOP/s")
    except Exception as e:  # noqa: BLE001
        print("timing failed:", e)


if __name__ == "__main__":
    t0 = time.time()
    main()
    print(f"done in {time.time() - t0:.1f}s")
