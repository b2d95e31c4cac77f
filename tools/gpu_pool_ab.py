"""A/B of the stem's fused max-pool epilogue (mpx_conv_set_mode bit 21): stem + max-pool kernel vs memset + pooled stem, and
the whole coarse forward under both schedules.  CUDA events, batch 576 at 240x320 (tensors >> L2)."""
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from megapose6d_b200 import _abi  # noqa: E402
from megapose6d_b200.backbone import ResNet34Engine  # noqa: E402
from workloads import weights as W  # noqa: E402


def time_ms(fn, iters=10, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    n, h, w = int(sys.argv[1]) if len(sys.argv) > 1 else 576, 240, 320
    lib, act = _abi.lib(), _abi.act_dtype()
    g = torch.Generator(device="cuda").manual_seed(0)
    hs, ws, c_pad = h // 2, w // 2, 16
    xm = torch.randn(n, hs, ws, 4 * c_pad, device="cuda", generator=g).to(act)
    wm = (torch.randn(64, 16 * 4 * c_pad, device="cuda", generator=g) / 21.0).to(act)
    bias = torch.randn(64, device="cuda", generator=g)
    full = torch.empty(n, hs, ws, 64, device="cuda", dtype=act)
    pooled = torch.empty(n, hs // 2, ws // 2, 64, device="cuda", dtype=act)
    pooled2 = torch.empty_like(pooled)
    s = _abi.stream_ptr()

    def stem(flags, out):
        rc = lib.mpx_conv2d(_abi.ptr(xm), n, hs, ws, 4 * c_pad, _abi.ptr(wm), _abi.ptr(bias), 64, 4, 4, 1, 2, 2, 1, 1, flags, None,
                            _abi.ptr(out), 0, 0, s)
        assert rc == 0, rc

    def pool():
        _abi.check(lib.mpx_maxpool3x3s2(_abi.ptr(full), n, hs, ws, 64, _abi.ptr(pooled), s))

    def fused():
        pooled2.zero_()
        stem(3 | 4, pooled2)

    rec = dict(batch=n)
    rec["stem_ms"] = time_ms(lambda: stem(3, full))
    rec["maxpool_ms"] = time_ms(pool)
    rec["stem_then_pool_ms"] = time_ms(lambda: (stem(3, full), pool()))
    rec["memset_ms"] = time_ms(lambda: pooled2.zero_())
    rec["memset_fused_stem_ms"] = time_ms(fused)
    stem(3, full); pool(); fused()
    torch.cuda.synchronize()
    rec["equal"] = bool(torch.equal(pooled.float(), pooled2.float()))
    sd = W.make_state_dict(W.COARSE_CFG, 1)
    for name, mode in (("separate_pool", 60866571 & ~2097152), ("fused_pool", 60866571)):
        lib.mpx_conv_set_mode(mode)
        eng = ResNet34Engine(sd, n_inputs=9, head="views_logits_head")
        x = eng.alloc_input(n, h, w)
        x.copy_(torch.rand(x.shape, device="cuda").to(act))
        rec[f"network_{name}_ms"] = time_ms(lambda: eng.forward(x, h, w), iters=8)
        del eng, x
        torch.cuda.empty_cache()
    lib.mpx_conv_set_mode(60866571)
    print(json.dumps(rec))


if __name__ == "__main__":
    main()
