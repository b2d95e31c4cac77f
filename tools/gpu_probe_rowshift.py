"""Probe: can a UMMA A-descriptor start r0 rows into a 128B-swizzled, TMA-written tile (and does it need base_offset)?"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from megapose6d_b200 import _abi  # noqa: E402

lib = _abi.lib()
lib.mpx_debug_umma_rowshift.argtypes = [_abi.c_void_p, _abi.c_void_p, _abi.c_int, _abi.c_int, _abi.c_void_p, _abi.c_void_p]
g = torch.Generator(device="cuda").manual_seed(0)
A = torch.randn(144, 64, device="cuda", generator=g).to(torch.bfloat16)
B = torch.randn(64, 64, device="cuda", generator=g).to(torch.bfloat16)
for r0 in range(0, 10):
    for bo in sorted({0, r0 & 7}):
        out = torch.full((128, 64), float("nan"), device="cuda")
        rc = lib.mpx_debug_umma_rowshift(_abi.ptr(A), _abi.ptr(B), r0, bo, _abi.ptr(out), _abi.stream_ptr())
        torch.cuda.synchronize()
        ref = A[r0:r0 + 128].float() @ B.float().t()
        err = (out - ref).abs().max().item()
        print(f"r0={r0} base_offset={bo} rc={rc} max_err={err:.4g} {'OK' if err < 1e-2 else 'MISMATCH'}")
