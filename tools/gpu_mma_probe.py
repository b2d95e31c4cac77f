"""Cycles per tcgen05.mma (kind::f16, both operands in shared memory) as a function of N, cta_group, the number of
independent accumulate chains an issuing thread interleaves and the number of issuing threads (mpx_debug_mma_probe).
The tensor floor is 128 * N / 256 cycles per SM for either cta_group (B300_MICROARCH.md); what the table shows on top of it
is the cost of dependent accumulates and of reading the operands from shared memory."""
import ctypes
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from megapose6d_b200 import _abi  # noqa: E402

lib = _abi.lib()
rows = []
# warp-uniform issue path (all lanes execute, elect.sync inside the instruction): cta_group 1
for n in (64, 128, 256):
    for issuers in (1, 2, 3, 4):
        for chains in (1, 2):
            if issuers * chains * n > 512:
                continue
            v, vi = ctypes.c_double(), ctypes.c_double()
            _abi.check(lib.mpx_debug_mma_probe(1, n, chains, issuers | 0x100, 8192, ctypes.byref(v), ctypes.byref(vi)))
            rows.append(dict(path="warp-uniform", cta_group=1, N=n, issuers=issuers, chains_per_issuer=chains,
                             cycles_per_mma=round(v.value, 1), issue_cycles_per_instruction=round(vi.value, 1),
                             tensor_floor=128 * n / 256, operand_bytes_per_sm=128 * 32 + n * 32))
            print(rows[-1], flush=True)
for cg in (1, 2):
    for n in (64, 128, 256):
        for issuers in (1, 2, 3, 4):
            for chains in (1, 2):
                if issuers * chains * n > 512:
                    continue
                v, vi = ctypes.c_double(), ctypes.c_double()
                _abi.check(lib.mpx_debug_mma_probe(cg, n, chains, issuers, 8192, ctypes.byref(v), ctypes.byref(vi)))
                floor = 128 * n / 256
                rows.append(dict(cta_group=cg, N=n, issuers=issuers, chains_per_issuer=chains, cycles_per_mma=round(v.value, 1),
                                 issue_cycles_per_instruction=round(vi.value, 1),
                                 tensor_floor=floor, operand_bytes_per_sm=128 * 32 + n * 32 // cg))
                print(rows[-1], flush=True)
out = Path(sys.argv[1]) if len(sys.argv) > 1 else Path("gpurun_out/mma_probe.json")
out.parent.mkdir(parents=True, exist_ok=True)
out.write_text(json.dumps(rows, indent=1))
