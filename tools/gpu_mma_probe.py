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
for cg in (1, 2):
    for n in (64, 128, 256):
        for issuers in (1, 2):
            for chains in (1, 2, 4):
                if issuers * chains * n > 512:
                    continue
                v = ctypes.c_double()
                _abi.check(lib.mpx_debug_mma_probe(cg, n, chains, issuers, 4096, ctypes.byref(v)))
                floor = 128 * n / 256
                rows.append(dict(cta_group=cg, N=n, issuers=issuers, chains_per_issuer=chains, cycles_per_mma=round(v.value, 1),
                                 tensor_floor=floor, operand_bytes_per_sm=128 * 32 + n * 32 // cg))
                print(rows[-1], flush=True)
out = Path(sys.argv[1]) if len(sys.argv) > 1 else Path("gpurun_out/mma_probe.json")
out.parent.mkdir(parents=True, exist_ok=True)
out.write_text(json.dumps(rows, indent=1))
