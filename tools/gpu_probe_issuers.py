"""Window kernel: one vs two MMA-issuing threads (mode bit 5 = single issuer), same box, interleaved."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import gpu_probe_window as W  # noqa: E402

for mode in (11, 43, 11 + 64, 11 + 128, 11 + 64 + 128):
    W.run(3, 60, 80, 3, 3, (1, 1, 1, 1), True, True, mode)
    W.run(2, 24, 32, 4, 4, (2, 2, 1, 1), True, False, mode)
    W.run(5, 7, 9, 3, 3, (1, 1, 1, 1), False, False, mode)
for rep in range(2):
    for mode in (11, 43, 11 + 64, 11 + 128, 11 + 64 + 128):
        W.run(576, 60, 80, 3, 3, (1, 1, 1, 1), True, False, mode, check=False, iters=10)
        W.run(576, 60, 80, 3, 3, (1, 1, 1, 1), True, True, mode, check=False, iters=10)
        W.run(576, 120, 160, 4, 4, (2, 2, 1, 1), True, False, mode, check=False, iters=5)
W.lib.mpx_conv_set_mode(11)
