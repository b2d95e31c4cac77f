"""Coarse stem (4x4 taps over the space-to-depth input): row-group choice of the window kernel (mode bit 10 = 1024 forces
the older rg = 2 / 2 stages), correctness + timing, interleaved on one box."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import gpu_probe_window as W  # noqa: E402

for mode in (11, 11 + 1024):
    W.run(2, 24, 32, 4, 4, (2, 2, 1, 1), True, False, mode)
    W.run(2, 120, 160, 4, 4, (2, 2, 1, 1), True, False, mode)
    W.run(3, 60, 80, 3, 3, (1, 1, 1, 1), True, True, mode)
for rep in range(2):
    for mode in (11, 11 + 1024):
        W.run(576, 120, 160, 4, 4, (2, 2, 1, 1), True, False, mode, check=False, iters=5)
        W.run(576, 60, 80, 3, 3, (1, 1, 1, 1), True, False, mode, check=False, iters=10)
W.lib.mpx_conv_set_mode(11)
