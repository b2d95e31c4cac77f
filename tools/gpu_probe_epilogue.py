"""Window-kernel epilogue variants (mode bits 16 = staged, 32/64 = warp sets - 1, 128 = force a single set):
correctness vs torch on small cases, then timing of layer1 (with / without residual) and the coarse stem."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import gpu_probe_window as W  # noqa: E402  (runs its own checks on import)

MODES = {"1set": 11 + 128, "2sets": 11 + 32, "3sets": 11, "4sets": 11 + 96, "1set_staged": 11 + 128 + 16,
         "2sets_staged": 11 + 32 + 16, "3sets_staged": 11 + 16}
for name, mode in MODES.items():
    print("==", name, mode)
    W.run(3, 60, 80, 3, 3, (1, 1, 1, 1), True, True, mode)
    W.run(2, 24, 32, 4, 4, (2, 2, 1, 1), True, False, mode)
    W.run(5, 7, 9, 3, 3, (1, 1, 1, 1), False, False, mode)
    W.run(576, 60, 80, 3, 3, (1, 1, 1, 1), True, False, mode, check=False, iters=10)
    W.run(576, 60, 80, 3, 3, (1, 1, 1, 1), True, True, mode, check=False, iters=10)
    W.run(576, 120, 160, 4, 4, (2, 2, 1, 1), True, False, mode, check=False, iters=5)
W.lib.mpx_conv_set_mode(11)
