"""Is the window kernel bound by DRAM-sourced TMA loads?  Same layer1 conv at batch sizes whose activations do / do
not fit the 126 MB L2 (time per 128-row tile should be equal if it is not)."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch  # noqa: E402
import gpu_probe_window as W  # noqa: E402

for mode in (11, 10):
    for n in (37, 74, 148, 296, 576):
        W.run(n, 60, 80, 3, 3, (1, 1, 1, 1), True, False, mode, check=False, iters=20)
W.lib.mpx_conv_set_mode(11)
