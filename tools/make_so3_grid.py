"""Convert the reference's SO(3) grid text files (x y z w unit quaternions, one per line;
/root/reference/src/megapose/data/data_<N>.qua, read by utils/transform_utils.py:27-50) into the
binary data assets megapose6d_b200/data/so3_grid_<N>.npy (float64 [N,4], xyzw).

Only runs where /root/reference exists.  The grids are data produced by the public SO(3) sampling
code of Yershova et al. (http://lavalle.pl/software/so3/so3.html); identical hypotheses are needed
for a drop-in coarse stage.
"""
from pathlib import Path

import numpy as np

SRC = Path("/root/reference/src/megapose/data")
DST = Path(__file__).resolve().parents[1] / "megapose6d_b200" / "data"

if __name__ == "__main__":
    DST.mkdir(parents=True, exist_ok=True)
    for n in (72, 576, 4608):
        rows = [[float(v) for v in line.split()] for line in (SRC / f"data_{n}.qua").read_text().splitlines() if line.strip()]
        q = np.asarray(rows, dtype=np.float64)
        assert q.shape == (n, 4), q.shape
        np.save(DST / f"so3_grid_{n}.npy", q)
        print(n, q.shape, np.abs(np.linalg.norm(q, axis=1) - 1).max())
