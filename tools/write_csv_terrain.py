#!/usr/bin/env python
"""Writes a golden terrain (tests/golden/terrain_<name>.npz) in the reference's CSV directory layout
(data/<name>/{x,y,z,dx,dy,dz}data.csv, rows = y, columns = x) so that gbp_plan / gbp_terrain_create_csv can load it on a
box without /root/reference.   python tools/write_csv_terrain.py rough_terrain /tmp/rough"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main(name, out):
    d = np.load(os.path.join(ROOT, "tests", "golden", f"terrain_{name}.npz"))
    x, y = d["x"].astype(np.float64), d["y"].astype(np.float64)
    os.makedirs(out, exist_ok=True)
    X, Y = np.meshgrid(x, y)
    lay = {"x": X, "y": Y}
    for k in ("z", "dx", "dy", "dz"):
        lay[k] = d[k].astype(np.float64).T
    for k, v in lay.items():
        with open(os.path.join(out, f"{k}data.csv"), "w") as f:
            for row in v:
                f.write(",".join("nan" if np.isnan(c) else repr(float(c)) for c in row) + "\n")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
