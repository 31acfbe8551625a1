"""Helpers shared by the golden-vector tests."""
import glob
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
RECORD = (1, 2, 50, 100, 200)


def golden_cases():
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz"))
                  if not os.path.basename(p).startswith(("smoke", "variant", "reset")))


def load(name):
    g = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    n, cen, cr, dt, flk = g["meta"]
    return g, dict(n=int(n), centralized=bool(cen), comm_radius=float(cr), dt=float(dt), knn=bool(flk))


def unpack_adj(packed, n):
    return np.unpackbits(packed, axis=1)[:, :n].astype(np.float64)
