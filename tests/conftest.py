import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLD = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    # keep the in-tree CUDA library in step with its sources (no-op when fresh; nvcc cross-compiles without a GPU)
    from oxdna_b200 import build as _b
    _b.build()


def load_golden(name):
    z = np.load(os.path.join(GOLD, name + ".npz"), allow_pickle=False)
    return {k: z[k] for k in z.files}


@pytest.fixture(scope="session")
def golden():
    return load_golden


def pair_set(pairs):
    p = np.sort(np.asarray(pairs, dtype=np.int64), axis=1)
    return set(map(tuple, p.tolist()))


def ext2_forces(pos):
    """the force list of the lattice8_ext2 fixture (same construction as oracle/make_golden.py:ext2_forces)"""
    xmin, zmin = float(pos[:, 0].min()), float(pos[:, 2].min())
    return [dict(type="repulsion_plane", particle="all", stiff=1.5, dir=(0.0, 0.0, 1.0), position=-(zmin + 1.5), v=0.002, end_position=-(zmin + 1.6)),
            dict(type="attraction_plane", particle=17, stiff=0.3, dir=(0.0, 1.0, 0.0), position=-3.0),
            dict(type="attraction_plane", particle=200, stiff=0.3, dir=(0.0, 1.0, 0.0), position=-30.0),
            dict(type="sphere", particle="all", stiff=2.0, r0=7.0, rate=-0.001, center=(10.0, 10.0, 10.0)),
            dict(type="sphere", particle=5, stiff=1.0, r0=0.5, r_ext=3.0, center=(1.0, 19.0, 2.0)),
            dict(type="LJ_wall", particle="all", stiff=0.5, dir=(1.0, 0.0, 0.0), position=-(xmin - 1.0), sigma=1.0, n=6, only_repulsive=1),
            dict(type="lowdim_trap", particle=33, stiff=0.7, rate=0.001, pos0=(5.0, 5.0, 5.0), dir=(1.0, 1.0, 0.0), visibility=(1, 0, 1)),
            dict(type="mutual_trap", particle=0, ref_particle=39, stiff=0.1, r0=1.2, PBC=1)]
