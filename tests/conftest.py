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


from oracle.fixtures import ext2_forces, ext3_forces  # noqa: E402,F401  (the force lists of the lattice8_ext2 / _ext3 fixtures)


def dna3_special_types(g, hb_multiplier=1.7):
    """The oxDNA3 fixture with base types outside 0..3: A-T pairs of the first duplex become the custom pair -300 / 303 (hydrogen bonding
    times hb_multiplier), two nucleotides become dummy bases (btype = type = 4).  Returns (btype, scalars)."""
    import numpy as np
    bt = np.array(g["btype"]).copy()
    for i in range(20):
        j = 39 - i
        if (bt[i], bt[j]) == (3, 0):
            bt[i], bt[j] = 303, -300
        elif (bt[i], bt[j]) == (0, 3):
            bt[i], bt[j] = -300, 303
    bt[45], bt[130] = 4, 4
    sc = np.array(g["dna3_scalars"]).copy()
    sc[4] = hb_multiplier
    return bt, sc
