"""TEST/BENCH INFRASTRUCTURE (generator script, runs only where /root/reference is mounted).

Fits the rigid "next nucleotide" and "paired nucleotide" transforms of an oxRNA A-form duplex from the reference's
example configuration examples/RNA_DUPLEX_MELT/init.conf (8 bp) and prints them as the constants used by
oxdna_b200/lattice.py:rna_duplex_lattice.  Frames: columns (a1, a2, a3) of nucleotide i; a transform (R, d) maps the frame
of nucleotide i to the frame of its neighbour: R_j = R_i R, x_j = x_i + R_i d.

    python oracle/fit_rna_helix.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oxdna_b200 import io as oio  # noqa: E402


def frames(c):
    a1, a3 = c["a1"], c["a3"]
    a2 = np.cross(a3, a1)
    return np.stack([a1, a2, a3], axis=2)  # (N, 3, 3), columns = axes


def mean_rotation(Rs):
    U, _, Vt = np.linalg.svd(np.mean(Rs, axis=0))
    R = U @ Vt
    if np.linalg.det(R) < 0:
        U[:, -1] *= -1
        R = U @ Vt
    return R


def main():
    ex = "/root/reference/examples/RNA_DUPLEX_MELT"
    c = oio.read_conf(os.path.join(ex, "init.conf"))
    R, x = frames(c), c["pos"]
    n = 8
    step_R, step_d, pair_R, pair_d = [], [], [], []
    for i in range(1, n - 2):  # interior steps of strand 0 (index increases 3' -> 5')
        step_R.append(R[i].T @ R[i + 1])
        step_d.append(R[i].T @ (x[i + 1] - x[i]))
    for i in range(1, n - 1):
        p = 2 * n - 1 - i
        pair_R.append(R[i].T @ R[p])
        d = x[p] - x[i]
        d -= c["box"] * np.rint(d / c["box"])
        pair_d.append(R[i].T @ d)
    np.set_printoptions(precision=12, suppress=True)
    print("RNA_STEP_R =", repr(mean_rotation(np.array(step_R))))
    print("RNA_STEP_D =", repr(np.mean(step_d, axis=0)))
    print("RNA_PAIR_R =", repr(mean_rotation(np.array(pair_R))))
    print("RNA_PAIR_D =", repr(np.mean(pair_d, axis=0)))
    print("# spread of the step translation:", np.std(step_d, axis=0))


if __name__ == "__main__":
    main()
