"""Synthetic duplex-lattice systems (BASELINE.json configs C2 / C4 / C5).

Ideal B-helix geometry follows the recipe of the reference's generator (utils/generate-sa.py:160-225):
rise 0.3897628551303122 per base pair, twist 35.9 degrees, nucleotide centre of mass 0.6 from the helix axis
along -a1.  Duplex axes lie along z, centres on a simple-cubic lattice, one random azimuth per duplex.
"""
import numpy as np

RISE = 0.3897628551303122
TWIST = np.deg2rad(35.9)
CM_CENTER_DS = 0.6


def duplex_lattice(n_duplex, bp=20, spacing=10.0, seed=12345, sites_per_side=None):
    """Returns dict(box, pos, a1, a3, btype, n3, n5, strand).  Strand 2d is the 'top' strand of duplex d (a3 = +z),
    strand 2d+1 its complement.  Old-style (3'->5') topology: nucleotide i of a strand has n3 = i-1, n5 = i+1."""
    rng = np.random.default_rng(seed)
    if sites_per_side is None:
        sites_per_side = int(np.ceil(n_duplex ** (1.0 / 3.0) - 1e-9))
    S = sites_per_side
    assert S ** 3 >= n_duplex
    L = S * spacing
    N = n_duplex * 2 * bp

    seq = rng.integers(0, 4, size=(n_duplex, bp))
    azim = rng.uniform(0, 2 * np.pi, size=n_duplex)
    site = np.arange(n_duplex)
    centre = np.stack([site % S, (site // S) % S, site // (S * S)], axis=1).astype(np.float64) * spacing + 0.5 * spacing

    k = np.arange(bp)
    ang = azim[:, None] + TWIST * k[None, :]                       # (D, bp)
    a1_top = np.stack([np.cos(ang), np.sin(ang), np.zeros_like(ang)], axis=2)
    z0 = -0.5 * (bp - 1) * RISE
    rb = centre[:, None, :] + np.stack([np.zeros_like(ang), np.zeros_like(ang), z0 + RISE * k[None, :] + 0 * ang], axis=2)
    pos_top = rb - CM_CENTER_DS * a1_top
    a3_top = np.broadcast_to(np.array([0.0, 0.0, 1.0]), a1_top.shape)
    # complementary strand: antiparallel, visits the base pairs in reverse order
    a1_bot = -a1_top[:, ::-1, :]
    pos_bot = rb[:, ::-1, :] - CM_CENTER_DS * a1_bot
    a3_bot = -a3_top
    seq_bot = 3 - seq[:, ::-1]

    pos = np.concatenate([pos_top, pos_bot], axis=1).reshape(N, 3)
    a1 = np.concatenate([a1_top, a1_bot], axis=1).reshape(N, 3)
    a3 = np.concatenate([a3_top, a3_bot], axis=1).reshape(N, 3)
    btype = np.concatenate([seq, seq_bot], axis=1).reshape(N).astype(np.int32)

    idx = np.arange(N)
    in_strand = idx % bp
    n3 = np.where(in_strand == 0, -1, idx - 1).astype(np.int32)
    n5 = np.where(in_strand == bp - 1, -1, idx + 1).astype(np.int32)
    strand = (idx // bp).astype(np.int32)
    return dict(box=np.array([L, L, L]), pos=pos, a1=np.ascontiguousarray(a1), a3=np.ascontiguousarray(a3),
                btype=btype, n3=n3, n5=n5, strand=strand, bp=bp, n_duplex=n_duplex)


# A-form oxRNA duplex: rigid transforms between the body frames (columns a1, a2, a3) of consecutive nucleotides of a strand
# (index increasing, i.e. 3' -> 5') and between a nucleotide and its Watson-Crick partner, R_j = R_i R, x_j = x_i + R_i d.
# Fitted by oracle/fit_rna_helix.py to the 8-bp duplex of the reference's examples/RNA_DUPLEX_MELT/init.conf.
RNA_STEP_R = np.array([[0.856848692155, -0.507194709428, 0.092541047541],
                       [0.515442346329, 0.846689337093, -0.132046787403],
                       [-0.011380086228, 0.160843691766, 0.986914282223]])
RNA_STEP_D = np.array([-0.128316409653, -0.367384926897, 0.42406827851])
RNA_PAIR_R = np.array([[-0.998108835305, 0.042447600826, -0.044462951657],
                       [0.050111106882, 0.980779445527, -0.188575067809],
                       [0.035603789868, -0.190446529022, -0.981051726328]])
RNA_PAIR_D = np.array([1.186844799649, -0.022931537307, -0.007604992795])


def rna_duplex(bp):
    """One ideal A-form duplex: (pos, a1, a3) of 2*bp nucleotides, strand A then strand B (antiparallel), helix axis along z
    through the origin, centred."""
    R, x = np.eye(3), np.zeros(3)
    RA, xA = [], []
    for _ in range(bp):
        RA.append(R)
        xA.append(x)
        x = x + R @ RNA_STEP_D
        R = R @ RNA_STEP_R
    RB = [RA[i] @ RNA_PAIR_R for i in range(bp - 1, -1, -1)]
    xB = [xA[i] + RA[i] @ RNA_PAIR_D for i in range(bp - 1, -1, -1)]
    Rs, xs = np.array(RA + RB), np.array(xA + xB)
    # helix axis = rotation axis of the step transform expressed in the lab frame (the same for every step: R_i u = u)
    w, v = np.linalg.eig(RNA_STEP_R)
    u = np.real(v[:, np.argmin(np.abs(w - 1.0))])
    u /= np.linalg.norm(u)
    if u @ RNA_STEP_D < 0:
        u = -u
    # rotate u onto z
    z = np.array([0.0, 0.0, 1.0])
    c, ax = u @ z, np.cross(u, z)
    K = np.array([[0, -ax[2], ax[1]], [ax[2], 0, -ax[0]], [-ax[1], ax[0], 0]])
    Q = np.eye(3) + K + K @ K / (1.0 + c)
    xs = xs @ Q.T
    Rs = np.einsum("ij,njk->nik", Q, Rs)
    # centre: the axis passes through the mean of the base-pair midpoints
    mid = 0.5 * (xs[:bp] + xs[bp:][::-1])
    xs = xs - mid.mean(axis=0)
    return xs, Rs[:, :, 0].copy(), Rs[:, :, 2].copy()


def rna_duplex_lattice(n_duplex, bp=16, spacing=10.0, seed=12345, sites_per_side=None):
    """BASELINE.json config C3: A-form RNA duplexes (random sequence, complement on the partner strand) on a simple-cubic
    lattice, axes along z, random azimuth.  Same topology conventions as duplex_lattice."""
    rng = np.random.default_rng(seed)
    if sites_per_side is None:
        sites_per_side = int(np.ceil(n_duplex ** (1.0 / 3.0) - 1e-9))
    S = sites_per_side
    assert S ** 3 >= n_duplex
    L = S * spacing
    N = n_duplex * 2 * bp
    x0, a10, a30 = rna_duplex(bp)
    seq = rng.integers(0, 4, size=(n_duplex, bp))
    azim = rng.uniform(0, 2 * np.pi, size=n_duplex)
    site = np.arange(n_duplex)
    centre = np.stack([site % S, (site // S) % S, site // (S * S)], axis=1).astype(np.float64) * spacing + 0.5 * spacing
    ca, sa = np.cos(azim), np.sin(azim)
    Rz = np.zeros((n_duplex, 3, 3))
    Rz[:, 0, 0], Rz[:, 0, 1], Rz[:, 1, 0], Rz[:, 1, 1], Rz[:, 2, 2] = ca, -sa, sa, ca, 1.0
    pos = np.einsum("dij,nj->dni", Rz, x0) + centre[:, None, :]
    a1 = np.einsum("dij,nj->dni", Rz, a10)
    a3 = np.einsum("dij,nj->dni", Rz, a30)
    btype = np.concatenate([seq, 3 - seq[:, ::-1]], axis=1).reshape(N).astype(np.int32)
    idx = np.arange(N)
    in_strand = idx % bp
    n3 = np.where(in_strand == 0, -1, idx - 1).astype(np.int32)
    n5 = np.where(in_strand == bp - 1, -1, idx + 1).astype(np.int32)
    strand = (idx // bp).astype(np.int32)
    return dict(box=np.array([L, L, L]), pos=pos.reshape(N, 3), a1=np.ascontiguousarray(a1.reshape(N, 3)),
                a3=np.ascontiguousarray(a3.reshape(N, 3)), btype=btype, n3=n3, n5=n5, strand=strand, bp=bp, n_duplex=n_duplex)


def mutual_traps(sys, stiff=0.1, r0=1.2, pbc=True):
    """C4's traps: for duplex d, first nt of strand 2d <-> last nt of strand 2d+1, both directions."""
    bp, D = sys["bp"], sys["n_duplex"]
    d = np.arange(D)
    a = 2 * d * bp
    b = (2 * d + 1) * bp + bp - 1
    out = []
    for p, q in zip(np.concatenate([a, b]), np.concatenate([b, a])):
        out.append(dict(type="mutual_trap", particle=int(p), ref_particle=int(q), stiff=stiff, r0=r0, PBC=int(pbc)))
    return out


def maxwell_velocities(N, T, seed=1):
    """MDBackend::_generate_vel (src/Backends/MDBackend.cpp:140-165): v, L ~ N(0, T) with unit mass and inertia."""
    rng = np.random.default_rng(seed)
    s = np.sqrt(T)
    return rng.normal(size=(N, 3)) * s, rng.normal(size=(N, 3)) * s
