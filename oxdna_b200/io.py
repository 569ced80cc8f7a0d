"""oxDNA topology / configuration file formats (docs/source/configurations.md:14-77 of the reference).

Host-side helpers used by the tests, the benchmark and the Python REMD driver.  The C++ host layer
(oxdna_b200/host) has its own reader; both follow the same rules as the reference's parser
(src/Interactions/DNAInteraction.cpp:1506-1577, src/Backends/SimBackend.cpp:545-660).
"""
import numpy as np

BASE_TO_BTYPE = {"A": 0, "G": 1, "C": 2, "T": 3, "U": 3, "D": 4}
BTYPE_TO_BASE = {0: "A", 1: "G", 2: "C", 3: "T", 4: "D"}  # D: the dummy base (Utils::decode_base, src/Utilities/Utils.cpp:25-38)


def read_topology(path):
    """Returns dict(N, n_strands, btype, n3, n5, strand) -- int32 arrays; -1 marks a missing neighbour."""
    with open(path) as f:
        lines = [l.strip() for l in f if l.strip()]
    head = lines[0].split()
    N, ns = int(head[0]), int(head[1])
    btype = np.zeros(N, dtype=np.int32)
    n3 = np.full(N, -1, dtype=np.int32)
    n5 = np.full(N, -1, dtype=np.int32)
    strand = np.zeros(N, dtype=np.int32)
    if len(head) > 2 and head[2] == "5->3":
        idx = 0
        for s, line in enumerate(lines[1:]):
            parts = line.split()
            seq = parts[0]
            opts = dict(p.split("=") for p in parts[1:] if "=" in p)
            circular = opts.get("circular", "false").lower() in ("true", "1", "yes")
            bts = []
            k = 0
            while k < len(seq):
                if seq[k] == "(":
                    j = seq.index(")", k)
                    bts.append(int(seq[k + 1:j]))
                    k = j + 1
                else:
                    bts.append(BASE_TO_BTYPE[seq[k].upper()])
                    k += 1
            L = len(bts)
            for i, b in enumerate(bts):
                btype[idx + i] = b
                strand[idx + i] = s
                if i > 0:
                    n5[idx + i] = idx + i - 1
                if i < L - 1:
                    n3[idx + i] = idx + i + 1
            if circular:
                n3[idx + L - 1] = idx
                n5[idx] = idx + L - 1
            idx += L
        assert idx == N, "topology: particle count mismatch"
    else:
        for i, line in enumerate(lines[1:1 + N]):
            sid, base, a, b = line.split()[:4]
            strand[i] = int(sid) - 1
            try:
                btype[i] = BASE_TO_BTYPE[base.upper()]
            except KeyError:
                btype[i] = int(base)
            n3[i], n5[i] = int(a), int(b)
    return dict(N=N, n_strands=ns, btype=btype, n3=n3, n5=n5, strand=strand)


def write_topology(path, btype, n3, n5, strand):
    N = len(btype)
    ns = int(strand.max()) + 1
    with open(path, "w") as f:
        f.write(f"{N} {ns}\n")
        for i in range(N):
            b = BTYPE_TO_BASE.get(int(btype[i]), str(int(btype[i])))
            f.write(f"{int(strand[i]) + 1} {b} {int(n3[i])} {int(n5[i])}\n")


def read_conf(path, N=None):
    """Returns dict(step, box, E, pos, a1, a3, vel, L) (float64)."""
    with open(path) as f:
        t = int(float(f.readline().split("=")[1]))
        box = np.array([float(x) for x in f.readline().split("=")[1].split()])
        E = np.array([float(x) for x in f.readline().split("=")[1].split()])
        data = np.loadtxt(f, ndmin=2, max_rows=N)
    if data.shape[1] == 9:
        data = np.hstack([data, np.zeros((data.shape[0], 6))])
    return dict(step=t, box=box, E=E, pos=data[:, 0:3].copy(), a1=data[:, 3:6].copy(), a3=data[:, 6:9].copy(),
                vel=data[:, 9:12].copy(), L=data[:, 12:15].copy())


def write_conf(path, box, pos, a1, a3, vel=None, L=None, step=0, E=(0.0, 0.0, 0.0)):
    N = pos.shape[0]
    vel = np.zeros((N, 3)) if vel is None else vel
    L = np.zeros((N, 3)) if L is None else L
    data = np.hstack([pos, a1, a3, vel, L])
    with open(path, "w") as f:
        f.write(f"t = {int(step)}\n")
        f.write("b = %.17g %.17g %.17g\n" % tuple(box))
        f.write("E = %.10g %.10g %.10g\n" % tuple(E))
        np.savetxt(f, data, fmt="%.17g")


def orthonormal_axes(a1, a3):
    """Same orthonormalisation as the reference's configuration reader (SimBackend.cpp:623-629).
    Returns (N, 9): a1, a2, a3."""
    a1 = a1 / np.linalg.norm(a1, axis=1, keepdims=True)
    a3 = a3 / np.linalg.norm(a3, axis=1, keepdims=True)
    a1 = a1 - a3 * np.sum(a1 * a3, axis=1, keepdims=True)
    a1 = a1 / np.linalg.norm(a1, axis=1, keepdims=True)
    a2 = np.cross(a3, a1)
    a2 = a2 / np.linalg.norm(a2, axis=1, keepdims=True)
    return np.hstack([a1, a2, a3])


def quaternions_from_axes(axes):
    """Rotation matrix (columns a1, a2, a3) -> unit quaternion (x, y, z, w); same branch rule as the reference's
    host-side marshalling (src/CUDA/Backends/MD_CUDABackend.cu:275-307)."""
    N = axes.shape[0]
    m = np.stack([axes[:, 0:3], axes[:, 3:6], axes[:, 6:9]], axis=2)  # m[:, r, c]: column c = axis c
    q = np.zeros((N, 4))
    tr = m[:, 0, 0] + m[:, 1, 1] + m[:, 2, 2]
    for i in range(N):
        M = m[i]
        if tr[i] > 0:
            s = 0.5 / np.sqrt(tr[i] + 1.0)
            q[i] = [(M[2, 1] - M[1, 2]) * s, (M[0, 2] - M[2, 0]) * s, (M[1, 0] - M[0, 1]) * s, 0.25 / s]
        elif M[0, 0] > M[1, 1] and M[0, 0] > M[2, 2]:
            s = 0.5 / np.sqrt(1.0 + M[0, 0] - M[1, 1] - M[2, 2])
            q[i] = [0.25 / s, (M[0, 1] + M[1, 0]) * s, (M[0, 2] + M[2, 0]) * s, (M[2, 1] - M[1, 2]) * s]
        elif M[1, 1] > M[2, 2]:
            s = 0.5 / np.sqrt(1.0 + M[1, 1] - M[0, 0] - M[2, 2])
            q[i] = [(M[0, 1] + M[1, 0]) * s, 0.25 / s, (M[1, 2] + M[2, 1]) * s, (M[0, 2] - M[2, 0]) * s]
        else:
            s = 0.5 / np.sqrt(1.0 + M[2, 2] - M[0, 0] - M[1, 1])
            q[i] = [(M[0, 2] + M[2, 0]) * s, (M[1, 2] + M[2, 1]) * s, 0.25 / s, (M[1, 0] - M[0, 1]) * s]
    return q


def axes_from_quaternions(q):
    """Inverse of the above (same expansion as src/CUDA/cuda_utils/CUDA_lr_common.cuh:40-61)."""
    x, y, z, w = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    a1 = np.stack([x * x - y * y - z * z + w * w, 2 * (x * y + z * w), 2 * (x * z - y * w)], axis=1)
    a2 = np.stack([2 * (x * y - z * w), -x * x + y * y - z * z + w * w, 2 * (y * z + x * w)], axis=1)
    a3 = np.stack([2 * (x * z + y * w), 2 * (y * z - x * w), -x * x - y * y + z * z + w * w], axis=1)
    return np.hstack([a1, a2, a3])


def read_binary_conf(path, N, frame=0):
    """one frame of the reference's binary configuration format (src/Observables/Configurations/BinaryConfiguration.cpp:20-92)"""
    rec = np.dtype([("pos", "<f8", 3), ("shift", "<i4", 3), ("a1", "<f8", 3), ("a2", "<f8", 3), ("a3", "<f8", 3), ("vel", "<f8", 3), ("L", "<f8", 3)])
    head = np.dtype([("step", "<i8"), ("rng", "<u2", 3), ("box", "<f8", 3), ("E", "<f8", 3)])
    size = head.itemsize + N * rec.itemsize
    with open(path, "rb") as f:
        f.seek(frame * size)
        h = np.frombuffer(f.read(head.itemsize), dtype=head)[0]
        p = np.frombuffer(f.read(N * rec.itemsize), dtype=rec)
    out = {k: np.array(p[k]) for k in rec.names}
    out.update(step=int(h["step"]), rng=np.array(h["rng"]), box=np.array(h["box"]), E=np.array(h["E"]))
    return out
