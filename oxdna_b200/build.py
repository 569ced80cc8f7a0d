"""Builds liboxdna_b200.so (CUDA kernels + C ABI) in-tree with nvcc for sm_100a."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SO = os.path.join(HERE, "liboxdna_b200.so")
SOURCES = ["context.cu", "forces.cu", "forces_dna3.cu", "integrate.cu", "lists.cu", "sort.cu", "marshal.cu", "params.cpp"]
# extra -D flags for tuning experiments (profiles/micro/occupancy_sweep.sh)
EXTRA = os.environ.get("OXB_EXTRA_NVCC", "").split()
NVCC_FLAGS = EXTRA + ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-ftz=true", "-lineinfo", "-Xcompiler", "-fPIC",
              "-Xcompiler", "-O3", "--expt-relaxed-constexpr", "-Wno-deprecated-gpu-targets"]


def needs_build():
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "oxdna_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return SO
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    objs = []
    procs = []
    for src in SOURCES:
        obj = os.path.join(objdir, src + ".o")
        cmd = ["nvcc"] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, src), "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    ok = True
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            ok = False
            sys.stderr.write(f"--- {src}\n{out}\n")
        elif verbose:
            sys.stderr.write(f"--- {src}\n{out}\n")
    if not ok:
        raise RuntimeError("nvcc failed")
    subprocess.check_call(["nvcc", "-shared", "-o", SO] + objs + ["-lcudart"])
    return SO


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(SO)
