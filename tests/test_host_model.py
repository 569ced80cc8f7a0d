"""CPU test of the product's FP32 pair-potential arithmetic (oxdna_b200/csrc/dna_model.cuh compiled for the host by nvcc)
against the double-precision oracle.  Tolerance: the north star's mixed-precision bound, |dF| <= 1e-5 * max|F|."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, load_golden
from oracle import oracle as O
from oxdna_b200 import capi
from oxdna_b200.sim import parse_temperature

SO = os.path.join(ROOT, "tests", "support", "_build", "libhostmodel.so")


@pytest.fixture(scope="module")
def hostlib():
    os.makedirs(os.path.dirname(SO), exist_ok=True)
    src = os.path.join(ROOT, "tests", "support", "host_model.cu")
    deps = [src, os.path.join(ROOT, "oxdna_b200", "csrc", "dna_model.cuh"), os.path.join(ROOT, "oxdna_b200", "csrc", "common.cuh")]
    if not os.path.exists(SO) or any(os.path.getmtime(d) > os.path.getmtime(SO) for d in deps):
        subprocess.check_call(["nvcc", "-O2", "-std=c++17", "-Wno-deprecated-gpu-targets", "-Xcompiler", "-fPIC", "-shared", "-o", SO, src,
                               os.path.join(ROOT, "oxdna_b200", "csrc", "params.cpp")])
    return C.CDLL(SO)


@pytest.mark.parametrize("case", ["force_field_dna/ref_dna2_nomesh", "lattice8", "lattice27_dense"])
def test_fp32_formulation_within_mixed_tolerance(hostlib, case):
    g = load_golden(case)
    T, salt = parse_temperature(str(g["T"])), float(g["salt"])
    M, rcut = capi.dna2_params(T, salt)
    N = len(g["pos"])
    ax = np.ascontiguousarray(O.axes_from_a1a3(g["a1"], g["a3"]))
    pairs = np.ascontiguousarray(g["pairs"], dtype=np.int32)
    F, Tl, ep = np.zeros((N, 3)), np.zeros((N, 3)), np.zeros(N)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    pos, box = np.ascontiguousarray(g["pos"]), np.ascontiguousarray(g["box"], dtype=np.float64)
    bt, n3, n5 = (np.ascontiguousarray(g[k], dtype=np.int32) for k in ("btype", "n3", "n5"))
    hostlib.host_dna2_forces(C.byref(M), N, p(pos), p(ax), p(bt), p(n3), p(n5), p(box), p(pairs), C.c_longlong(len(pairs)), p(F), p(Tl), p(ep))
    fmax = np.linalg.norm(g["force"], axis=1).max()
    tmax = np.linalg.norm(g["torque_lab"], axis=1).max()
    assert np.linalg.norm(F - g["force"], axis=1).max() <= 1e-5 * fmax
    assert np.linalg.norm(Tl - g["torque_lab"], axis=1).max() <= 1e-5 * tmax
    assert abs(ep.sum() - float(g["U"])) <= 1e-6 * abs(float(g["U"]))
