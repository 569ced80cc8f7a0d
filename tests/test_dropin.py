"""Drop-in test of the C++ host side: the reference's own main()/SimManager/input parser linked with OUR BackendFactory,
MD_CUDABackend and operator classes (oxdna_b200/host) runs a stock input file with `backend = CUDA`, and the trajectory
is compared with the unmodified reference CPU binary (oracle/_ref/oxDNA, `backend = CPU`, analytic DNA2_nomesh potential)
started from the same files and seed.  Config C1 of BASELINE.json (examples/HAIRPIN, 18 nt, max_backbone_force = 10).

Both executables are build artefacts of this container (they link the reference's objects) and travel to the GPU box;
where they are missing the test is skipped.  Tolerances: FP32 pair arithmetic vs FP64 over 300 NVE steps of a chaotic
system -> positions 2e-3, CPU-evaluated energies of the two trajectories 2e-4 per nucleotide."""
import os
import shutil
import subprocess

import numpy as np
import pytest

from oxdna_b200 import io as oio

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OURS = os.path.join(ROOT, "oxdna_b200", "host", "_build", "oxDNA_b200")
REF = os.path.join(ROOT, "oracle", "_ref", "oxDNA")
FIX = os.path.join(ROOT, "tests", "golden", "hairpin")

INPUT = """backend = {backend}
backend_precision = mixed
sim_type = MD
interaction_type = {itype}
salt_concentration = 0.5
T = 334K
dt = 0.003
steps = {steps}
thermostat = {thermostat}
newtonian_steps = 103
diff_coeff = 2.5
verlet_skin = 0.05
max_backbone_force = 10
CUDA_list = verlet
CUDA_sort_every = {sort_every}
use_edge = {use_edge}
seed = 42
refresh_vel = 1
topology = initial.top
conf_file = initial.conf
trajectory_file = trajectory.dat
lastconf_file = last_conf.dat
energy_file = energy.dat
print_energy_every = 100
print_conf_interval = 100000
restart_step_counter = 1
time_scale = linear
no_stdout_energy = 1
{extra}
"""

FORCES = """{
type = mutual_trap
particle = 0
ref_particle = 17
stiff = 0.5
r0 = 1.5
PBC = 1
}
{
type = mutual_trap
particle = 17
ref_particle = 0
stiff = 0.5
r0 = 1.5
PBC = 1
}
"""


FORCES2 = """{
type = repulsion_plane
particle = all
stiff = 1.0
dir = 0,0,1
position = -23.5
}
{
type = sphere
particle = all
stiff = 2.0
r0 = 4.0
center = 25.,25.,25.
}
{
type = LJ_wall
particle = 3
stiff = 0.4
dir = 1,0,0
position = -20.0
sigma = 1.5
n = 4
}
{
type = lowdim_trap
particle = 9
stiff = 0.5
rate = 0.001
pos0 = 25., 25., 25.
dir = 0., 1., 0.
visibility = 0, 1, 1
}
{
type = attraction_plane
particle = 0
stiff = 0.2
dir = 0,1,0
position = -20.
}
{
type = twist
particle = 17
stiff = 0.3
rate = 0.001
base = 0.2
axis = 0., 0., 1.
pos0 = 8., 13., 36.
center = 7., 12., 36.
mask = 1., 1., 0.
}
{
type = ellipsoid
particle = all
stiff = 0.05
r_2 = 3., 4., 3.
center = 8., 13., 36.
}
"""


FORCES3 = """{
type = repulsion_plane_moving
particle = all
ref_particle = 9
stiff = 0.5
dir = 1,0,0
}
{
type = generic_central_force
particle = all
center = 8.9,13.9,34.8
force_type = gravity
F0 = 0.02
inner_cut_off = 0.5
}
{
type = LJ_cone
particle = all
stiff = 2.0
sigma = 4.0
alpha = 0.5
n = 6
dir = 0,0,1
pos0 = 8.9,13.9,12.0
}
{
type = com
com_list = 0,1,2
ref_list = 15,16,17
stiff = 0.3
r0 = 1.0
}
{
type = yukawa_sphere
particle = all
radius = 9.0
center = 8.9,13.9,34.8
debye_length = 2.0
debye_A = 1.0
}
{
type = repulsive_sphere_moving
particle = all
stiff = 0.5
r0 = 3.8
rate = 0.0005
origin = 8.9,13.9,28.0
target = 8.9,13.9,28.3
steps = 200
}
"""


OP_FILE = "{\norder_parameter = bond\nname = stem\n" + "".join(f"pair{k + 1} = {k}, {17 - k}\n" for k in range(6)) + "}\n"
FORCES4 = """{
type = meta_coordination
op_file = op.txt
coordination_type = mixed
mixed_weight = 0.7
hb_energy_cutoff = -0.1
d0 = 0.4
r0 = 0.5
n = 6
coord_min = 0.0
coord_max = 6.06
N_grid = 31
potential_grid = """ + ",".join("%.8f" % (0.5 * (x - 3.0) ** 2) for x in np.linspace(0.0, 6.06, 31)) + """
}
"""


def run(binary, d, fix=FIX, files=("initial.top", "initial.conf"), **kw):
    os.makedirs(d, exist_ok=True)
    for f, name in zip(files, ("initial.top", "initial.conf")):
        shutil.copy(os.path.join(fix, f), os.path.join(d, name))
    with open(os.path.join(d, "forces.txt"), "w") as f:
        f.write(FORCES)
    with open(os.path.join(d, "forces2.txt"), "w") as f:
        f.write(FORCES2)
    with open(os.path.join(d, "forces3.txt"), "w") as f:
        f.write(FORCES3)
    with open(os.path.join(d, "forces4.txt"), "w") as f:
        f.write(FORCES4)
    with open(os.path.join(d, "op.txt"), "w") as f:
        f.write(OP_FILE)
    with open(os.path.join(d, "input"), "w") as f:
        f.write(INPUT.format(**kw))
    p = subprocess.run([binary, "input"], cwd=d, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    return p


def energies(d):
    return np.loadtxt(os.path.join(d, "energy.dat"), ndmin=2)


needs_binaries = pytest.mark.skipif(not (os.path.exists(OURS) and os.path.exists(REF)), reason="drop-in executables not built (need /root/reference)")


@pytest.mark.gpu
@needs_binaries
@pytest.mark.parametrize("use_edge,sort_every,extra", [(1, 1, ""), (0, 0, ""), (1, 1, "external_forces = 1\nexternal_forces_file = forces.txt"),
                                                       (1, 1, "external_forces = 1\nexternal_forces_file = forces2.txt"),
                                                       (1, 1, "external_forces = 1\nexternal_forces_file = forces3.txt"),
                                                       (1, 1, "fix_diffusion_every = 100"),
                                                       (1, 1, "external_forces = 1\nexternal_forces_file = forces4.txt")])
def test_stock_input_file_matches_reference_cpu(tmp_path, use_edge, sort_every, extra):
    a = run(OURS, str(tmp_path / "ours"), backend="CUDA", itype="DNA2", steps=300, thermostat="no", use_edge=use_edge, sort_every=sort_every, extra=extra)
    assert a.returncode == 0, a.stdout[-2000:]
    b = run(REF, str(tmp_path / "ref"), backend="CPU", itype="DNA2_nomesh", steps=300, thermostat="no", use_edge=0, sort_every=0, extra=extra)
    assert b.returncode == 0, b.stdout[-2000:]
    ca, cb = oio.read_conf(str(tmp_path / "ours" / "last_conf.dat")), oio.read_conf(str(tmp_path / "ref" / "last_conf.dat"))
    assert np.abs(ca["pos"] - cb["pos"]).max() < 2e-3
    assert np.abs(ca["a1"] - cb["a1"]).max() < 5e-3
    ea, eb = energies(str(tmp_path / "ours")), energies(str(tmp_path / "ref"))
    assert ea.shape == eb.shape and ea.shape[0] >= 3
    assert np.abs(ea[:, 1:] - eb[:, 1:]).max() < 2e-4


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(os.path.join(os.path.dirname(OURS), "plugins", "CUDAToy.so")), reason="toy plugin not built (needs /root/reference)")
def test_third_party_interaction_through_the_plugin_manager(tmp_path):
    """The plugin seam of the drop-in backend: `interaction_type = Toy` is unknown to both factories, so the reference's
    InteractionFactory loads Toy.so (make_Toy) and our CUDAInteractionFactory loads CUDAToy.so (make_CUDAToy) through the reference's
    own PluginManager (src/CUDA/Interactions/CUDAInteractionFactory.cu:44-51, src/PluginManagement/PluginManager.cpp:89-180).  The toy
    interaction (tests/plugin) computes a soft repulsion with its own kernel from the raw device arrays (poss float4, column-major
    matrix_neighs, number_neighs); 300 NVE steps are compared with a plain numpy velocity-Verlet run of the same potential."""
    from conftest import load_golden
    from test_gpu_plugin_seam import toy_reference
    g = load_golden("lattice8")
    src = tmp_path / "src"
    os.makedirs(src)
    oio.write_topology(str(src / "t.top"), g["btype"], g["n3"], g["n5"], g["strand"])
    oio.write_conf(str(src / "t.conf"), g["box"], g["pos"], g["a1"], g["a3"], g["vel"], g["L"])
    plug = os.path.join(os.path.dirname(OURS), "plugins")
    d = str(tmp_path / "toy")
    a = run(OURS, d, fix=str(src), files=("t.top", "t.conf"), backend="CUDA", itype="Toy", steps=300, thermostat="no", use_edge=0, sort_every=1,
            extra=f"plugin_search_path = {plug}\nrefresh_vel = 0\ntoy_k = 3.0\ntoy_rc = 1.4\n")
    assert a.returncode == 0, a.stdout[-2000:]
    last = oio.read_conf(os.path.join(d, "last_conf.dat"))
    p_ref, v_ref, _ = toy_reference(g["pos"], g["vel"], g["n3"], g["n5"], g["box"], 0.003, 300)
    assert np.abs(last["pos"] - p_ref).max() < 2e-5 and np.abs(last["vel"] - v_ref).max() < 2e-5
    assert np.abs(last["pos"] - g["pos"]).max() > 1e-2  # it did move
    # an interaction nobody provides fails the way the reference fails
    bad = run(OURS, str(tmp_path / "none"), fix=str(src), files=("t.top", "t.conf"), backend="CUDA", itype="DNA2", steps=10, thermostat="no", use_edge=0,
              sort_every=0, extra=f"plugin_search_path = {plug}\ninteraction_type = Nope\n")
    assert bad.returncode != 0


@pytest.mark.gpu
@needs_binaries
def test_stock_input_file_thermostat_and_errors(tmp_path):
    a = run(OURS, str(tmp_path / "t"), backend="CUDA", itype="DNA2", steps=5000, thermostat="brownian", use_edge=1, sort_every=1, extra="")
    assert a.returncode == 0, a.stdout[-2000:]
    e = energies(str(tmp_path / "t"))
    assert np.all(np.isfinite(e)) and e.shape[0] >= 40
    # the strained start structure (max_backbone_force) dumps ~6 units of energy per nucleotide into kinetic energy; the
    # Andersen-like thermostat (pt = 0.0137 every 103 steps -> relaxation time ~7500 steps) must be draining it
    assert e[-5:, 2].mean() < 0.8 * e[:5, 2].mean()
    # reference incompatibilities are reported as the reference reports them (fatal oxDNAException, non-zero exit)
    bad = run(OURS, str(tmp_path / "bad"), backend="CUDA", itype="DNA2", steps=10, thermostat="no", use_edge=1, sort_every=0, extra="CUDA_list = no")
    assert bad.returncode != 0 and "incompatible" in bad.stdout
    # CUDA_list = no (src/CUDA/Lists/CUDANoList.cu) without use_edge is accepted: same forces as the Verlet list, so the NVE energies of a
    # short run are identical to the CUDA_list = verlet run (the later key overrides the template's)
    nl = run(OURS, str(tmp_path / "nolist"), backend="CUDA", itype="DNA2", steps=300, thermostat="no", use_edge=0, sort_every=0, extra="CUDA_list = no")
    vl = run(OURS, str(tmp_path / "verlet"), backend="CUDA", itype="DNA2", steps=300, thermostat="no", use_edge=0, sort_every=0, extra="")
    assert nl.returncode == 0 and vl.returncode == 0, nl.stdout[-1500:]
    assert np.array_equal(energies(str(tmp_path / "nolist")), energies(str(tmp_path / "verlet")))
    bad = run(OURS, str(tmp_path / "bad2"), backend="CUDA", itype="DNA2", steps=10, thermostat="no", use_edge=1, sort_every=0, extra="reload_from = x")
    assert bad.returncode != 0
    bad = run(OURS, str(tmp_path / "bad3"), backend="CUDA", itype="LJ", steps=10, thermostat="no", use_edge=1, sort_every=0, extra="")
    assert bad.returncode != 0 and "CUDA interaction 'CUDALJ' not found" in bad.stdout  # the reference's own message (CUDAInteractionFactory.cu:49)


@pytest.mark.gpu
@needs_binaries
@pytest.mark.parametrize("use_edge,sort_every,extra", [(1, 1, ""), (0, 0, "use_average_seq = 0\nseq_dep_file = seq.txt\nmismatch_repulsion = 1")])
def test_stock_rna_input_file_matches_reference_cpu(tmp_path, use_edge, sort_every, extra):
    """interaction_type = RNA2 through the drop-in executable (CUDARNAInteraction on the reference's RNA2Interaction), on the
    16-nt system of the reference's test/RNA/FORCE_FIELD, against the reference CPU binary.  The CPU class interpolates the
    hydrogen-bonding factors on 6-12 point meshes and its force deviates from the gradient in two places (oracle/oxdna_oracle.h),
    so the trajectories part faster than for DNA: 100 steps, looser bounds."""
    from oxdna_b200 import seqdep
    rfix = os.path.join(ROOT, "tests", "golden", "force_field_rna")
    kw = dict(fix=rfix, files=("init.top", "init.dat"), steps=100, thermostat="no", extra=extra)
    for d in ("ours", "ref"):
        os.makedirs(str(tmp_path / d), exist_ok=True)
        seqdep.write_file(str(tmp_path / d / "seq.txt"), seqdep.RNA_SEQ_DEP)
    a = run(OURS, str(tmp_path / "ours"), backend="CUDA", itype="RNA2", use_edge=use_edge, sort_every=sort_every, **kw)
    assert a.returncode == 0, a.stdout[-2000:]
    b = run(REF, str(tmp_path / "ref"), backend="CPU", itype="RNA2", use_edge=0, sort_every=0, **kw)
    assert b.returncode == 0, b.stdout[-2000:]
    ca, cb = oio.read_conf(str(tmp_path / "ours" / "last_conf.dat")), oio.read_conf(str(tmp_path / "ref" / "last_conf.dat"))
    assert np.abs(ca["pos"] - cb["pos"]).max() < 5e-3
    assert np.abs(ca["a1"] - cb["a1"]).max() < 2e-2
    ea, eb = energies(str(tmp_path / "ours")), energies(str(tmp_path / "ref"))
    assert ea.shape == eb.shape
    assert np.abs(ea[:, 1:] - eb[:, 1:]).max() < 1e-3


@needs_binaries
def test_no_cpu_fallback_in_the_dropin_binary(tmp_path):
    """Without a CUDA device the backend must refuse to run (this test is meaningful on the CPU-only container)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    a = run(OURS, str(tmp_path / "nogpu"), backend="CUDA", itype="DNA2", steps=10, thermostat="no", use_edge=1, sort_every=0, extra="")
    assert a.returncode != 0 and "no CPU fallback" in a.stdout


@pytest.mark.gpu
@needs_binaries
def test_stock_input_file_npt_matches_reference_cpu(tmp_path):
    """use_barostat = 1.  Both executables draw the per-step activation, the box change and the acceptance number
    from drand48 in the same order, so as long as the Metropolis decisions agree the boxes follow the same sequence: final box side
    and acceptance ratio are compared with the reference CPU backend (whose VolumeMove acts at mid-step; ours between steps)."""
    extra = "use_barostat = 1\nP = 0.01\ndelta_L = 0.5\nbarostat_probability = 0.2"
    a = run(OURS, str(tmp_path / "ours"), backend="CUDA", itype="DNA2", steps=300, thermostat="no", use_edge=0, sort_every=1, extra=extra)
    assert a.returncode == 0, a.stdout[-2000:]
    b = run(REF, str(tmp_path / "ref"), backend="CPU", itype="DNA2_nomesh", steps=300, thermostat="no", use_edge=0, sort_every=0, extra=extra)
    assert b.returncode == 0, b.stdout[-2000:]
    ca, cb = oio.read_conf(str(tmp_path / "ours" / "last_conf.dat")), oio.read_conf(str(tmp_path / "ref" / "last_conf.dat"))
    assert cb["box"][0] < 49.0
    assert abs(ca["box"][0] - cb["box"][0]) < 0.3, (ca["box"], cb["box"])
    ea, eb = energies(str(tmp_path / "ours")), energies(str(tmp_path / "ref"))
    assert ea.shape == eb.shape and ea.shape[1] == 6           # time, U, K, E, density, barostat acceptance
    assert abs(ea[-1, 5] - eb[-1, 5]) < 0.1
    assert abs(ea[-1, 4] / eb[-1, 4] - 1.0) < 0.03


@pytest.mark.gpu
@needs_binaries
def test_device_side_energy_stream_matches_reference_cpu(tmp_path):
    """CUDA_device_observables = 1: the default energy streams print U, K, U + K evaluated on the device and the print
    skips the download / CPU list rebuild / upload bracket of SimBackend::print_observables.  Same columns, same values (FP32 pair
    arithmetic vs FP64) as the reference CPU binary; the trajectory is unaffected."""
    a = run(OURS, str(tmp_path / "ours"), backend="CUDA", itype="DNA2", steps=300, thermostat="no", use_edge=1, sort_every=1, extra="CUDA_device_observables = 1")
    assert a.returncode == 0, a.stdout[-2000:]
    b = run(REF, str(tmp_path / "ref"), backend="CPU", itype="DNA2_nomesh", steps=300, thermostat="no", use_edge=0, sort_every=0, extra="")
    assert b.returncode == 0, b.stdout[-2000:]
    ea, eb = energies(str(tmp_path / "ours")), energies(str(tmp_path / "ref"))
    assert ea.shape == eb.shape and ea.shape[0] >= 3
    assert np.abs(ea[:, 1:] - eb[:, 1:]).max() < 2e-4
    ca, cb = oio.read_conf(str(tmp_path / "ours" / "last_conf.dat")), oio.read_conf(str(tmp_path / "ref" / "last_conf.dat"))
    assert np.abs(ca["pos"] - cb["pos"]).max() < 2e-3


QUICK_INPUT = """backend = CUDA
backend_precision = mixed
CUDA_list = verlet
CUDA_sort_every = 0
use_edge = {use_edge}
CUDA_device_observables = {dev_obs}
steps = {steps}
newtonian_steps = 103
diff_coeff = 2.50
thermostat = john
T = {T}
dt = 0.005
verlet_skin = 0.05
topology = {top}
conf_file = {conf}
trajectory_file = trajectory.dat
refresh_vel = 1
log_file = log.dat
no_stdout_energy = 1
restart_step_counter = 1
energy_file = energy.dat
print_conf_interval = 1e5
print_energy_every = 1e3
time_scale = linear
external_forces = 0
seed = 20261017
"""


@pytest.mark.gpu
@needs_binaries
@pytest.mark.parametrize("name,T,checks,use_edge,dev_obs,steps", [
    ("dsdna8", "20C", [(2, -1.37970256144, 0.15)], 1, 0, 400000),
    ("ssdna15", "300K", [(2, -0.700783241758, 0.26), (3, 0.30, 0.015)], 0, 1, 1000000)])
def test_reference_quick_md_tests_on_the_gpu_backend(tmp_path, name, T, checks, use_edge, dev_obs, steps):
    """The reference's OWN statistical MD tests (test/DNA/DSDNA8/MD, test/DNA/SSDNA15/MD: quick_input + quick_compare, ColumnAverage of
    test/TestSuite.py:141-196) with `backend = CUDA`: stock input keys, first-generation oxDNA, john thermostat, the expected column
    averages and tolerances are the reference's.  The duplex is shortened from 1e6 to 4e5 steps (its tolerance is tens of standard errors
    wide); the 15-nt single strand folds and unfolds transient hairpins, so it keeps the reference's 1e6 steps (at 4e5 steps and a
    time-based seed one run in five or so lands outside the reference's tolerance) and both runs carry a fixed seed."""
    gold = os.path.join(ROOT, "tests", "golden", "quick_md")
    d = str(tmp_path)
    shutil.copy(os.path.join(gold, name + ".top"), os.path.join(d, name + ".top"))
    shutil.copy(os.path.join(gold, name + "_init.dat"), os.path.join(d, "init.dat"))
    with open(os.path.join(d, "input"), "w") as f:
        f.write(QUICK_INPUT.format(use_edge=use_edge, dev_obs=dev_obs, steps=steps, T=T, top=name + ".top", conf="init.dat"))
    p = subprocess.run([OURS, "input"], cwd=d, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-2000:]
    e = energies(d)
    assert e.shape[0] >= steps // 1000
    for col, want, tol in checks:
        avg = e[:, col - 1].mean()
        assert want - tol <= avg <= want + tol, (name, col, avg, want, tol)


SEQ3 = "/root/reference/oxDNA3_sequence_dependent_parameters.txt"


@pytest.mark.gpu
@needs_binaries
@pytest.mark.parametrize("use_edge,sort_every,seqdep", [(0, 0, False), (1, 1, False), (1, 1, True)])
def test_stock_input_file_oxdna3_matches_reference_cpu(tmp_path, use_edge, sort_every, seqdep):
    """interaction_type = DNA3 through the drop-in executable (CUDADNA3Interaction on the reference's DNA3Interaction_nomesh: the tables the
    CPU class derives are handed to oxb_set_model_dna3) against the reference CPU binary with DNA3_nomesh: 300 NVE steps from the thermalised
    oxDNA3 fixture (8 duplexes).  Average-sequence tables everywhere; the sequence-dependent parameter file only where /root/reference is
    mounted."""
    if seqdep and not os.path.exists(SEQ3):
        pytest.skip("oxDNA3 parameter file not available (needs /root/reference)")
    g = np.load(os.path.join(ROOT, "tests", "golden", "dna3_lattice8.npz"))
    fix = str(tmp_path / "fix")
    os.makedirs(fix)
    oio.write_topology(os.path.join(fix, "initial.top"), g["btype"], g["n3"], g["n5"], g["strand"])
    oio.write_conf(os.path.join(fix, "initial.conf"), g["box"], g["pos"], g["a1"], g["a3"], g["vel"], g["L"])
    extra = "refresh_vel = 0\n" + (f"use_average_seq = 0\nseq_dep_file = {SEQ3}" if seqdep else "")
    a = run(OURS, str(tmp_path / "ours"), fix=fix, backend="CUDA", itype="DNA3", steps=300, thermostat="no", use_edge=use_edge, sort_every=sort_every, extra=extra)
    assert a.returncode == 0, a.stdout[-2000:]
    b = run(REF, str(tmp_path / "ref"), fix=fix, backend="CPU", itype="DNA3_nomesh", steps=300, thermostat="no", use_edge=0, sort_every=0, extra=extra)
    assert b.returncode == 0, b.stdout[-2000:]
    ca, cb = oio.read_conf(str(tmp_path / "ours" / "last_conf.dat")), oio.read_conf(str(tmp_path / "ref" / "last_conf.dat"))
    assert np.abs(ca["pos"] - cb["pos"]).max() < 2e-3
    assert np.abs(ca["a1"] - cb["a1"]).max() < 5e-3
    ea, eb = energies(str(tmp_path / "ours")), energies(str(tmp_path / "ref"))
    assert ea.shape == eb.shape and ea.shape[0] >= 3
    assert np.abs(ea[:, 1:] - eb[:, 1:]).max() < 2e-4
