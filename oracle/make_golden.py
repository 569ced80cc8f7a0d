"""TEST INFRASTRUCTURE ONLY.  Generates the committed fixtures under tests/golden/ by running the UNMODIFIED reference
CPU implementation (oracle/_ref/liboxref.so, built by oracle/Makefile.ref from /root/reference).  Only runnable in the
build container; the GPU box uses the committed .npz files.

    python oracle/make_golden.py
"""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle.refharness import Reference  # noqa: E402
from oracle import refharness as RH  # noqa: E402
from oxdna_b200 import io as oio, lattice  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def dump(ref, topo, path, extra=None):
    RH.lib().oxref_rebuild_lists()
    st = ref.state()
    out = ref.compute_forces()
    d = dict(pos=st["pos"], a1=st["a1"], a3=st["a3"], vel=st["vel"], L=st["L"], box=ref.box(), rcut=ref.rcut(),
             btype=topo["btype"], n3=topo["n3"], n5=topo["n5"], strand=topo["strand"],
             force=out["force"], torque_body=out["torque_body"], torque_lab=out["torque_lab"], U=out["U"],
             energy_split=ref.energy_split(), pairs=ref.pairs())
    if extra:
        d.update(extra)
    np.savez_compressed(path, **d)
    print("wrote", path, "N =", ref.N, "U/N =", out["U"] / ref.N, "pairs =", len(d["pairs"]))


def force_field():
    top = os.path.join(GOLD, "force_field_dna", "init.top")
    conf = os.path.join(GOLD, "force_field_dna", "init.dat")
    r = Reference(top, conf, interaction_type="DNA2_nomesh", salt_concentration=1.0, T="20C")
    dump(r, r.topology(), os.path.join(GOLD, "force_field_dna", "ref_dna2_nomesh.npz"), dict(T="20C", salt=1.0))
    r.close()


def lattice_case(name, n_duplex, spacing, steps, T="300K", salt=0.5, seed=3, nve_steps=200, ext=None, itype="DNA2_nomesh", more_keys=None, dna3=False):
    sysm = lattice.duplex_lattice(n_duplex, bp=20, spacing=spacing, seed=seed)
    d = tempfile.mkdtemp()
    top, conf = os.path.join(d, "l.top"), os.path.join(d, "l.dat")
    oio.write_topology(top, sysm["btype"], sysm["n3"], sysm["n5"], sysm["strand"])
    from oxdna_b200.sim import parse_temperature
    v, L = lattice.maxwell_velocities(len(sysm["pos"]), parse_temperature(T), 5)
    oio.write_conf(conf, sysm["box"], sysm["pos"], sysm["a1"], sysm["a3"], v, L)
    r = Reference(top, conf, interaction_type=itype, salt_concentration=salt, T=T, thermostat="brownian",
                  newtonian_steps=103, diff_coeff=2.5, seed=7, **(more_keys or {}))
    r.step(steps)
    st = r.state()
    topo = r.topology()
    r.close()
    # restart without thermostat from the thermalised state: forces + an NVE segment
    conf2 = os.path.join(d, "t.dat")
    oio.write_conf(conf2, sysm["box"], st["pos"], st["a1"], st["a3"], st["vel"], st["L"])
    keys = dict(interaction_type=itype, salt_concentration=salt, T=T, thermostat="no", dt=0.003, **(more_keys or {}))
    if ext:
        fpath = os.path.join(d, "forces.txt")
        with open(fpath, "w") as f:
            for e in ext:
                f.write("{\n")
                for k, val in e.items():
                    if isinstance(val, (tuple, list)):
                        val = ",".join(str(x) for x in val)
                    f.write(f"{k} = {val}\n")
                f.write("}\n")
        keys.update(external_forces=1, external_forces_file=fpath)
    r = Reference(top, conf2, **keys)
    st0 = r.state()
    RH.lib().oxref_rebuild_lists()
    f0 = r.compute_forces()
    base = dict(pos=st0["pos"], a1=st0["a1"], a3=st0["a3"], vel=st0["vel"], L=st0["L"], box=r.box(), rcut=r.rcut(),
                btype=topo["btype"], n3=topo["n3"], n5=topo["n5"], strand=topo["strand"], force=f0["force"],
                torque_body=f0["torque_body"], torque_lab=f0["torque_lab"], U=f0["U"], energy_split=r.energy_split(), pairs=r.pairs(),
                T=T, salt=salt)
    if dna3:
        # the tetramer-indexed tables + scalars the live DNA3Interaction holds after init() (what CUDADNA3Interaction::cuda_init uploads)
        tab, sc = np.zeros((215, 900)), np.zeros(40)
        k = RH.lib().oxref_dna3_tables(RH._p(tab), RH._p(sc))
        assert k == 29
        base.update(dna3_tables=tab, dna3_scalars=sc[:k])
    r.step(nve_steps)
    st1 = r.state()
    base.update(nve_steps=nve_steps, pos1=st1["pos"], a11=st1["a1"], a31=st1["a3"], vel1=st1["vel"], L1=st1["L"],
                U1=r.system_energy(), n_updates=r.n_updates())
    r.close()
    path = os.path.join(GOLD, name + ".npz")
    np.savez_compressed(path, **base)
    print("wrote", path, "N =", len(st0["pos"]), "U/N =", f0["U"] / len(st0["pos"]), "pairs =", len(base["pairs"]))


SEQDEP_RNA = "/root/reference/rna_sequence_dependent_parameters.txt"


def read_seq_dep_rna(path=SEQDEP_RNA):
    vals = {}
    for line in open(path):
        if "=" in line:
            k, v = line.split("=")
            vals[k.strip()] = float(v)
    B = "AGCT"
    return dict(stck=np.array([vals[f"STCK_{a}_{b}"] for a in B for b in B]), cross=np.array([vals[f"CROSS_{a}_{b}"] for a in B for b in B]),
                st_t_dep=vals["ST_T_DEP"], hb_AT=vals["HYDR_A_T"], hb_GC=vals["HYDR_C_G"], hb_GT=vals["HYDR_G_T"])


def force_field_rna():
    """test/RNA/FORCE_FIELD of the reference: the 16-nt configuration, its golden per-term energies, and the full CPU output."""
    import shutil
    src = "/root/reference/test/RNA/FORCE_FIELD"
    dst = os.path.join(GOLD, "force_field_rna")
    os.makedirs(dst, exist_ok=True)
    for f in ("init.top", "init.dat"):
        shutil.copy(os.path.join(src, f), os.path.join(dst, f))
    shutil.copy(os.path.join(src, "AVG_SEQ", "reference.dat"), os.path.join(dst, "reference_avg_seq.dat"))
    top, conf = os.path.join(dst, "init.top"), os.path.join(dst, "init.dat")
    r = Reference(top, conf, interaction_type="RNA2", salt_concentration=1.0, T="20C")
    dump(r, r.topology(), os.path.join(dst, "ref_rna2.npz"), dict(T="20C", salt=1.0))
    r.close()
    r = Reference(top, conf, interaction_type="RNA2", salt_concentration=1.0, T="20C", use_average_seq=0, seq_dep_file=SEQDEP_RNA)
    sd = read_seq_dep_rna()
    dump(r, r.topology(), os.path.join(dst, "ref_rna2_seqdep.npz"), dict(T="20C", salt=1.0, **{"sd_" + k: v for k, v in sd.items()}))
    r.close()


def rna_lattice_case(name, n_duplex, steps, T="300K", salt=0.5, seed=3, nve_steps=100, no_hb=False, seq_dep=False, mismatch=None):
    """A-form RNA duplex lattice thermalised by the reference CPU backend (RNA2), then forces + an NVE segment from the
    thermalised state.  no_hb: same state re-evaluated with an all-A sequence (no hydrogen bonding => no meshed term, so
    forces and trajectories pin the restatement to rounding)."""
    sysm = lattice.rna_duplex_lattice(n_duplex, bp=16, spacing=10.0, seed=seed)
    d = tempfile.mkdtemp()
    top, conf = os.path.join(d, "l.top"), os.path.join(d, "l.dat")
    oio.write_topology(top, sysm["btype"], sysm["n3"], sysm["n5"], sysm["strand"])
    from oxdna_b200.sim import parse_temperature
    v, L = lattice.maxwell_velocities(len(sysm["pos"]), parse_temperature(T), 5)
    oio.write_conf(conf, sysm["box"], sysm["pos"], sysm["a1"], sysm["a3"], v, L)
    keys = dict(interaction_type="RNA2", salt_concentration=salt, T=T)
    if seq_dep:
        keys.update(use_average_seq=0, seq_dep_file=SEQDEP_RNA)
    if mismatch is not None:
        keys.update(mismatch_repulsion=1, mismatch_repulsion_strength=mismatch)
    r = Reference(top, conf, thermostat="brownian", newtonian_steps=103, diff_coeff=2.5, seed=7, **keys)
    r.step(steps)
    st = r.state()
    r.close()
    if no_hb:
        oio.write_topology(top, np.zeros_like(sysm["btype"]), sysm["n3"], sysm["n5"], sysm["strand"])
    conf2 = os.path.join(d, "t.dat")
    oio.write_conf(conf2, sysm["box"], st["pos"], st["a1"], st["a3"], st["vel"], st["L"])
    r = Reference(top, conf2, thermostat="no", dt=0.003, **keys)
    topo = r.topology()
    st0 = r.state()
    RH.lib().oxref_rebuild_lists()
    f0 = r.compute_forces()
    base = dict(pos=st0["pos"], a1=st0["a1"], a3=st0["a3"], vel=st0["vel"], L=st0["L"], box=r.box(), rcut=r.rcut(),
                btype=topo["btype"], n3=topo["n3"], n5=topo["n5"], strand=topo["strand"], force=f0["force"],
                torque_body=f0["torque_body"], torque_lab=f0["torque_lab"], U=f0["U"], energy_split=r.energy_split(), pairs=r.pairs(),
                T=T, salt=salt, seq_dep=int(seq_dep), mismatch=-1.0 if mismatch is None else float(mismatch))
    if seq_dep:
        base.update({"sd_" + k: v for k, v in read_seq_dep_rna().items()})
    r.step(nve_steps)
    st1 = r.state()
    base.update(nve_steps=nve_steps, pos1=st1["pos"], a11=st1["a1"], a31=st1["a3"], vel1=st1["vel"], L1=st1["L"],
                U1=r.system_energy(), n_updates=r.n_updates())
    r.close()
    path = os.path.join(GOLD, name + ".npz")
    np.savez_compressed(path, **base)
    print("wrote", path, "N =", len(st0["pos"]), "U/N =", f0["U"] / len(st0["pos"]), "pairs =", len(base["pairs"]),
          "terms/N", np.round(base["energy_split"] / len(st0["pos"]), 4))


from oracle.fixtures import ext2_forces, ext3_forces  # noqa: E402


def ext2_case(name="lattice8_ext2", force_list=ext2_forces):
    """lattice8 state + forces file with the further force types, evaluated and stepped by the reference CPU backend"""
    g = dict(np.load(os.path.join(GOLD, "lattice8.npz")))
    d = tempfile.mkdtemp()
    top, conf = os.path.join(d, "l.top"), os.path.join(d, "l.dat")
    oio.write_topology(top, g["btype"], g["n3"], g["n5"], g["strand"])
    oio.write_conf(conf, g["box"], g["pos"], g["a1"], g["a3"], g["vel"], g["L"])
    ext = force_list(g["pos"])
    fpath = os.path.join(d, "forces.txt")
    with open(fpath, "w") as f:
        for e in ext:
            f.write("{\n")
            for k, val in e.items():
                if isinstance(val, (tuple, list)):
                    val = ",".join(str(x) for x in val)
                f.write(f"{k} = {val}\n")
            f.write("}\n")
    r = Reference(top, conf, interaction_type="DNA2_nomesh", salt_concentration=float(g["salt"]), T=str(g["T"]), thermostat="no", dt=0.003,
                  external_forces=1, external_forces_file=fpath)
    st0 = r.state()
    RH.lib().oxref_rebuild_lists()
    f0 = r.compute_forces()
    base = {k: g[k] for k in ("btype", "n3", "n5", "strand", "box", "T", "salt")}
    base.update(pos=st0["pos"], a1=st0["a1"], a3=st0["a3"], vel=st0["vel"], L=st0["L"], rcut=r.rcut(), force=f0["force"], torque_lab=f0["torque_lab"],
                force_noext=g["force"], U=f0["U"])
    n = 100
    r.step(n)
    st1 = r.state()
    base.update(nve_steps=n, pos1=st1["pos"], a11=st1["a1"], vel1=st1["vel"], L1=st1["L"])
    r.close()
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), **base)
    dF = np.abs(base["force"] - base["force_noext"])
    print("wrote", name, ": particles feeling an external force:", int((dF.max(axis=1) > 1e-12).sum()), "max |dF|", dF.max())


def rna_quirks_case(name="rna_quirks"):
    """A configuration on which the reference CPU class's force is NOT the gradient of its energy: the mirrored coaxial theta1 term
    enters with the opposite sign (RNAInteraction.cpp:1046 vs the mesh builder :1302) -- the reference's CUDA kernels and ours use the
    gradient (CUDA_RNA.cuh:896).  (The other such spot, the phi2 stacking term without the theta-B factors, RNAInteraction.cpp:620, is
    dormant with the stock parameters: inside the theta-B2 window, |theta_B2| < 0.9615 rad around -p5 = (0.104, 0.842, -0.530), the
    backbone direction has a2 . bhat >= 0.04, where f5(phi2) = 1 and its derivative vanishes; 60,000 random draws confirm it.)
    All-A sequence (no hydrogen bonding: no meshed factor); the reference's 16-nt RNA test system (a nicked duplex: the nick is coaxially
    stacked) with every nucleotide rotated about its own backbone site (the FENE bonds keep their lengths, the angles move a lot).  Out
    of 40,000 seeded draws with moderate forces the one on which the term changes the torques most is kept."""
    from oracle import oracle as O
    d = tempfile.mkdtemp()
    top = os.path.join(d, "aaaa.top")
    with open(top, "w") as f:
        f.write("16 3 5->3\nAAAA circular=False type=RNA\nAAAA circular=False type=RNA\nAAAAAAAA circular=False type=RNA\n")
    r = Reference(top, os.path.join(GOLD, "force_field_rna", "init.dat"), interaction_type="RNA2", salt_concentration=0.3, T="37C")
    st, topo = r.state(), r.topology()
    rng = np.random.default_rng(5)
    P = [O.rna2_params(O.celsius(37.0), 0.3, cpu_quirks=q) for q in (0, 1, 2)]

    def rot(axis, ang):
        axis = axis / np.linalg.norm(axis)
        K = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
        return np.eye(3) + np.sin(ang) * K + (1 - np.cos(ang)) * (K @ K)

    best, phi2_seen = None, 0.0
    for it in range(40000):
        pos, a1, a3 = st["pos"].copy(), st["a1"].copy(), st["a3"].copy()
        for i in range(len(pos)):
            R = rot(rng.normal(size=3), rng.normal(scale=[0.3, 0.5, 0.8][it % 3]))
            back = pos[i] - 0.4 * a1[i] + 0.2 * a3[i]  # oxRNA backbone site (rna_model.h: RNA_POS_BACK_a1, _a3)
            a1[i], a3[i] = R @ a1[i], R @ a3[i]
            pos[i] = back + 0.4 * a1[i] - 0.2 * a3[i]
        ax = O.axes_from_a1a3(a1, a3)
        pairs = O.verlet_pairs(pos, topo["n3"], topo["n5"], r.box(), P[0].rcut + 0.1)
        o = [O.forces(p, pos, ax, topo["btype"], topo["n3"], topo["n5"], r.box(), pairs) for p in P]
        phi2_seen = max(phi2_seen, np.abs(o[1]["torque_lab"] - o[0]["torque_lab"]).max(), np.abs(o[1]["force"] - o[0]["force"]).max())
        # moderate forces and intact bonds (no FENE blow-up) so that the 1e-5 criterion is meaningful
        if not (abs(o[0]["U"]) < 1e3 and np.linalg.norm(o[0]["force"], axis=1).max() < 150):
            continue
        sc = np.linalg.norm(o[2]["torque_lab"] - o[0]["torque_lab"], axis=1).max() / np.linalg.norm(o[0]["torque_lab"], axis=1).max()
        if best is None or sc > best[0]:
            best = (sc, pos, ax)
    score, pos, ax = best
    print("relative torque change by the mirrored theta1 term on the kept configuration: %g; largest change by the phi2 term over all draws: %g" % (score, phi2_seen))
    r.set_state(pos, ax[:, 0:3], ax[:, 6:9])
    dump(r, topo, os.path.join(GOLD, name + ".npz"), dict(T="37C", salt=0.3))
    r.close()


def rna():
    force_field_rna()
    rna_lattice_case("rna_lattice8", 8, 3000)
    rna_lattice_case("rna_lattice8_nohb", 8, 3000, no_hb=True)
    rna_lattice_case("rna_lattice8_seqdep", 8, 3000, T="310K", salt=1.0, seq_dep=True, mismatch=1.0)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "rna":
        rna()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "rna_quirks":
        rna_quirks_case()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "dna1":
        lattice_case("lattice8_dna1", 8, 8.0, 3000, T="310K", itype="DNA_nomesh", nve_steps=100)
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "dna3":
        # oxDNA3 (DNA3Interaction_nomesh: the class the CUDA backend instantiates, InteractionFactory.cpp:63-65), sequence-dependent tables
        # FIRST, while the process is fresh: with use_average_seq = 1 the reference fills the stacking F1_SD_EPS from a temperature that is
        # not set yet and never assigns F1_SD_SHIFT (DNA3Interaction.cpp:82-83): the tables are what the heap holds -- zeros in a new
        # process (what the reference's CLI and its CUDA backend see), leftovers of an earlier interaction object otherwise
        # the average-sequence tables (use_average_seq = 1: no parameter file) at the bench's temperature and salt: bench.py --workload c2_dna3
        sysm = lattice.duplex_lattice(8, bp=20, spacing=10.0, seed=3)
        d = tempfile.mkdtemp()
        top, conf = os.path.join(d, "l.top"), os.path.join(d, "l.dat")
        oio.write_topology(top, sysm["btype"], sysm["n3"], sysm["n5"], sysm["strand"])
        v, L = lattice.maxwell_velocities(len(sysm["pos"]), 0.1, 5)
        oio.write_conf(conf, sysm["box"], sysm["pos"], sysm["a1"], sysm["a3"], v, L)
        r = Reference(top, conf, interaction_type="DNA3_nomesh", salt_concentration=0.5, T="300K")
        tab, sc = np.zeros((215, 900)), np.zeros(40)
        k = RH.lib().oxref_dna3_tables(RH._p(tab), RH._p(sc))
        r.close()
        np.savez_compressed(os.path.join(GOLD, "dna3_tables_avg_300K_salt05.npz"), dna3_tables=tab, dna3_scalars=sc[:k], T="300K", salt=0.5)
        sd = dict(use_average_seq=0, seq_dep_file="/root/reference/oxDNA3_sequence_dependent_parameters.txt")
        lattice_case("dna3_lattice8", 8, 10.0, 3000, itype="DNA3_nomesh", more_keys=sd, dna3=True, nve_steps=100)
        lattice_case("dna3_lattice27_dense", 27, 8.5, 4000, T="330K", salt=0.2, itype="DNA3_nomesh", more_keys=sd, dna3=True, nve_steps=100)
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "ext2":
        ext2_case()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "ext3":
        ext2_case("lattice8_ext3", ext3_forces)
        sys.exit(0)
    force_field()
    lattice_case("lattice8", 8, 10.0, 3000)
    lattice_case("lattice27_dense", 27, 8.5, 4000, T="330K")
    ext = [dict(type="mutual_trap", particle=0, ref_particle=39, stiff=0.1, r0=1.2, PBC=1),
           dict(type="mutual_trap", particle=39, ref_particle=0, stiff=0.1, r0=1.2, PBC=1),
           dict(type="trap", particle=45, pos0=(5.0, 5.0, 5.0), stiff=0.5, rate=0.001, dir=(1.0, 0.0, 0.0)),
           dict(type="string", particle=80, F0=0.2, rate=0.0001, dir=(0.0, 1.0, 1.0))]
    lattice_case("lattice8_ext", 8, 10.0, 2000, ext=ext, nve_steps=100)
