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


def lattice_case(name, n_duplex, spacing, steps, T="300K", salt=0.5, seed=3, nve_steps=200, ext=None):
    sysm = lattice.duplex_lattice(n_duplex, bp=20, spacing=spacing, seed=seed)
    d = tempfile.mkdtemp()
    top, conf = os.path.join(d, "l.top"), os.path.join(d, "l.dat")
    oio.write_topology(top, sysm["btype"], sysm["n3"], sysm["n5"], sysm["strand"])
    from oxdna_b200.sim import parse_temperature
    v, L = lattice.maxwell_velocities(len(sysm["pos"]), parse_temperature(T), 5)
    oio.write_conf(conf, sysm["box"], sysm["pos"], sysm["a1"], sysm["a3"], v, L)
    r = Reference(top, conf, interaction_type="DNA2_nomesh", salt_concentration=salt, T=T, thermostat="brownian",
                  newtonian_steps=103, diff_coeff=2.5, seed=7)
    r.step(steps)
    st = r.state()
    topo = r.topology()
    r.close()
    # restart without thermostat from the thermalised state: forces + an NVE segment
    conf2 = os.path.join(d, "t.dat")
    oio.write_conf(conf2, sysm["box"], st["pos"], st["a1"], st["a3"], st["vel"], st["L"])
    keys = dict(interaction_type="DNA2_nomesh", salt_concentration=salt, T=T, thermostat="no", dt=0.003)
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
    r.step(nve_steps)
    st1 = r.state()
    base.update(nve_steps=nve_steps, pos1=st1["pos"], a11=st1["a1"], a31=st1["a3"], vel1=st1["vel"], L1=st1["L"],
                U1=r.system_energy(), n_updates=r.n_updates())
    r.close()
    path = os.path.join(GOLD, name + ".npz")
    np.savez_compressed(path, **base)
    print("wrote", path, "N =", len(st0["pos"]), "U/N =", f0["U"] / len(st0["pos"]), "pairs =", len(base["pairs"]))


if __name__ == "__main__":
    force_field()
    lattice_case("lattice8", 8, 10.0, 3000)
    lattice_case("lattice27_dense", 27, 8.5, 4000, T="330K")
    ext = [dict(type="mutual_trap", particle=0, ref_particle=39, stiff=0.1, r0=1.2, PBC=1),
           dict(type="mutual_trap", particle=39, ref_particle=0, stiff=0.1, r0=1.2, PBC=1),
           dict(type="trap", particle=45, pos0=(5.0, 5.0, 5.0), stiff=0.5, rate=0.001, dir=(1.0, 0.0, 0.0)),
           dict(type="string", particle=80, F0=0.2, rate=0.0001, dir=(0.0, 1.0, 1.0))]
    lattice_case("lattice8_ext", 8, 10.0, 2000, ext=ext, nve_steps=100)
