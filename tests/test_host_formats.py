"""Host-side data formats and generators either side of the hot path (no GPU): the reference's topology / configuration text
formats (docs/source/configurations.md:14-34, old and `5->3` topologies), the binary configuration record
(src/Observables/Configurations/BinaryConfiguration.cpp:20-92), the particle-list strings of the forces file
(Utils::get_particles_from_string) and the synthetic lattices of BASELINE.json's configs."""
import os
import struct

import numpy as np

from oxdna_b200 import capi, io as oio, lattice
from oxdna_b200.remd import geometric_ladder
from oxdna_b200.sim import brownian_params, langevin_params, parse_temperature

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_topology_and_configuration_round_trip(tmp_path):
    s = lattice.duplex_lattice(3, bp=5, spacing=10.0, seed=2)
    top, conf = str(tmp_path / "t.top"), str(tmp_path / "c.dat")
    oio.write_topology(top, s["btype"], s["n3"], s["n5"], s["strand"])
    t = oio.read_topology(top)
    assert t["N"] == 30 and t["n_strands"] == 6
    for k in ("btype", "n3", "n5", "strand"):
        assert np.array_equal(t[k], s[k]), k
    rng = np.random.default_rng(0)
    vel, L = rng.normal(size=(30, 3)), rng.normal(size=(30, 3))
    oio.write_conf(conf, s["box"], s["pos"], s["a1"], s["a3"], vel, L, step=12, E=(1.0, -2.0, 3.0))
    c = oio.read_conf(conf)
    assert c["step"] == 12 and np.array_equal(c["box"], s["box"])
    for k, ref in (("pos", s["pos"]), ("a1", s["a1"]), ("a3", s["a3"]), ("vel", vel), ("L", L)):
        assert np.array_equal(c[k], ref), k  # %.17g is lossless
    # a frame without momenta (9 columns) reads with zero velocities
    with open(conf, "w") as f:
        f.write("t = 0\nb = 30 30 30\nE = 0 0 0\n")
        np.savetxt(f, np.hstack([s["pos"], s["a1"], s["a3"]]))
    assert not oio.read_conf(conf)["vel"].any()


def test_new_style_topology_matches_the_old_one(tmp_path):
    """the `5->3` topology of the reference (one line per strand, sequence 5'->3', optional circular=true) against the old format: in
    the old format nucleotide i of a strand has n3 = i - 1 (3'->5' listing), in the new one the listing runs the other way"""
    p = tmp_path / "new.top"
    p.write_text("7 2 5->3\nACGT type=DNA\nGG(-17) circular=true\n")
    t = oio.read_topology(str(p))
    assert t["N"] == 7 and list(t["strand"]) == [0, 0, 0, 0, 1, 1, 1]
    assert list(t["btype"]) == [0, 2, 1, 3, 1, 1, -17]
    assert list(t["n5"][:4]) == [-1, 0, 1, 2] and list(t["n3"][:4]) == [1, 2, 3, -1]
    assert t["n3"][6] == 4 and t["n5"][4] == 6  # the circular strand closes on itself


def test_binary_configuration_reader_follows_the_reference_record(tmp_path):
    N = 3
    p = tmp_path / "f.bin"
    rng = np.random.default_rng(1)
    vals = rng.normal(size=(N, 18))
    with open(p, "wb") as f:
        f.write(struct.pack("<q3H3d3d", 77, 5, 6, 7, 10.0, 11.0, 12.0, 0.5, -1.5, 2.0))
        for i in range(N):
            f.write(struct.pack("<3d3i15d", *vals[i, :3], i, -i, 2 * i, *vals[i, 3:]))
    b = oio.read_binary_conf(str(p), N)
    assert b["step"] == 77 and list(b["rng"]) == [5, 6, 7] and list(b["box"]) == [10.0, 11.0, 12.0] and list(b["E"]) == [0.5, -1.5, 2.0]
    assert np.array_equal(b["pos"], vals[:, :3]) and np.array_equal(b["a1"], vals[:, 3:6]) and np.array_equal(b["a2"], vals[:, 6:9])
    assert np.array_equal(b["a3"], vals[:, 9:12]) and np.array_equal(b["vel"], vals[:, 12:15]) and np.array_equal(b["L"], vals[:, 15:18])
    assert np.array_equal(b["shift"], [[0, 0, 0], [1, -1, 2], [2, -2, 4]])


def test_particle_list_strings():
    f = capi._index_list
    assert f(5) == [5] and f("7") == [7] and f("1,2, 9") == [1, 2, 9] and f("3-6") == [3, 4, 5, 6] and f("0,4-5") == [0, 4, 5]
    assert f([2, 3]) == [2, 3] and f(np.int32(4)) == [4]


def test_lattice_generators_give_bonded_complementary_duplexes():
    for gen, bp in ((lattice.duplex_lattice, 20), (lattice.rna_duplex_lattice, 16)):
        s = gen(27, bp=bp, spacing=10.0, seed=3)
        N = 27 * 2 * bp
        assert len(s["pos"]) == N and np.allclose(s["box"], 30.0)
        assert np.allclose(np.linalg.norm(s["a1"], axis=1), 1) and np.allclose(np.linalg.norm(s["a3"], axis=1), 1)
        assert np.abs(np.einsum("ij,ij->i", s["a1"], s["a3"])).max() < 1e-10  # the A-form helix parameters are fitted numbers
        # every strand is one chain: n3 / n5 are mutually consistent and stay inside the strand
        for i in range(N):
            if s["n3"][i] >= 0:
                assert s["n5"][s["n3"][i]] == i and s["strand"][s["n3"][i]] == s["strand"][i]
        assert (s["n3"] < 0).sum() == 54 and (s["n5"] < 0).sum() == 54
        # Watson-Crick partners: nucleotide k of strand 2d pairs with nucleotide bp-1-k of strand 2d+1 (btype sum 3)
        d0 = np.arange(bp)
        assert np.all(s["btype"][d0] + s["btype"][2 * bp - 1 - d0] == 3)
        # bonded neighbours sit within the FENE range of the backbone spring (|r_bb - r0| < Delta = 0.25); centre distance as a proxy
        j = s["n3"][s["n3"] >= 0]
        i = np.flatnonzero(s["n3"] >= 0)
        dist = np.linalg.norm(s["pos"][i] - s["pos"][j], axis=1)
        assert dist.max() < 1.0 and dist.min() > 0.3
    traps = lattice.mutual_traps(dict(bp=20, n_duplex=2))
    assert len(traps) == 4 and {(t["particle"], t["ref_particle"]) for t in traps} == {(0, 39), (39, 0), (40, 79), (79, 40)}


def test_input_value_parsing_follows_the_reference():
    # src/Utilities/Utils.cpp:316-346
    assert abs(parse_temperature("300K") - 0.1) < 1e-15 and abs(parse_temperature("27C") - 300.15 * 0.1 / 300) < 1e-15
    assert parse_temperature("0.11") == 0.11 and parse_temperature(0.2) == 0.2
    pt, pr, resc = brownian_params(0.1, 0.003, 103, diff_coeff=2.5)
    assert 0 < pr < pt < 1 and abs(resc - np.sqrt(0.1)) < 1e-15
    g, gr, rt, rr = langevin_params(0.1, 0.003, diff_coeff=2.5)
    assert abs(g - 0.1 / 2.5) < 1e-7 and abs(gr - 0.1 / 7.5) < 1e-7 and rt > 0 and rr > 0
    lad = geometric_ladder(290.0, 350.0, 64)
    assert lad[0] == 290.0 and abs(lad[-1] - 350.0) < 1e-9 and np.allclose(lad[1:] / lad[:-1], lad[1] / lad[0])
