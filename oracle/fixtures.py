"""TEST INFRASTRUCTURE ONLY.  Definitions shared by the fixture generator (oracle/make_golden.py) and the tests."""


def ext2_forces(pos):
    """The external-force list of the lattice8_ext2 fixture: the further force types, placed so that each
    one acts on the thermalised lattice8 state; keys are the reference's (docs/source/forces.md)."""
    xmin, zmin = float(pos[:, 0].min()), float(pos[:, 2].min())
    return [dict(type="repulsion_plane", particle="all", stiff=1.5, dir=(0.0, 0.0, 1.0), position=-(zmin + 1.5), v=0.002, end_position=-(zmin + 1.6)),
            dict(type="attraction_plane", particle=17, stiff=0.3, dir=(0.0, 1.0, 0.0), position=-3.0),
            dict(type="attraction_plane", particle=200, stiff=0.3, dir=(0.0, 1.0, 0.0), position=-30.0),
            dict(type="sphere", particle="all", stiff=2.0, r0=7.0, rate=-0.001, center=(10.0, 10.0, 10.0)),
            dict(type="sphere", particle=5, stiff=1.0, r0=0.5, r_ext=3.0, center=(1.0, 19.0, 2.0)),
            dict(type="LJ_wall", particle="all", stiff=0.5, dir=(1.0, 0.0, 0.0), position=-(xmin - 1.0), sigma=1.0, n=6, only_repulsive=1),
            dict(type="lowdim_trap", particle=33, stiff=0.7, rate=0.001, pos0=(5.0, 5.0, 5.0), dir=(1.0, 1.0, 0.0), visibility=(1, 0, 1)),
            dict(type="mutual_trap", particle=0, ref_particle=39, stiff=0.1, r0=1.2, PBC=1),
            dict(type="twist", particle=60, stiff=0.4, rate=0.002, base=0.3, axis=(0.0, 0.0, 1.0), pos0=(6.0, 5.0, 4.0), center=(5.0, 5.0, 5.0), mask=(1.0, 1.0, 0.0)),
            dict(type="twist", particle=61, stiff=0.2, rate=-0.001, base=0.0, axis=(1.0, 1.0, 0.0), pos0=(4.0, 6.0, 5.0), center=(5.0, 5.0, 5.0), mask=(1.0, 1.0, 1.0)),
            dict(type="sphere_smooth", particle="all", stiff=0.05, r0=6.5, r_ext=9.0, center=(10.0, 10.0, 10.0)),
            dict(type="ellipsoid", particle="all", stiff=0.1, r_2=(9.0, 8.0, 7.5), center=(10.0, 10.0, 10.0)),
            dict(type="ellipsoid", particle=100, stiff=0.3, r_2=(30.0, 30.0, 30.0), r_1=(2.0, 2.0, 2.0), center=(15.0, 15.0, 15.0))]
