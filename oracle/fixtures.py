"""TEST INFRASTRUCTURE ONLY.  Definitions shared by the fixture generator (oracle/make_golden.py) and the tests."""
import numpy as np


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


def ext3_forces(pos):
    """The external-force list of the lattice8_ext3 fixture (second batch): repulsion_plane_moving, generic_central_force
    (gravity), LJ_cone, com, yukawa_sphere, repulsive_sphere_moving; geometry derived from the thermalised lattice8 positions so that every
    force acts on some particles without blowing the 100-step NVE run up."""
    pos = np.asarray(pos, dtype=np.float64)
    centre = np.array([10.0, 10.0, 10.0])
    # cone along +z around the box centre: apex placed so that the closest particle sits 0.95 sigma inside the surface
    alpha = 0.6
    rad = np.hypot(pos[:, 0] - centre[0], pos[:, 1] - centre[1])
    z0 = float(((pos[:, 2] * np.sin(alpha) - rad * np.cos(alpha)).min() - 0.95) / np.sin(alpha))
    # Yukawa sphere just outside the outermost particle
    R = float(np.linalg.norm(pos - centre, axis=1).max() + 1.0)
    # moving WCA sphere next to particle 150: radius such that the closest surface gap is 0.9
    org = pos[150] + np.array([2.0, 0.0, 0.0])
    r0 = float(np.linalg.norm(pos - org, axis=1).min() - 0.9)
    assert r0 > 0
    return [dict(type="repulsion_plane_moving", particle="all", ref_particle="100,101,102", stiff=0.8, dir=(1.0, 0.0, 0.0)),
            dict(type="repulsion_plane_moving", particle=7, ref_particle="250", stiff=0.5, dir=(0.0, 1.0, 1.0)),
            dict(type="generic_central_force", particle="all", center=tuple(centre), force_type="gravity", F0=0.05, inner_cut_off=4.0, outer_cut_off=9.0),
            dict(type="generic_central_force", particle=12, center=(0.0, 0.0, 0.0), force_type="gravity", F0=-0.2),
            dict(type="LJ_cone", particle="all", stiff=0.3, sigma=1.0, alpha=alpha, n=6, dir=(0.0, 0.0, 1.0), pos0=(10.0, 10.0, round(z0, 6)), only_repulsive=1),
            dict(type="com", com_list="0,1,2,3,4,5", ref_list="40,41,42,43", stiff=0.5, r0=2.0, rate=0.001),
            dict(type="com", com_list="300,301", ref_list="310", stiff=1.0, r0=0.5),
            dict(type="yukawa_sphere", particle="all", radius=round(R, 6), center=tuple(centre), debye_length=1.0, debye_A=0.2, WCA_epsilon=1.0, WCA_sigma=1.0),
            dict(type="repulsive_sphere_moving", particle="all", stiff=0.4, r0=round(r0, 6), rate=0.0005, origin=tuple(np.round(org, 6)),
                 target=tuple(np.round(org + np.array([0.2, 0.1, 0.0]), 6)), steps=400)] + _meta_traps(pos)


def _meta_traps(pos):
    """a metadynamics bias between the two strands of duplex 4 (tabulated Gaussian hill around their current separation), declared
    once per mode as the reference's metadynamics interface does"""
    p1a, p2a = list(range(160, 170)), list(range(190, 200))
    x0 = float(np.linalg.norm(pos[p1a].mean(axis=0) - pos[p2a].mean(axis=0)))
    xs = np.linspace(0.0, 5.0, 51)
    grid = ",".join("%.8f" % v for v in 1.5 * np.exp(-0.5 * ((xs - (x0 + 0.15)) / 0.4) ** 2))
    common = dict(type="meta_com_trap", p1a=",".join(map(str, p1a)), p2a=",".join(map(str, p2a)), xmin=0.0, xmax=5.0, N_grid=51, potential_grid=grid, PBC=1)
    return [dict(common, mode=1), dict(common, mode=2)]
