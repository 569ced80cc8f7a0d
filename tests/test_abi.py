"""CPU test: the C-ABI shared library loads and exports every symbol include/oxdna_b200.h declares (no compute calls)."""
import ctypes
import os
import re

import pytest

from conftest import ROOT


def _declared():
    text = open(os.path.join(ROOT, "include", "oxdna_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(oxb_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_operator_surface():
    names = _declared()
    for must in ["oxb_create", "oxb_run", "oxb_compute_forces", "oxb_update_lists", "oxb_sort", "oxb_first_step", "oxb_second_step",
                 "oxb_thermostat", "oxb_set_ext_forces", "oxb_get_pairs", "oxb_set_state", "oxb_get_state"]:
        assert must in names


def test_library_exports_every_declared_symbol():
    from oxdna_b200 import capi
    if not os.path.exists(capi.SO_PATH):
        import __graft_entry__
        __graft_entry__.build()
    L = ctypes.CDLL(capi.SO_PATH)
    missing = [n for n in _declared() if not hasattr(L, n)]
    assert not missing, missing
    assert sorted(capi.EXPORTED) == _declared()


def test_no_cpu_fallback_without_gpu():
    """On a machine without a CUDA device the product must fail loudly, never compute on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    from oxdna_b200 import capi
    with pytest.raises(capi.OxbError, match="no CUDA device|CPU fallback"):
        capi.Context(16)


def test_parameter_block_matches_oracle():
    """oxb_dna2_params_init is host-only code: compare with the oracle's independent derivation."""
    import numpy as np
    from oxdna_b200 import capi
    from oracle import oracle as O
    for T, salt in [(0.1, 0.5), (O.celsius(20.0), 1.0), (O.celsius(37.0), 0.3), (O.kelvin(350.0), 0.15)]:
        P, rcut = capi.dna2_params(T, salt)
        Q = O.dna2_params(T, salt)
        assert rcut == Q.rcut
        assert abs(P.dh_rc - Q.dh_rc) < 1e-6 and abs(P.dh_b - Q.dh_b) < 1e-8 and abs(P.dh_minus_kappa - Q.dh_minus_kappa) < 1e-6
        assert abs(P.stck_eps[0] - Q.stck.eps[0][0]) < 1e-6 and abs(P.stck_shift[7] - Q.stck.shift[1][2]) < 1e-6
        assert abs(P.hb_shift[3] - Q.hb.shift[0][3]) < 1e-6
        assert abs(P.f4[capi.NF4 - 3].t0 - Q.cxst_t1.t0) < 1e-7 and abs(P.cxst_t1_sb - Q.cxst_t1_sb) < 1e-7
        assert np.isclose(P.base_a1, Q.base_a1) and np.isclose(P.back_a2, Q.back_a2)


def test_external_force_table_mapping_matches_the_oracle_side():
    """The dict -> table-entry mapping exists twice on purpose (the product must not import oracle/): both copies have to fill identical
    entries, index pools and bias-table pools for every force type of the fixtures, and reject the same malformed input."""
    import ctypes as C

    import numpy as np

    from conftest import ext2_forces, ext3_forces, load_golden
    from oracle import oracle as O
    from oxdna_b200 import capi

    g = load_golden("lattice8")
    forces = ext2_forces(g["pos"]) + ext3_forces(g["pos"]) + [dict(type="string", particle=3, F0=0.1, rate=0.01, dir=(0, 0, 2.0)),
                                                               dict(type="trap", particle=4, stiff=1.0, rate=0.0, pos0=(1, 2, 3), dir=(1, 0, 0)),
                                                               dict(type="meta_coordination", pairs=[(0, 39), (1, 38)], coordination_type="mixed", mixed_weight=0.7,
                                                                    N_grid=3, potential_grid="0,1,4", coord_max=2.02, d0=0.35, r0=0.45, n=4)]
    assert set(capi.EXT_TYPES) == set(O.EXT_TYPES) and all(capi.EXT_TYPES[k] == O.EXT_TYPES[k] for k in capi.EXT_TYPES)
    assert {f["type"] for f in forces} == set(capi.EXT_TYPES)  # every type the library knows is exercised by a fixture
    pool_a, grid_a, pool_b, grid_b = [], [], [], []
    for f in forces:
        a, b = capi.ExtForce(), O.ExtForce()
        capi.fill_ext_entry(a, f, pool_a, grid_a)
        O.fill_ext_entry(b, f, pool_b, grid_b)
        assert C.sizeof(a) == C.sizeof(b) == capi.lib().oxb_sizeof(2)
        assert bytes(a) == bytes(b), f["type"]
    assert pool_a == pool_b and np.array_equal(grid_a, grid_b) and len(pool_a) > 0 and len(grid_a) > 0
    for bad in (dict(type="repulsion_plane_moving", particle=0, ref_particle="3,5", stiff=1.0, dir=(1, 0, 0)),
                dict(type="generic_central_force", particle=0, center=(0, 0, 0), force_type="interpolated", potential_file="x", interpolated_N=10),
                dict(type="meta_com_trap", p1a="0", p2a="1", xmin=0, xmax=1, N_grid=3, potential_grid="0,1", mode=1)):
        for fill, E in ((capi.fill_ext_entry, capi.ExtForce), (O.fill_ext_entry, O.ExtForce)):
            try:
                fill(E(), bad, [], [])
            except ValueError:
                continue
            raise AssertionError(f"{bad['type']}: malformed entry accepted")
