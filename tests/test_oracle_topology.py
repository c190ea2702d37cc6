"""CPU: the pure-Python topology restatement (oracle/topology_oracle.py) against the vectors the
unmodified reference produced (tests/golden/topo_*.npz, made by oracle/make_golden.py)."""
import numpy as np

from oracle import topology_oracle as to


def test_topology_oracle_matches_reference(case, golden_topo, param_dir):
    g = golden_topo(case)
    t = to.build_topology(param_dir, [str(s) for s in g["in_names"]], [str(s) for s in g["in_resnames"]],
                          [int(r) for r in g["in_resids"]])
    for k in ("bond_idxs", "angle_idxs", "torsion_idxs", "bond_14_idxs"):
        assert np.array_equal(t[k], g[k]), k                       # values AND order
    for k in ("bond_para", "angle_para", "torsion_para"):
        assert np.array_equal(t[k], g[k]), k                       # bit-exact fp64
    assert np.array_equal(t["amber_types"], g["amber_types"])
    assert np.array_equal(t["out_names"], g["out_names"]) and np.array_equal(t["out_resnames"], g["out_resnames"])
    assert np.array_equal(t["atom_R"], g["atom_R"]) and np.array_equal(t["atom_e"], g["atom_e"])
    assert np.array_equal(t["atom_charges"].astype(np.float32), g["atom_charges"].astype(np.float32))


def test_nb_pair_parameters(case, golden_topo, param_dir):
    """A_ij / q_i q_j of the reference's dense matrices (banded + random sample) incl. exclusions."""
    g = golden_topo(case)
    t = to.build_topology(param_dir, [str(s) for s in g["in_names"]], [str(s) for s in g["in_resnames"]],
                          [int(r) for r in g["in_resids"]])
    A, Q = to.nb_pair_params(t, g["nb_i"], g["nb_j"])
    assert np.array_equal(A == 0, g["nb_A"] == 0)                  # exclusion pattern exact
    assert np.array_equal(Q, g["nb_q"])                            # products of fp32 charges: exact
    nz = g["nb_A"] != 0
    assert np.max(np.abs(A[nz] / g["nb_A"][nz] - 1)) < 5e-6        # reference A carries fp32 cdist noise
    n = len(g["in_names"])
    # non-zero A entries = non-excluded pairs whose two atoms both have a well depth (hydroxyl H has eps = 0)
    ex = to.exclusion_pairs(t)
    live = g["atom_e"] > 0
    n_live = int(live.sum())
    expect = n_live * (n_live - 1) // 2 - int((live[ex[:, 0]] & live[ex[:, 1]]).sum())
    assert int(g["nb_nnz_A"]) == expect
    assert int(g["nb_nnz_q"]) == n * (n - 1) // 2 - len(ex)             # no template atom has zero charge
