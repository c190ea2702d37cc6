"""CPU: the product's C++ topology builder (C ABI mlp_topology_*) - bit-exact against the
reference's vectors, error behaviour, and the exported symbol table."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT, atominfo_from


def test_library_exports_every_declared_symbol():
    from molearn_b200 import _lib
    header = open(os.path.join(ROOT, "include", "molearn_b200.h")).read()
    declared = set(re.findall(r"\b(mlp_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    L = ctypes.CDLL(_lib.library_path())
    for name in declared:
        assert hasattr(L, name), name
    assert _lib.lib().mlp_abi_version() == 1


def test_builder_bit_exact(case, golden_topo, param_dir):
    from molearn_b200 import Topology
    g = golden_topo(case)
    info = atominfo_from(g)
    before = info.copy()
    t = Topology(info, param_dir=param_dir)
    assert np.array_equal(info, before)                             # caller's array untouched
    assert np.array_equal(t.bond_idxs, g["bond_idxs"])
    assert np.array_equal(t.angle_idxs, g["angle_idxs"])
    assert np.array_equal(t.torsion_idxs, g["torsion_idxs"])
    assert np.array_equal(t.bond_14_idxs, g["bond_14_idxs"])
    assert np.array_equal(t.bond_para, g["bond_para"])
    assert np.array_equal(t.angle_para, g["angle_para"])
    assert np.array_equal(t.torsion_para, g["torsion_para"])
    assert np.array_equal(t.amber_types, g["amber_types"])
    assert np.array_equal(t.names, g["out_names"]) and np.array_equal(t.resnames, g["out_resnames"])
    assert np.array_equal(t.atom_R, g["atom_R"]) and np.array_equal(t.atom_e, g["atom_e"])
    assert np.array_equal(t.atom_charges.astype(np.float32), g["atom_charges"].astype(np.float32))
    n = t.n_atoms
    assert t.n_nb_pairs == int(g["nb_nnz_q"]) == n * (n - 1) // 2 - len(t.exclusions)
    ex = t.exclusions
    assert np.all(ex[:, 0] < ex[:, 1]) and len(np.unique(ex, axis=0)) == len(ex)


@pytest.mark.parametrize("case14", ["bbcb", "pep28"])
def test_scaled_14_parameters_bit_exact(case14, golden_topo, param_dir):
    """utils.py:833-836: vdw_14R, vdw_14e (sum under the root, as written there), q1q2_14 - fp32, bit for bit."""
    from conftest import load_golden
    from molearn_b200 import Topology
    g = load_golden("energy_full_" + case14)
    t = Topology(atominfo_from(golden_topo(case14)), param_dir=param_dir)
    assert np.array_equal(t.bond_14_idxs, g["bond_14_idxs"])
    assert np.array_equal(t.vdw_14R, g["vdw_14R"].astype(np.float32))
    assert np.array_equal(t.vdw_14e, g["vdw_14e"].astype(np.float32))
    assert np.array_equal(t.q1q2_14, g["q1q2_14"].astype(np.float32))


def test_nb_pair_table_reproduces_the_reference_lj_matrix(case, golden_topo, param_dir):
    """The type-pair table of the repulsive pair kernel (``mlp_nb_pair_table``: w^2 per pair of Lennard-Jones classes,
    fp64 on the host, rounded once) against the reference's own dense ``vdw_A`` matrix (utils.py:813-816, sampled into
    the golden files by oracle/make_golden.py): (w_ij^2)^6 / 12 = eps_ij R_ij^12 for every sampled non-excluded pair,
    and atoms share a class exactly when they share (R, eps)."""
    from molearn_b200 import Topology
    g = golden_topo(case)
    t = Topology(atominfo_from(g), param_dir=param_dir)
    cls, w2 = t.nb_pair_table()
    C = w2.shape[0]
    assert cls.min() == 0 and cls.max() == C - 1 and np.array_equal(w2, w2.T)
    keys = np.stack([g["atom_R"], g["atom_e"]], 1)
    assert len(np.unique(keys, axis=0)) == C
    for c in range(C):
        assert len(np.unique(keys[cls == c], axis=0)) == 1          # one (R, eps) per class
    nz = g["nb_A"] != 0                                              # excluded pairs are zero in the reference's matrix
    i, j = g["nb_i"][nz], g["nb_j"][nz]
    A = w2[cls[i], cls[j]].astype(np.float64) ** 6 / 12.0
    assert np.max(np.abs(A / g["nb_A"][nz] - 1)) < 5e-6             # the reference builds A in fp32; w^2 is rounded once


def test_builder_matches_python_oracle_on_edge_inputs(param_dir):
    """Ragged / unusual inputs the golden files do not hold: OXT, HIS variants, CHARMM names,
    hydrogens with fix_h, a single residue, an empty list."""
    from molearn_b200 import Topology
    from oracle import topology_oracle as to
    cases = {
        "single_gly": [("N", "GLY", 1), ("CA", "GLY", 1), ("C", "GLY", 1), ("O", "GLY", 1)],
        "oxt_his": [("N", "HIS", 5), ("CA", "HIS", 5), ("CB", "HIS", 5), ("CG", "HIS", 5), ("ND1", "HIS", 5),
                    ("HD1", "HIS", 5), ("CD2", "HIS", 5), ("CE1", "HIS", 5), ("NE2", "HIS", 5), ("C", "HIS", 5),
                    ("O", "HIS", 5), ("N", "HSD", 6), ("CA", "HSD", 6), ("C", "HSD", 6), ("O", "HSD", 6), ("OXT", "HSD", 6)],
        "hydrogens": [("N", "ALA", 1), ("HN", "ALA", 1), ("CA", "ALA", 1), ("HA", "ALA", 1), ("CB", "ALA", 1),
                      ("HB1", "ALA", 1), ("HB2", "ALA", 1), ("HB3", "ALA", 1), ("C", "ALA", 1), ("O", "ALA", 1),
                      ("N", "SER", 2), ("H", "SER", 2), ("CA", "SER", 2), ("CB", "SER", 2), ("OG", "SER", 2),
                      ("HG1", "SER", 2), ("C", "SER", 2), ("O", "SER", 2)],
        "pro_ring": [("N", "PRO", 3), ("CD", "PRO", 3), ("CG", "PRO", 3), ("CB", "PRO", 3), ("CA", "PRO", 3),
                     ("C", "PRO", 3), ("O", "PRO", 3), ("N", "TRP", 4), ("CA", "TRP", 4), ("C", "TRP", 4)],
    }
    for name, atoms in cases.items():
        info = np.array(atoms, dtype=object)
        fix_h = name == "hydrogens"
        t = Topology(info, fix_h=fix_h, param_dir=param_dir)
        o = to.build_topology(param_dir, list(info[:, 0]), list(info[:, 1]), [int(r) for r in info[:, 2]], fix_h=fix_h)
        for k in ("bond_idxs", "angle_idxs", "torsion_idxs", "bond_para", "angle_para", "torsion_para"):
            assert np.array_equal(getattr(t, k), o[k]), (name, k)
        assert np.array_equal(t.amber_types, o["amber_types"]), name
        assert np.array_equal(t.exclusions, to.exclusion_pairs(o).reshape(-1, 2)), name
    empty = Topology(np.empty((0, 3), dtype=object), param_dir=param_dir)
    assert empty.n_atoms == 0 and len(empty.bond_idxs) == 0


def test_builder_errors(param_dir):
    from molearn_b200 import Topology
    with pytest.raises(KeyError):                                   # unknown residue (reference: KeyError, utils.py:489)
        Topology(np.array([("N", "XYZ", 1)], dtype=object), param_dir=param_dir)
    with pytest.raises(KeyError):                                   # unknown atom name
        Topology(np.array([("QQ", "ALA", 1)], dtype=object), param_dir=param_dir)
    with pytest.raises(FileNotFoundError):
        Topology(np.array([("N", "ALA", 1)], dtype=object), param_dir="/nonexistent")


def test_no_cpu_fallback(golden_topo):
    import torch
    from molearn_b200 import TorchProteinEnergy
    info = atominfo_from(golden_topo("bbcb"))
    with pytest.raises(RuntimeError):
        TorchProteinEnergy(None, info, device=torch.device("cpu"))
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError):                           # no GPU: create fails loudly
            TorchProteinEnergy(None, info, device=torch.device("cuda:0"))


def test_builder_matches_python_oracle_on_random_peptides(param_dir):
    """Bit-exact agreement (values, order, fp64 parameters) on seeded random sequences over all 28 templates,
    all-atom and heavy-atom selections, non-consecutive and repeated residue numbers."""
    from molearn_b200 import Topology
    from oracle import topology_oracle as to
    at, _, _, _ = to.parse_forcefield(param_dir)
    residues = sorted(at)
    rng = np.random.default_rng(7)
    for trial in range(8):
        heavy = trial % 2 == 1
        seq = [residues[i] for i in rng.integers(0, len(residues), int(rng.integers(5, 60)))]
        rid, atoms = int(rng.integers(-5, 50)), []
        for r in seq:
            rid += int(rng.integers(1, 4))                                # gaps in the numbering
            atoms += [(a, r, rid) for a in at[r] if not (heavy and a.startswith("H"))]
        info = np.array(atoms, dtype=object)
        t = Topology(info, param_dir=param_dir)
        o = to.build_topology(param_dir, list(info[:, 0]), list(info[:, 1]), [int(r) for r in info[:, 2]])
        for k in ("bond_idxs", "angle_idxs", "torsion_idxs", "bond_14_idxs", "bond_para", "angle_para", "torsion_para"):
            assert np.array_equal(getattr(t, k), o[k]), (trial, k)
        assert np.array_equal(t.amber_types, o["amber_types"]) and np.array_equal(t.atom_charges, o["atom_charges"])
        assert np.array_equal(t.atom_R, o["atom_R"]) and np.array_equal(t.atom_e, o["atom_e"])
        assert np.array_equal(t.exclusions, to.exclusion_pairs(o).reshape(-1, 2))


def _check_bonded_plan(t, tile, label):
    """Every tile evaluates exactly the terms that touch one of its atoms, each once; an item's angle / bond sit on
    the item's own atoms in the positions the kernel assumes (angle = (i,j,k) centred on j, bond = (i,j) or (j,k)).
    Returns False if the topology is too dense for this tile size (mlp_energy_create then retries smaller)."""
    try:
        items, st = t.bonded_plan(tile)
    except ValueError as exc:          # more than 2 items per thread
        assert "too dense" in str(exc), label
        return False
    n = t.n_atoms
    assert st["tiles"] == -(-n // tile)
    amp = t.torsion_para[:, 1, :] / t.torsion_para[:, 0, :]
    live = np.flatnonzero(np.any(amp.astype(np.float32) != 0, axis=1))
    assert st["live_torsions"] == len(live)
    for T in range(st["tiles"]):
        lo, hi = T * tile, min(n, (T + 1) * tile)
        own = lambda idx: np.any((idx >= lo) & (idx < hi), axis=-1)
        rows = items[items[:, 0] == T]
        for col, want in ((5, live[own(t.torsion_idxs[live])] if len(live) else live),
                          (6, np.flatnonzero(own(t.angle_idxs)) if len(t.angle_idxs) else np.zeros(0, int)),
                          (7, np.flatnonzero(own(t.bond_idxs)) if len(t.bond_idxs) else np.zeros(0, int))):
            got = rows[rows[:, col] >= 0, col]
            assert len(got) == len(set(got.tolist())), (label, T, col)             # each term once per tile
            assert set(got.tolist()) == set(want.tolist()), (label, T, col)       # and exactly the tile's terms
        per_atom = np.zeros(n, np.int64)
        for r in rows:
            a = r[1:5]
            if r[5] >= 0:
                q = t.torsion_idxs[r[5]]
                assert list(a) == list(q) or list(a) == list(q[::-1]), label
            if r[6] >= 0:
                q = t.angle_idxs[r[6]]
                assert a[1] == q[1] and {a[0], a[2]} == {q[0], q[2]}, label
            if r[7] >= 0:
                pair = (a[1], a[2]) if r[8] else (a[0], a[1])
                assert set(pair) == set(t.bond_idxs[r[7]].tolist()), label
            roles = 4 if r[5] >= 0 else 3 if r[6] >= 0 else 2
            assert roles > 2 or r[7] >= 0, label                                  # no empty items
            for x in a[:roles]:
                per_atom[x] += 1
        # gradient slots: contributions of a tile's own atoms fit the slot table
        assert per_atom[lo:hi].max(initial=0) <= st["max_contributions_per_atom"], label
    n_terms = len(live) + len(t.angle_idxs) + len(t.bond_idxs)
    assert len(items) <= n_terms                                                  # the matching never adds work
    return True


@pytest.mark.parametrize("tile", [128, 96, 40])
def test_bonded_work_plan_covers_every_term_once(case, golden_topo, param_dir, tile):
    """Host logic of the bonded kernel (mlp_bonded_plan) on the golden systems."""
    from molearn_b200 import Topology
    t = Topology(atominfo_from(golden_topo(case)), param_dir=param_dir)
    ok = _check_bonded_plan(t, tile, (case, tile))
    assert ok or (case == "pep28" and tile > 64)
    if ok and case == "bbcb":
        items, _ = t.bonded_plan(tile)
        assert len(items) < 0.45 * (len(t.torsion_idxs) + len(t.angle_idxs) + len(t.bond_idxs))   # backbone: ~3 terms per item


def test_bonded_work_plan_on_random_peptides(param_dir):
    """The same invariants on seeded random sequences over all 28 residue templates (all-atom and heavy-atom), with
    tile sizes that put the tile boundaries anywhere."""
    from molearn_b200 import Topology
    from oracle import topology_oracle as to
    at, _, _, _ = to.parse_forcefield(param_dir)
    residues = sorted(at)
    rng = np.random.default_rng(19)
    checked = 0
    for trial in range(6):
        heavy = trial % 2 == 0
        seq = [residues[i] for i in rng.integers(0, len(residues), int(rng.integers(3, 40)))]
        rid, atoms = 0, []
        for r in seq:
            rid += 1
            atoms += [(a, r, rid) for a in at[r] if not (heavy and a.startswith("H"))]
        t = Topology(np.array(atoms, dtype=object), param_dir=param_dir)
        for tile in (int(rng.integers(8, 64)), 64, 128):
            checked += bool(_check_bonded_plan(t, tile, (trial, heavy, tile)))
    assert checked >= 8
