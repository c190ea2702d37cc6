"""TEST INFRASTRUCTURE ONLY - pure-Python restatement of the reference's force-field parser
and topology construction (rows T-1, T-2, T-4 of SURVEY.md section 8).

It is the checker for the product's C++ builder (``molearn_b200/csrc/topology.cpp``); the
product never imports it.  Pinned against the reference's own output in
``tests/golden/topo_*.npz`` (``tests/test_oracle_topology.py``).

Reference being restated (all under src/molearn/loss_functions/):
  read_lib_file                 torch_protein_energy_utils.py:19-83
  get_amber_parameters          torch_protein_energy_utils.py:86-189
  amber_card_type_2..10B        torch_protein_energy_utils.py:192-356
  get_convolutions (v=5 path)   torch_protein_energy_utils.py:420-504, 578-607, 659-745, 813-836
The Amber text files themselves are read at run time from a directory the caller names.
"""
from __future__ import annotations

import math
import os

import numpy as np

DEG = math.pi / 180.0  # np.deg2rad multiplies by this constant in fp64 (utils.py:180-184)


# ----------------------------------------------------------------------------- amino12.lib
def parse_lib(path):
    """utils.py:19-83.  Returns (atom_type[res][pdbname], charge_by_type[res][ambertype],
    connectivity[res][pdbname] -> list of pdb names)."""
    with open(path) as fh:
        lines = fh.readlines()
    atom_type, charge, conn, by_index = {}, {}, {}, {}
    # residue index block (utils.py:30-50)
    for k, line in enumerate(lines):
        if line.split() == ["!!index", "array", "str"]:
            for l2 in lines[k + 1:]:
                if l2[0] != " ":
                    break
                tok = l2.split()
                if len(tok) != 1 and len(tok[0]) != 5:
                    break
                res = tok[0][1:-1]
                atom_type[res], charge[res], conn[res], by_index[res] = {}, {}, {}, {}
            break
    for k, line in enumerate(lines):
        if line[0:7] != "!entry.":
            continue
        res = line[7:10]
        if line[10:22] == ".unit.atoms ":
            for l2 in lines[k + 1:]:
                if l2[0] != " ":
                    break
                tok = l2.split()
                pdb_name, amber = tok[0][1:-1], tok[1][1:-1]
                atom_type[res][pdb_name] = amber
                charge[res][amber] = float(tok[7])  # keyed by TYPE: last atom of a type wins (:66)
                by_index[res][int(tok[5])] = pdb_name
                conn[res][pdb_name] = []
        elif line[10:29] == ".unit.connectivity ":
            for l2 in lines[k + 1:]:
                if l2[0] != " ":
                    break
                tok = l2.split()
                if len(tok) != 3:
                    break
                a, b = by_index[res][int(tok[0])], by_index[res][int(tok[1])]
                conn[res][a].append(b)
                conn[res][b].append(a)
    return atom_type, charge, conn


# ------------------------------------------------------------------ parm10.dat / frcmod cards
class ForceField:
    def __init__(self):
        self.bond = {}      # (t1,t2) -> (K, r0)
        self.angle = {}     # (t1,t2,t3) -> (K, theta0 in degrees until finish())
        self.torsion = {}   # (t1,t2,t3,t4) -> [factor list, barrier list, phase list, period list]
        self.equiv = {}     # first symbol -> all symbols of the line
        self.vdw = {}       # symbol -> (R*, eps)


def _until_blank(it):
    for line in it:
        if line == "\n" or line.strip() == "":
            return
        yield line


def _card_bonds(it, ff):  # utils.py:217-233
    for line in _until_blank(it):
        k, r0 = line[5:25].split()
        ff.bond[(line[0:2].strip(), line[3:5].strip())] = (float(k), float(r0))


def _card_angles(it, ff):  # utils.py:237-256
    for line in _until_blank(it):
        k, th = line[8:28].split()
        ff.angle[(line[0:2].strip(), line[3:5].strip(), line[6:8].strip())] = (float(k), float(th))


def _card_torsions(it, ff):  # utils.py:259-293
    for line in _until_blank(it):
        key = (line[0:2].strip(), line[3:5].strip(), line[6:8].strip(), line[9:11].strip())
        c = line[11:55].split()
        if len(c) != 4:
            raise Exception("I wanted four values here?")
        new = [int(c[0]), float(c[1]), float(c[2]), float(c[3])]
        if key in ff.torsion:
            last = ff.torsion[key][3][-1]
            if last > 0:       # previous series was complete: this line starts a replacement
                ff.torsion[key] = [[v] for v in new]
            elif last < 0:     # negative period = "more terms follow"
                for lst, v in zip(ff.torsion[key], new):
                    lst.append(v)
        else:
            ff.torsion[key] = [[v] for v in new]


def _card_skip(it):
    for _ in _until_blank(it):
        pass


def _card_equiv(it, ff):  # utils.py:341-347
    for line in _until_blank(it):
        c = line.split()
        ff.equiv[c[0]] = c


def _card_vdw(it, ff):  # utils.py:350-356
    for line in _until_blank(it):
        c = line.split()
        ff.vdw[c[0]] = (float(c[1]), float(c[2]))


def parse_forcefield(param_dir):
    """utils.py:86-189 for ('amino12.lib', 'parm10.dat', 'frcmod.ff14SB')."""
    atom_type, charge, conn = parse_lib(os.path.join(param_dir, "amino12.lib"))
    ff = ForceField()
    with open(os.path.join(param_dir, "parm10.dat")) as fh:
        it = iter(fh)
        next(it)            # title
        _card_skip(it)      # masses (card 2) - unused by the energy
        next(it)            # hydrophilic symbols (card 3, one line)
        _card_bonds(it, ff)
        _card_angles(it, ff)
        _card_torsions(it, ff)
        _card_skip(it)      # impropers (card 7) - parsed by the reference, never used
        _card_skip(it)      # H-bond 10-12 (card 8) - unused
        _card_equiv(it, ff)
        for line in it:
            tok = line.split()
            if len(tok) > 1 and tok[1] == "RE":
                _card_vdw(it, ff)
    with open(os.path.join(param_dir, "frcmod.ff14SB")) as fh:
        it = iter(fh)
        next(it)
        for line in it:     # utils.py:158-177
            tag = line[:4]
            if tag == "MASS":
                _card_skip(it)
            if tag == "BOND":
                _card_bonds(it, ff)
            if tag == "ANGL":
                _card_angles(it, ff)
            if tag == "DIHE":
                _card_torsions(it, ff)
            if tag == "IMPR":
                _card_skip(it)
            if tag == "HBON":
                _card_skip(it)
            if tag == "NONB":
                _card_vdw(it, ff)
    ff.angle = {k: (v[0], v[1] * DEG) for k, v in ff.angle.items()}
    for v in ff.torsion.values():
        v[2] = [p * DEG for p in v[2]]
    return atom_type, charge, conn, ff


# ------------------------------------------------------------------------- name normalisation
_FIX_H_ANY = [("HN", "H"), ("1HD2", "HD21"), ("2HD2", "HD22"), ("1HG2", "HG21"), ("2HG2", "HG22"),
              ("3HG2", "HG23"), ("3HG1", "HG13"), ("1HG1", "HG11"), ("2HG1", "HG12"),
              ("1HD1", "HD11"), ("2HD1", "HD12"), ("3HD1", "HD13"), ("3HD2", "HD23"),
              ("1HH1", "HH11"), ("2HH1", "HH12"), ("1HH2", "HH21"), ("2HH2", "HH22"),
              ("1HE2", "HE21"), ("2HE2", "HE22")]
_FIX_H_PRE = [("HB1", "MET", "HB3"), ("HG1", "MET", "HG3"), ("HB1", "ASN", "HB3")]
_FIX_H_POST = [("HG11", "ILE", "HG13"), ("CD", "ILE", "CD1"), ("HD1", "ILE", "HD11"),
               ("HD2", "ILE", "HD12"), ("HD3", "ILE", "HD13"), ("HB1", "PHE", "HB3"),
               ("HB1", "GLU", "HB3"), ("HG1", "GLU", "HG3"), ("HB1", "LEU", "HB3"),
               ("HB1", "ARG", "HB3"), ("HG1", "ARG", "HG3"), ("HD1", "ARG", "HD3"),
               ("HB1", "ASP", "HB3"), ("HA1", "GLY", "HA3"), ("HB1", "LYS", "HB3"),
               ("HG1", "LYS", "HG3"), ("HD1", "LYS", "HD3"), ("HE1", "LYS", "HE3"),
               ("HB1", "TYR", "HB3"), ("HB1", "HIP", "HB3"), ("HB1", "SER", "HB3"),
               ("HG1", "SER", "HG"), ("HB1", "PRO", "HB3"), ("HG1", "PRO", "HG3"),
               ("HD1", "PRO", "HD3"), ("HB1", "LEU", "HB3"), ("HB1", "GLN", "HB3"),
               ("HG1", "GLN", "HG3"), ("HB1", "TRP", "HB3")]


def normalise_names(names, resnames, resids, fix_h=False):
    """utils.py:420-487 (the reference does this in place on the caller's array)."""
    names, resnames = list(names), list(resnames)
    n = len(names)
    for i in range(n):
        if names[i] == "OXT":
            names[i] = "O"
    for i in range(n):
        if resnames[i] == "HSD":
            resnames[i] = "HID"
    for i in range(n):
        if resnames[i] == "HSE":
            resnames[i] = "HIE"
    # grouped by resid VALUE over the whole array (np.unique, utils.py:425-433)
    groups = {}
    for i in range(n):
        groups.setdefault(resids[i], []).append(i)
    for rid in sorted(groups):
        idx = groups[rid]
        if all(resnames[i] == "HIS" for i in idx):
            hd1 = any(names[i] == "HD1" for i in idx)
            he2 = any(names[i] == "HE2" for i in idx)
            new = "HIP" if (hd1 and he2) else "HID" if hd1 else "HIE" if he2 else None
            if new:
                for i in idx:
                    resnames[i] = new
    for i in range(n):
        if resnames[i] == "HIS":
            resnames[i] = "HIE"
    if fix_h:
        for a, r, b in _FIX_H_PRE:
            for i in range(n):
                if names[i] == a and resnames[i] == r:
                    names[i] = b
        for a, b in _FIX_H_ANY:
            for i in range(n):
                if names[i] == a:
                    names[i] = b
        for a, r, b in _FIX_H_POST:
            for i in range(n):
                if names[i] == a and resnames[i] == r:
                    names[i] = b
    return names, resnames


# --------------------------------------------------------------------------------- topology
def build_topology(param_dir, names, resnames, resids, fix_h=False):
    """Restates get_convolutions(v=5) up to the explicit index/parameter lists.

    Returns a dict with the same arrays ``oracle/make_golden.py`` captures from the reference.
    """
    atom_type, charge, conn, ff = parse_forcefield(param_dir)
    names, resnames = normalise_names(names, resnames, resids, fix_h)
    n = len(names)
    pnames = [a if a not in ("H2", "H3") else "H" for a in names]            # utils.py:490
    types = [atom_type[r][a] for a, r in zip(pnames, resnames)]              # utils.py:489
    charges = [charge[r][t] for t, r in zip(types, resnames)]                # utils.py:493
    inv_equiv = {}
    for first, syms in ff.equiv.items():                                     # utils.py:498-502
        for s in syms:
            inv_equiv[s] = first
    atom_R = np.array([ff.vdw[inv_equiv.get(t, t)][0] for t in types], dtype=np.float32)
    atom_e = np.array([ff.vdw[inv_equiv.get(t, t)][1] for t in types], dtype=np.float32)

    # bonds, v=5 (utils.py:578-607)
    bonds, tracker = [], [[] for _ in range(n)]
    cur_resid, current, previous = -9999, [], []
    for i1 in range(n):
        a1, res, rid = pnames[i1], resnames[i1], resids[i1]
        assert a1 in conn[res]
        if rid != cur_resid:
            previous, current, cur_resid = current, [], rid
        if a1 == "N":
            for a2, i2 in previous:
                if a2 == "C":
                    tracker[i1].append(i2)
                    tracker[i2].append(i1)
                    bonds.append((i2, i1))
        for a2, i2 in current:
            if a2 in conn[res][a1] and a1 in conn[res][a2]:
                tracker[i1].append(i2)
                tracker[i2].append(i1)
                bonds.append((i2, i1))
        current.append((a1, i1))

    # angles / torsions / 1-4 (utils.py:659-687)
    angles, torsions, pairs14 = [], [], []
    for a1 in range(n):
        for a2 in tracker[a1]:
            for a3 in tracker[a2]:
                if a3 > a1:
                    angles.append((a1, a2, a3))
                if a3 != a1:
                    for a4 in tracker[a3]:
                        if a4 > a1 and a2 != a4:
                            torsions.append((a1, a2, a3, a4))
                            pairs14.append((a1, a4))

    # parameters (utils.py:705-745)
    bond_para = np.zeros((len(bonds), 2))
    for k, (i, j) in enumerate(bonds):
        key = (types[i], types[j])
        kf, r0 = ff.bond[key] if key in ff.bond else ff.bond[(key[1], key[0])]
        bond_para[k] = (r0, kf)
    angle_para = np.zeros((len(angles), 2))
    for k, (i, j, l) in enumerate(angles):
        key = (types[i], types[j], types[l])
        kf, th = ff.angle[key] if key in ff.angle else ff.angle[(key[2], key[1], key[0])]
        angle_para[k] = (th, kf)
    found, pmax = [], 0
    for (i, j, k_, l) in torsions:
        t = (types[i], types[j], types[k_], types[l])
        for key in (t, t[::-1], ("X", t[2], t[1], "X"), ("X", t[1], t[2], "X")):
            if key in ff.torsion:
                found.append(ff.torsion[key])
                pmax = max(pmax, len(ff.torsion[key][1]))
                break
        else:
            raise KeyError(t)
    torsion_para = np.zeros((len(torsions), 4, pmax))
    torsion_para[:, 0, :] = 1.0
    for k, para in enumerate(found):
        torsion_para[k, :, :len(para[0])] = para
    torsion_para[:, 3, :] = np.abs(torsion_para[:, 3, :])

    return dict(out_names=np.array(names), out_resnames=np.array(resnames),
                amber_types=np.array(types),
                bond_idxs=np.array(bonds, dtype=np.int64).reshape(-1, 2),
                angle_idxs=np.array(angles, dtype=np.int64).reshape(-1, 3),
                torsion_idxs=np.array(torsions, dtype=np.int64).reshape(-1, 4),
                bond_14_idxs=np.array(pairs14, dtype=np.int64).reshape(-1, 2),
                bond_para=bond_para, angle_para=angle_para, torsion_para=torsion_para,
                atom_charges=np.array(charges, dtype=np.float64), atom_R=atom_R, atom_e=atom_e)


def exclusion_pairs(topo):
    """Unique unordered (i<j) pairs zeroed at utils.py:818-830 with pair indexing
    (the intended semantics): 1-2, 1-3 and 1-4 ends."""
    p = np.concatenate([topo["bond_idxs"], topo["angle_idxs"][:, (0, 2)], topo["torsion_idxs"][:, (0, 3)]])
    p = np.sort(p, axis=1)
    return np.unique(p, axis=0)


def nb_pair_params(topo, ii, jj):
    """A_ij and q_i q_j for the pairs (ii, jj) as the reference's dense matrices hold them
    (utils.py:813-830, torch_protein_energy.py:125-127), computed directly in fp32 from the
    per-atom fp32 values (the reference's own A carries ~1e-6 relative cdist noise)."""
    R, e = topo["atom_R"], topo["atom_e"]
    q = topo["atom_charges"].astype(np.float32)
    ii, jj = np.asarray(ii), np.asarray(jj)
    upper = ii < jj
    excl = {(int(a), int(b)) for a, b in exclusion_pairs(topo)}
    keep = np.array([u and (int(a), int(b)) not in excl for u, a, b in zip(upper, ii, jj)])
    Rij = np.float32(0.5) * np.abs(R[ii] + R[jj])
    eij = np.sqrt(e[ii] * e[jj])
    A = (eij * Rij ** 12).astype(np.float32)
    Q = (q[ii] * q[jj]).astype(np.float32)
    return np.where(keep, A, np.float32(0)), np.where(keep, Q, np.float32(0))
