"""Host-side topology: Python face of ``mlp_topology_*`` (C++ builder in ``csrc/topology.cpp``).

Mirrors what ``get_convolutions(..., v=5)`` computes in the reference before its roll/conv
packing (src/molearn/loss_functions/torch_protein_energy_utils.py:359-745, 813-836): ordered
bond/angle/torsion index lists, their fp64 parameters, per-atom charge / vdW parameters and the
1-2/1-3/1-4 exclusion pairs.  Runs on the CPU; no GPU needed.
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import _lib
from .build import param_dir as _default_param_dir


def _vp(a):
    return a.ctypes.data_as(ctypes.c_void_p)


class Topology:
    """Amber ff14SB topology of a list of atoms.

    :param atominfo: object/str array ``[N, 3]`` of ``(atom name, residue name, resid)`` - what
        ``PDBData.get_atominfo()`` returns (src/molearn/data/pdb_data.py:273-286).  Unlike the
        reference (utils.py:420-487) the caller's array is NOT modified; the normalised names
        are available as ``.names`` / ``.resnames``.
    :param fix_h: CHARMM->Amber hydrogen renames (``get_convolutions(fix_h=True)``).
    :param param_dir: directory with amino12.lib / parm10.dat / frcmod.ff14SB.
    """

    def __init__(self, atominfo, fix_h: bool = False, param_dir: str | None = None):
        L = _lib.lib()
        info = np.asarray(atominfo, dtype=object)
        if info.ndim != 2 or info.shape[1] < 3:
            raise ValueError("atominfo must have shape [N, 3] = (atom name, residue name, resid)")
        n = info.shape[0]
        names = (ctypes.c_char_p * n)(*[str(a).encode() for a in info[:, 0]])
        res = (ctypes.c_char_p * n)(*[str(a).encode() for a in info[:, 1]])
        resids = np.ascontiguousarray([int(r) for r in info[:, 2]], dtype=np.int32)
        self._h = ctypes.c_void_p()
        pdir = param_dir or _default_param_dir()
        _lib.check(L.mlp_topology_build(pdir.encode(), n, names, res, _vp(resids),
                                        _lib.TOPO_FIX_H if fix_h else 0, ctypes.byref(self._h)))
        c = np.zeros(8, dtype=np.int64)
        _lib.check(L.mlp_topology_counts(self._h, _vp(c)))
        self.n_atoms, nb, na, nt, P, ne, n14 = (int(v) for v in c[:7])
        self.P = P
        self.bond_idxs = np.zeros((nb, 2), np.int32)
        self.bond_para = np.zeros((nb, 2), np.float64)
        self.angle_idxs = np.zeros((na, 3), np.int32)
        self.angle_para = np.zeros((na, 2), np.float64)
        self.torsion_idxs = np.zeros((nt, 4), np.int32)
        self.torsion_para = np.zeros((nt, 4, P), np.float64)
        self.exclusions = np.zeros((ne, 2), np.int32)
        _lib.check(L.mlp_topology_export(self._h, _vp(self.bond_idxs), _vp(self.bond_para), _vp(self.angle_idxs),
                                         _vp(self.angle_para), _vp(self.torsion_idxs), _vp(self.torsion_para),
                                         _vp(self.exclusions)))
        self.bond_14_idxs = np.zeros((n14, 2), np.int32)
        _lib.check(L.mlp_topology_pairs14(self._h, _vp(self.bond_14_idxs)))
        # 1-4 scaled parameters (utils.py:833-836; computed by the reference, used by no energy term)
        self.vdw_14R, self.vdw_14e, self.q1q2_14 = (np.zeros(n14, np.float32) for _ in range(3))
        _lib.check(L.mlp_topology_params14(self._h, _vp(self.vdw_14R), _vp(self.vdw_14e), _vp(self.q1q2_14)))
        nm, rn, ty = (np.zeros((n, 8), np.uint8) for _ in range(3))
        self.atom_charges = np.zeros(n, np.float64)
        self.atom_R = np.zeros(n, np.float32)
        self.atom_e = np.zeros(n, np.float32)
        _lib.check(L.mlp_topology_atoms(self._h, _vp(nm), _vp(rn), _vp(ty), _vp(self.atom_charges),
                                        _vp(self.atom_R), _vp(self.atom_e)))
        dec = lambda a: np.array([bytes(r).split(b"\0", 1)[0].decode() for r in a])
        self.names, self.resnames, self.amber_types = dec(nm), dec(rn), dec(ty)
        self.resids = resids

    @property
    def handle(self):
        return self._h

    @property
    def n_nb_pairs(self) -> int:
        """Unordered non-excluded pairs N(N-1)/2 - n_excl (the NB kernel's algorithmic work)."""
        n = self.n_atoms
        return n * (n - 1) // 2 - len(self.exclusions)

    def bonded_plan(self, tile_atoms: int = 128):
        """Work items of the bonded kernel for tiles of ``tile_atoms`` atoms (host-only; ``mlp_bonded_plan``).

        Returns ``(items [n, 9] int32, stats)``; columns: tile, i, j, k, l, torsion id | -1, angle id | -1,
        bond id | -1, bond-on-(j,k) flag.  An item is up to one torsion + one angle (i,j,k) + one bond
        ((i,j) or (j,k)) on the same four atoms."""
        L = _lib.lib()
        n = ctypes.c_int64(0)
        st = np.zeros(4, np.int64)
        _lib.check(L.mlp_bonded_plan(self._h, tile_atoms, ctypes.byref(n), None, 0, _vp(st)))
        items = np.zeros((n.value, 9), np.int32)
        _lib.check(L.mlp_bonded_plan(self._h, tile_atoms, ctypes.byref(n), _vp(items), n.value, _vp(st)))
        return items, dict(tiles=int(st[0]), max_contributions_per_atom=int(st[1]), torsion_sines=bool(st[2]),
                           live_torsions=int(st[3]))

    def nb_pair_table(self):
        """Lennard-Jones classes and the squared pair lengths of the repulsive pair kernel's type-pair table
        (host-only; ``mlp_nb_pair_table``).  Returns ``(atom_class [N] int32, w2 [C, C] float32)`` with
        ``w2[a, b]**6 / 12 == vdw_A[i, j]`` (utils.py:816) for non-excluded atoms i, j of classes a, b."""
        L = _lib.lib()
        n = ctypes.c_int64(0)
        cls = np.zeros(self.n_atoms, np.int32)
        _lib.check(L.mlp_nb_pair_table(self._h, _vp(cls), None, 0, ctypes.byref(n)))
        w2 = np.zeros((n.value, n.value), np.float32)
        _lib.check(L.mlp_nb_pair_table(self._h, None, _vp(w2), w2.size, ctypes.byref(n)))
        return cls, w2

    def as_dict(self):
        """Arrays under the names the golden fixtures / oracle use."""
        return dict(bond_idxs=self.bond_idxs, angle_idxs=self.angle_idxs, torsion_idxs=self.torsion_idxs,
                    bond_14_idxs=self.bond_14_idxs, bond_para=self.bond_para, angle_para=self.angle_para,
                    torsion_para=self.torsion_para, atom_charges=self.atom_charges, atom_R=self.atom_R,
                    atom_e=self.atom_e, amber_types=self.amber_types, out_names=self.names,
                    out_resnames=self.resnames)

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            try:
                _lib.lib().mlp_topology_destroy(h)
            except Exception:
                pass
