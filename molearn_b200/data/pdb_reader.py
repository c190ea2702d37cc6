"""Minimal multi-frame PDB / CHARMM-DCD reader and a PDBData-shaped container.

The reference's ``PDBData`` (src/molearn/data/pdb_data.py:207-400) sits on MDAnalysis and
biobox, which are outside the hot path.  The energy and scan
path needs exactly three things from it: the coordinates ``dataset`` ``[F, N, 3]`` float32
(pdb_data.py:268-270), ``get_atominfo()`` -> object array ``[N, 3]`` of
``(atom name, residue name, int resid)`` (pdb_data.py:273-286) and the scalar ``mean`` /
``std`` used to standardise.  This module provides those for fixed-column PDB text and
little-endian CHARMM DCD files (the formats of the reference's test fixtures).
"""
from __future__ import annotations

import struct
from dataclasses import dataclass, field

import numpy as np


def read_pdb_frames(path):
    """Parse ATOM/HETATM records; frames are separated by END/ENDMDL.

    Returns (atominfo [N,3] object array, coords [F,N,3] float32).  Atom order is file order
    (the order ``select_atoms`` keeps in the reference).
    """
    names, frames, cur = None, [], []
    cur_names = []
    with open(path) as fh:
        for line in fh:
            rec = line[:6]
            if rec in ("ATOM  ", "HETATM"):
                cur.append((float(line[30:38]), float(line[38:46]), float(line[46:54])))
                if names is None:
                    cur_names.append((line[12:16].strip(), line[17:20].strip(), int(line[22:26])))
            elif rec[:3] == "END":
                if cur:
                    frames.append(cur)
                    if names is None:
                        names = cur_names
                    cur = []
    if cur:
        frames.append(cur)
        if names is None:
            names = cur_names
    if not frames:
        raise ValueError(f"no ATOM records in {path}")
    n = len(frames[0])
    for f in frames:
        if len(f) != n:
            raise ValueError("frames with differing atom counts")
    info = np.empty((n, 3), dtype=object)
    for i, (a, r, k) in enumerate(names):
        info[i, 0], info[i, 1], info[i, 2] = a, r, k
    return info, np.asarray(frames, dtype=np.float32)


def read_dcd(path):
    """Little-endian CHARMM DCD without unit-cell records -> [F, N, 3] float32."""
    with open(path, "rb") as fh:
        data = fh.read()
    off = 0

    def record():
        nonlocal off
        (n,) = struct.unpack_from("<i", data, off)
        body = data[off + 4: off + 4 + n]
        (n2,) = struct.unpack_from("<i", data, off + 4 + n)
        if n != n2:
            raise ValueError("corrupt DCD record markers")
        off += 8 + n
        return body

    hdr = record()
    if hdr[:4] != b"CORD":
        raise ValueError("not a CORD DCD file")
    icntrl = struct.unpack_from("<20i", hdr, 4)
    nset, has_cell = icntrl[0], icntrl[10]
    record()  # titles
    (natom,) = struct.unpack("<i", record())
    out = np.empty((nset, natom, 3), dtype=np.float32)
    for f in range(nset):
        if has_cell:
            record()
        for c in range(3):
            out[f, :, c] = np.frombuffer(record(), dtype="<f4", count=natom)
    return out


def select_atoms(atominfo, coords, atoms):
    """Keep atoms whose name is in ``atoms`` (file order preserved), like
    ``PDBData.atomselect(atoms=[...])`` (pdb_data.py:236-256)."""
    keep = np.array([a in set(atoms) for a in atominfo[:, 0]])
    return atominfo[keep].copy(), np.ascontiguousarray(coords[:, keep])


@dataclass
class CoordinateSet:
    """The slice of ``PDBData`` the energy / scan path consumes."""

    atominfo: np.ndarray            # [N,3] object (name, resname, resid)
    coords: np.ndarray              # [F,N,3] float32, Angstrom
    mean: float = 0.0
    std: float = 1.0
    standardize: bool = True
    dataset: "object" = field(default=None, repr=False)  # torch [F,N,3] float32 (standardised)

    def prepare_dataset(self):
        import torch

        t = torch.from_numpy(np.ascontiguousarray(self.coords)).float()
        if self.standardize:
            self.mean = float(t.mean())
            self.std = float(t.std())
            t = (t - self.mean) / self.std
        self.dataset = t
        return self

    def get_atominfo(self):
        return self.atominfo.copy()

    @property
    def atoms(self):
        return sorted(set(self.atominfo[:, 0]))
