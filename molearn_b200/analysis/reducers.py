"""Device-side scan reducers (C ABI ``mlp_rmsd_eval``).

Replace the host NumPy / biobox loops of ``MolearnAnalysis.scan_error`` and
``scan_error_from_target`` (src/molearn/analysis/analyser.py:626-737): only ``[B, K]`` scalars
leave the GPU instead of the decoded ``[B, N, 3]`` structures.
"""
from __future__ import annotations

import torch

from .. import _lib


def _check_structures(x, what="x"):
    if not x.is_cuda:
        raise RuntimeError(f"{what}: the scan reducers need CUDA tensors (there is no CPU fallback)")
    if x.dtype != torch.float32:
        raise TypeError(f"{what} must be float32, got {x.dtype}")
    if x.dim() != 3 or x.shape[2] != 3:
        raise ValueError(f"{what}: expected [B, N, 3], got {tuple(x.shape)}")


def rmsd_to_targets(x, targets, scale: float = 1.0, shift: float = 0.0, align: bool = False, paired: bool = False):
    """RMSD ``sqrt(mean_atoms sum_xyz (x*scale+shift - t)^2)`` of every structure to every target
    (the intended per-point formula, SURVEY S-3/S-4).  ``align=True``: after optimal rigid
    superposition (Kabsch), the quantity biobox ``Molecule.rmsd`` returns in the reference.

    :param x: ``[B, N, 3]`` float32 CUDA tensor (any strides - e.g. a slice of a decoder output).
    :param targets: ``[K, N, 3]`` float32 CUDA tensor on the same device, in the units of ``x*scale+shift``.
    :param paired: ``K == B`` and structure ``b`` is scored against target ``b`` only.
    :returns: ``[B, K]`` float32 (``[B]`` when ``paired``).
    """
    _check_structures(x)
    if targets.dim() == 2:
        targets = targets.unsqueeze(0)
    if targets.device != x.device:
        raise ValueError(f"targets live on {targets.device}, x on {x.device}")
    B, N, _ = x.shape
    if tuple(targets.shape[1:]) != (N, 3):
        raise ValueError(f"targets {tuple(targets.shape)} do not match x {tuple(x.shape)}")
    targets = targets.contiguous().float()
    K = targets.shape[0]
    if paired and K != B:
        raise ValueError(f"paired RMSD needs one target per structure, got {K} targets for {B} structures")
    out = torch.empty((B,) if paired else (B, K), dtype=torch.float32, device=x.device)
    flags = (_lib.RMSD_ALIGN if align else 0) | (_lib.RMSD_PAIRED if paired else 0)
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().mlp_rmsd_eval(x.data_ptr(), B, N, x.stride(0), x.stride(2), x.stride(1),
                                            float(scale), float(shift), targets.data_ptr(), K, out.data_ptr(),
                                            flags, torch.cuda.current_stream(x.device).cuda_stream))
    return out


BOND_TYPES = ("N-CA", "CA-C", "C-N", "CA-CB")


def backbone_indices(atominfo):
    """Atom index lists the reference builds from ``self.mol.data`` (analyser.py:357-373, 423-449): per residue (in
    order of first appearance of its resid) the first N / CA / C / CB; ``C-N`` links a residue to resid+1.
    Returns ``(bonds: {type: [(i, j), ...]}, quads: [(N, CA, C, CB), ...])`` - CA-CB and the chirality quadruples
    skip GLY; types whose atoms are not selected are absent."""
    names = [str(a) for a in atominfo[:, 0]]
    resn = [str(a) for a in atominfo[:, 1]]
    resid = [int(a) for a in atominfo[:, 2]]
    atoms = set(names)
    if not {"CA", "C", "N"}.issubset(atoms):
        raise ValueError("Selected atoms should contain at least N, CA, and C.")
    first = {}
    order = []
    for i, (a, r) in enumerate(zip(names, resid)):
        if r not in first:
            first[r] = {"resname": resn[i]}
            order.append(r)
        first[r].setdefault(a, i)
    bonds = {"N-CA": [], "CA-C": [], "C-N": []}
    if "CB" in atoms:
        bonds["CA-CB"] = []
    quads = []
    last = max(order)
    for r in order:
        d = first[r]
        bonds["N-CA"].append((d["N"], d["CA"]))
        bonds["CA-C"].append((d["CA"], d["C"]))
        if d["resname"] != "GLY" and "CB" in atoms:
            bonds["CA-CB"].append((d["CA"], d["CB"]))
            quads.append((d["N"], d["CA"], d["C"], d["CB"]))
        if r != last:
            bonds["C-N"].append((d["C"], first[r + 1]["N"]))
    return bonds, quads


class GeometryStats:
    """Device tables + launcher of ``mlp_geometry_eval`` for one atom selection."""

    def __init__(self, atominfo, device):
        import numpy as np
        self.bonds, self.quads = backbone_indices(atominfo)
        self.types = [t for t in BOND_TYPES if t in self.bonds]
        pairs = np.concatenate([np.asarray(self.bonds[t], dtype=np.int32).reshape(-1, 2) for t in self.types])
        group = np.concatenate([np.full(len(self.bonds[t]), g, dtype=np.int32) for g, t in enumerate(self.types)])
        self.pairs = torch.from_numpy(pairs).to(device)
        self.group = torch.from_numpy(group).to(device)
        self.quad_t = torch.from_numpy(np.asarray(self.quads, dtype=np.int32).reshape(-1, 4)).to(device)

    def __call__(self, x, scale: float = 1.0):
        """x [B, N, 3] CUDA float32 (any strides) -> [B, 2*G+1]: (mean, std) per bond type, then inversions."""
        _check_structures(x)
        if x.device != self.pairs.device:
            raise ValueError(f"x lives on {x.device}, the geometry tables on {self.pairs.device}")
        G = len(self.types)
        out = torch.empty((x.shape[0], 2 * G + 1), dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            _lib.check(_lib.lib().mlp_geometry_eval(
                x.data_ptr(), x.shape[0], x.stride(0), x.stride(2), x.stride(1), float(scale),
                self.pairs.data_ptr(), self.group.data_ptr(), self.pairs.shape[0], G,
                self.quad_t.data_ptr() if len(self.quads) else None, len(self.quads), out.data_ptr(),
                torch.cuda.current_stream(x.device).cuda_stream))
        return out


class InternalCoords:
    """Device tables + launcher of ``mlp_internal_coords_eval``: named groups of atom pairs (distances), triples
    (angles) and quadruples (dihedrals) evaluated for every structure of a batch on the GPU.

    ``groups``: ``{name: int array [n, 2 | 3 | 4]}``.  ``__call__(x [B,N,3], scale)`` returns ``{name: [B, n]}``
    float32 CUDA tensors (views of one output buffer; distances are multiplied by ``scale``)."""

    def __init__(self, groups, device):
        import numpy as np
        self.device = torch.device(device)
        self.slices = {}
        parts = {2: [], 3: [], 4: []}
        for name, idx in groups.items():
            a = np.asarray(idx, dtype=np.int32).reshape(len(idx), -1)
            if a.shape[1] not in parts:
                raise ValueError(f"group {name}: expected index rows of 2, 3 or 4 atoms, got {a.shape}")
            if a.size and a.min() < 0:
                raise ValueError(f"group {name}: negative atom index")
            parts[a.shape[1]].append((name, a))
        self.counts = {}
        self.tables = {}
        off = 0
        for k in (2, 3, 4):
            rows = 0
            for name, a in parts[k]:
                self.slices[name] = (off + rows, off + rows + len(a))
                rows += len(a)
            self.counts[k] = rows
            cat = np.concatenate([a for _, a in parts[k]]) if rows else np.zeros((0, k), np.int32)
            self.tables[k] = torch.from_numpy(np.ascontiguousarray(cat)).to(self.device)
            off += rows
        self.n_all = off
        self.max_index = max([int(t.max()) for t in self.tables.values() if t.numel()] + [-1])

    def __call__(self, x, scale: float = 1.0):
        _check_structures(x)
        if x.device != self.device:
            raise ValueError(f"x lives on {x.device}, the index tables on {self.device}")
        if x.shape[1] <= self.max_index:
            raise ValueError(f"structures have {x.shape[1]} atoms, the index lists address atom {self.max_index}")
        B = x.shape[0]
        out = torch.empty((B, self.n_all), dtype=torch.float32, device=x.device)
        L = _lib.lib()
        with torch.cuda.device(x.device):
            for lo in range(0, B, 32768):
                xs, os_ = x[lo:lo + 32768], out[lo:lo + 32768]
                _lib.check(L.mlp_internal_coords_eval(
                    xs.data_ptr(), xs.shape[0], xs.stride(0), xs.stride(2), xs.stride(1), float(scale),
                    self.tables[2].data_ptr() if self.counts[2] else None, self.counts[2],
                    self.tables[3].data_ptr() if self.counts[3] else None, self.counts[3],
                    self.tables[4].data_ptr() if self.counts[4] else None, self.counts[4],
                    os_.data_ptr(), torch.cuda.current_stream(x.device).cuda_stream))
        return {name: out[:, a:b] for name, (a, b) in self.slices.items()}
