"""Synthetic multi-chain systems for the large-N benchmark configurations.

A tile is the MurD backbone+CB selection with the last residue's ``C`` and ``O`` removed: the reference
links a residue's ``N`` to any ``C`` of the previous residue on every resid change without a chain or
distance test (torch_protein_energy_utils.py:588-598), so dropping that ``C`` is what makes consecutive
tiles separate chains.  Resids run consecutively across tiles; tile ``t`` is translated on an 80 A lattice.
"""
from __future__ import annotations

import numpy as np


def tile_system(atominfo, coords, n_tiles, frame_of_tile=lambda t, b: (b + t) % 16, batch=1, noise=0.0, seed=1234):
    """atominfo [N,3] object, coords [F,N,3] -> (tiled atominfo [n_tiles*(N-2),3], x [batch, N', 3] float32)."""
    last = atominfo[-1, 2]
    keep = np.array([not (r == last and a in ("C", "O")) for a, _, r in atominfo])
    tinfo, tcoords = atominfo[keep], coords[:, keep]
    nres = int(tinfo[-1, 2]) - int(tinfo[0, 2]) + 1
    infos = []
    for t in range(n_tiles):
        ti = tinfo.copy()
        ti[:, 2] = [int(r) + t * nres for r in tinfo[:, 2]]
        infos.append(ti)
    n = tcoords.shape[1]
    x = np.empty((batch, n_tiles * n, 3), dtype=np.float32)
    for t in range(n_tiles):
        shift = 80.0 * np.array([t % 3, (t // 3) % 3, t // 9], dtype=np.float32)
        for b in range(batch):
            x[b, t * n:(t + 1) * n] = tcoords[frame_of_tile(t, b) % tcoords.shape[0]] + shift
    if noise:
        import torch
        g = torch.Generator().manual_seed(seed)
        x += noise * torch.randn(x.shape, generator=g).numpy()
    return np.concatenate(infos), x
