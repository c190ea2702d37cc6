"""TEST / BASELINE INFRASTRUCTURE ONLY - PyTorch CPU port of the reference's energy path, used as
the timed ``cpu_baseline`` of ``bench.py`` (``kind: "port"``) and as a second fp32 checker.

The reference is pure Python on ATen ops and cannot travel to the GPU box, so its algorithm is
restated here with the same op structure and therefore the same cost profile:

* bonded terms follow the *roll* formulation (``_roll_bond_angle_torsion_loss``,
  src/molearn/loss_functions/torch_protein_energy.py:276-305): terms are grouped by their
  relative index pattern; every group is evaluated over the FULL ``[B,3,N]`` array with rolled
  copies of the coordinates and a zero-padded parameter row per group (force constant / barrier 0
  where no term starts), exactly the packing ``get_convolutions`` produces
  (torch_protein_energy_utils.py:749-811) - 2+6+11 groups for MurD backbone+CB.
* the non-bonded term is ``_cdist_nb`` / ``_cdist_nb_full`` verbatim in structure (:224-239):
  ``torch.cdist`` -> ``elu`` warps -> dense ``[B,N,N]`` divisions against dense ``[N,N]`` vdw_A /
  q1q2 matrices built as at utils.py:813-830 (pair exclusions, SURVEY NB-3) -> ``nansum``.

Never imported by the product.
"""
from __future__ import annotations

import numpy as np
import torch


class TorchCPUPort:
    def __init__(self, topo: dict, dtype=torch.float32, nb_full: bool = False, device="cpu"):
        """topo: arrays named as in tests/golden/topo_*.npz (or ``Topology.as_dict()``).
        ``device``: "cpu" (the timed CPU baseline) or a CUDA device - the same ATen op chain on the GPU, which is
        what the reference itself runs with ``device='cuda'`` (bench.py's ``torch_cuda_baseline``)."""
        self.dtype = dtype
        self.device = torch.device(device)
        self.nb_full = nb_full
        N = len(topo["atom_R"])
        self.N = N
        f32 = lambda a: torch.tensor(np.asarray(a, dtype=np.float32)).to(dtype)  # reference: .float() first

        def groups(idx, para_rows):
            """group terms by offsets relative to the lowest atom index -> [(offsets, rows[N,...])]"""
            idx = np.asarray(idx)
            lo = idx.min(axis=1)
            rel = idx - lo[:, None]
            out = []
            for pat in sorted({tuple(r) for r in rel}):
                sel = np.all(rel == np.array(pat), axis=1)
                rows = []
                for p, fill in para_rows:
                    row = np.full((N,) + p.shape[1:], fill, dtype=np.float64)
                    row[lo[sel]] = p[sel]
                    rows.append(f32(row))
                out.append((pat, rows))
            return out

        bp, ap, tp = (np.asarray(topo[k], dtype=np.float64) for k in ("bond_para", "angle_para", "torsion_para"))
        self.bond_groups = groups(topo["bond_idxs"], [(bp[:, 0], 0.0), (bp[:, 1], 0.0)])
        self.angle_groups = groups(topo["angle_idxs"], [(ap[:, 0], 0.0), (ap[:, 1], 0.0)])
        tpad = tp.copy()  # [T,4,P]; padding rows: factor 1, everything else 0 (utils.py:733-737)
        base = np.zeros((4, tp.shape[2]))
        base[0] = 1.0
        idx = np.asarray(topo["torsion_idxs"])
        lo = idx.min(axis=1)
        rel = idx - lo[:, None]
        self.torsion_groups = []
        for pat in sorted({tuple(r) for r in rel}):
            sel = np.all(rel == np.array(pat), axis=1)
            row = np.broadcast_to(base, (N,) + base.shape).copy()
            row[lo[sel]] = tpad[sel]
            self.torsion_groups.append((pat, f32(row)))
        # dense NB matrices (utils.py:813-830; energy.py:125-127)
        R, e, q = f32(topo["atom_R"]).float(), f32(topo["atom_e"]).float(), f32(topo["atom_charges"]).float()
        vdw_R = 0.5 * torch.cdist(R.view(-1, 1), -R.view(-1, 1)).triu(diagonal=1)
        vdw_e = (e.view(1, -1) * e.view(-1, 1)).triu(diagonal=1).sqrt()
        q1q2 = (q.view(1, -1) * q.view(-1, 1)).triu(diagonal=1)
        ex = np.concatenate([np.asarray(topo["bond_idxs"]), np.asarray(topo["angle_idxs"])[:, (0, 2)],
                             np.asarray(topo["torsion_idxs"])[:, (0, 3)]]).astype(np.int64)
        for m in (vdw_R, vdw_e, q1q2):
            m[ex[:, 0], ex[:, 1]] = 0.0
        self.vdw_A = (vdw_e * vdw_R ** 12).to(dtype)
        self.vdw_B = (2 * vdw_e * vdw_R ** 6).to(dtype)
        self.q1q2 = q1q2.to(dtype)
        if self.device.type != "cpu":
            mv = lambda t: t.to(self.device)
            self.bond_groups = [(pat, [mv(r) for r in rows]) for pat, rows in self.bond_groups]
            self.angle_groups = [(pat, [mv(r) for r in rows]) for pat, rows in self.angle_groups]
            self.torsion_groups = [(pat, mv(row)) for pat, row in self.torsion_groups]
            self.vdw_A, self.vdw_B, self.q1q2 = mv(self.vdw_A), mv(self.vdw_B), mv(self.q1q2)

    # x: [B,3,N]
    def bonded(self, x):
        r = lambda o: x if o == 0 else x.roll(-o, 2)
        bl = x.new_zeros(())
        al = x.new_zeros(())
        tl = x.new_zeros(())
        for (oi, oj), (r0, k) in self.bond_groups:
            v = r(oi) - r(oj)
            bl = bl + (((v.norm(dim=1) - r0) ** 2) * k).sum()
        for (oi, oj, ok), (th0, k) in self.angle_groups:
            xj = r(oj)
            v1, v2 = r(oi) - xj, r(ok) - xj
            c = torch.sum(v1 * v2, dim=1) / (torch.norm(v1, dim=1) * torch.norm(v2, dim=1))
            th = torch.acos(torch.clamp(c, min=-0.999999, max=0.999999))
            al = al + (k * (th - th0) ** 2).sum()
        for (oi, oj, ok, ol), p in self.torsion_groups:
            xj, xk = r(oj), r(ok)
            b1, b2, b3 = r(oi) - xj, xj - xk, r(ol) - xk
            c32 = torch.linalg.cross(b3, b2, dim=1)
            c12 = torch.linalg.cross(b1, b2, dim=1)
            phi = torch.atan2((b2 * torch.linalg.cross(c32, c12, dim=1)).sum(dim=1),
                              b2.norm(dim=1) * ((c12 * c32).sum(dim=1)))
            pp = p.unsqueeze(0)
            tl = tl + ((pp[:, :, 1] / pp[:, :, 0]) * (1 + torch.cos(pp[:, :, 3] * phi.unsqueeze(2) - pp[:, :, 2]))).sum()
        return bl, al, tl

    @staticmethod
    def _warp(d, k):
        return torch.nn.functional.elu(d - k, 1.0) + k

    def nonbonded(self, x):
        xt = x.permute(0, 2, 1)
        d = torch.cdist(xt, xt)
        if self.nb_full:
            d6 = self._warp(d, 1.9) ** 6
            return torch.nansum(self.vdw_A / (d6 ** 2) - self.vdw_B / d6 + self.q1q2 / self._warp(d, 0.4))
        return torch.nansum(self.vdw_A / (self._warp(d, 1.9) ** 12) + self.q1q2 / self._warp(d, 0.4))

    def energy_and_grad(self, x, nonbonded=True):
        """One forward + backward of the sum of all terms; returns ([4] batch sums, grad [B,3,N])."""
        x = x.detach().to(self.device, self.dtype).requires_grad_(True)
        terms = list(self.bonded(x))
        if nonbonded:
            terms.append(self.nonbonded(x))
        sum(terms).backward()
        return torch.stack([t.detach() for t in terms]), x.grad
