"""``physics_energy`` - per-conformation total energy as an ``nn.Module`` with the call contract of the
reference's ``openmm_energy`` (src/molearn/loss_functions/openmm_thread.py:388-419): ``forward(x)`` takes the
decoder's ``[B, n, 3]`` output in standardised units, rescales by ``std`` and returns ``[B]`` energies,
differentiable in ``x`` - so ``OpenMM_Physics_Trainer.common_physics_step``
(src/molearn/trainers/openmm_physics_trainer.py:111-131: ``energy = self.physics_loss(generated)``, then
inf -> 1e35, clamp, ``nanmean``) can use the sm_100a kernels without OpenMM or its plugin.

The energy is this package's ff14SB-style bond + angle + torsion + soft non-bonded sum, not OpenMM's full
force field: it replaces the dependency, not the numbers.
"""
from __future__ import annotations

import torch

from .torch_protein_energy import TorchProteinEnergy


class physics_energy(torch.nn.Module):
    def __init__(self, atominfo, std=1.0, mean=0.0, device=None, clamp=None, nonbonded=True, NB="repulsive", param_dir=None):
        super().__init__()
        self.std, self.mean = float(std), float(mean)
        self.clamp = clamp                      # {'min': .., 'max': ..} on the per-atom gradient, like the reference's clamp
        self.nonbonded = nonbonded
        self.energy = TorchProteinEnergy(None, atominfo, device=device, NB=NB, param_dir=param_dir)

    def forward(self, x):
        """x: [B, n, 3] float32 CUDA (any strides, e.g. ``decode(z)[:, :n, :]``) -> [B] energies (kcal/mol)."""
        xa = x * self.std + self.mean           # translation does not change the energy; kept for clarity of units
        if self.clamp is not None and xa.requires_grad:
            xa.register_hook(lambda g: g.clamp(**self.clamp))
        return self.energy.energy_per_conformation(xa.permute(0, 2, 1), nonbonded=self.nonbonded).sum(dim=1)
