"""``Torch_Physics_Trainer`` on the sm_100a energy kernels.

Keeps the consumer API of the reference trainer (src/molearn/trainers/torch_physics_trainer.py:
``prepare_physics``, ``common_physics_step``, ``train_step``, ``valid_step``, result keys
``physics_loss, bond_energy, angle_energy, torsion_energy, mse_loss, loss``) and its loss mixing
(train: ``mse + sg(psf*mse/(phys+1e-5)) * phys``, :62-64; valid: ``log(mse) + psf*log(phys)``, :84).

The body is written for the current ``[B, N, 3]`` data layout (``Trainer.common_step``,
trainers/trainer.py:556-571) following the maintained sibling ``OpenMM_Physics_Trainer.
common_physics_step`` (openmm_physics_trainer.py:111-131).  The reference's own body still indexes
``[B, 3, N]`` and rebuilds the three energies with ``torch.tensor([...])``, which detaches them from
the graph; here the physics gradient does reach the decoder.

The epoch loop, logging and checkpointing of the reference ``Trainer`` (trainers/trainer.py) are
host-side code outside the hot path.  If molearn is importable this class can be mixed into its
``Trainer``; stand-alone, the small ``Trainer`` below provides the hooks the step functions use.
"""
from __future__ import annotations

import torch

from ..loss_functions import TorchProteinEnergy


class Trainer:
    """The slice of ``molearn.trainers.Trainer`` the physics trainer relies on
    (set_autoencoder :308-318, set_data :330-350, prepare_optimiser, train/valid steps :540-607)."""

    def __init__(self, device=None, **_):
        self.device = torch.device(device) if device is not None else torch.device(
            "cuda" if torch.cuda.is_available() else "cpu")
        self.epoch = 0
        self._internal = {}

    def set_autoencoder(self, autoencoder, **kwargs):
        self.autoencoder = (autoencoder(**kwargs) if isinstance(autoencoder, type) else autoencoder).to(self.device)
        self._autoencoder_kwargs = kwargs

    def set_data(self, data, **_):
        """``data``: PDBData-like (``dataset`` [F,N,3], ``std``, ``mean``, ``get_atominfo()``)."""
        if getattr(data, "dataset", None) is None:
            data.prepare_dataset()
        self._data = data
        self.std, self.mean = float(data.std), float(data.mean)
        self.mol = None

    def prepare_optimiser(self, lr=1e-3, weight_decay=0.0001, **kwargs):
        self.optimiser = torch.optim.AdamW(self.autoencoder.parameters(), lr=lr, weight_decay=weight_decay, **kwargs)

    def common_step(self, batch):
        self._internal = {}
        encoded = self.autoencoder.encode(batch)
        self._internal["encoded"] = encoded
        decoded = self.autoencoder.decode(encoded)[:, :batch.size(1), :]
        self._internal["decoded"] = decoded
        return dict(mse_loss=((batch - decoded) ** 2).mean())

    def train_step(self, batch):
        results = self.common_step(batch)
        results["loss"] = results["mse_loss"]
        return results

    valid_step = train_step

    def training_iteration(self, batch):
        """One optimiser update on a device batch (the body of ``_run_phase``, trainer.py:107-141)."""
        self.autoencoder.train()
        self.optimiser.zero_grad(set_to_none=True)
        results = self.train_step(batch)
        results["loss"].backward()
        self.optimiser.step()
        return results


class Torch_Physics_Trainer(Trainer):
    def prepare_physics(self, physics_scaling_factor=0.1, NB="repulsive", param_dir=None):
        """Create ``self.physics_loss`` (torch_physics_trainer.py:15-22)."""
        self.psf = physics_scaling_factor
        frame = (self._data.dataset[0] * self.std).T                  # [3, N] as the constructor documents
        self.physics_loss = TorchProteinEnergy(frame, pdb_atom_names=self._data.get_atominfo(),
                                               device=self.device, method="roll", NB=NB, param_dir=param_dir)

    def common_physics_step(self, batch, latent):
        """Decode random interpolations between adjacent latent vectors and score them
        (torch_physics_trainer.py:24-45).  ``batch`` [B, N, 3]; ``latent`` [B, 2]."""
        alpha = torch.rand(int(len(batch) // 2), 1).type_as(latent)
        latent_interpolated = (1 - alpha) * latent[:-1:2] + alpha * latent[1::2]
        generated = self.autoencoder.decode(latent_interpolated)[:, :batch.size(1), :]
        self._internal["generated"] = generated
        # decoder output is a strided view; the kernels take it as is ([B,3,N] logical view, Angstrom)
        x = (generated * self.std + self.mean).permute(0, 2, 1)
        bond, angle, torsion = self.physics_loss._roll_bond_angle_torsion_loss(x)
        n = len(generated)
        bond, angle, torsion = bond / n, angle / n, torsion / n
        _all = torch.stack([bond, angle, torsion])
        _all = torch.where(_all.isinf(), torch.full_like(_all, 1e35), _all)
        total_physics = _all.nansum()
        return {"physics_loss": total_physics, "bond_energy": bond, "angle_energy": angle, "torsion_energy": torsion}

    def train_step(self, batch):
        results = self.common_step(batch)
        results.update(self.common_physics_step(batch, self._internal["encoded"]))
        with torch.no_grad():
            scale = self.psf * results["mse_loss"] / (results["physics_loss"] + 1e-5)
        results["loss"] = results["mse_loss"] + scale * results["physics_loss"]
        return results

    def valid_step(self, batch):
        results = self.common_step(batch)
        results.update(self.common_physics_step(batch, self._internal["encoded"]))
        results["loss"] = torch.log(results["mse_loss"]) + self.psf * torch.log(results["physics_loss"])
        return results
