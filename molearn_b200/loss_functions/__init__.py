from .torch_protein_energy import TorchProteinEnergy
from .physics_energy import physics_energy

__all__ = ["TorchProteinEnergy", "physics_energy"]
