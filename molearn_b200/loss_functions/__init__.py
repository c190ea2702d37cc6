from .torch_protein_energy import TorchProteinEnergy
from .physics_energy import physics_energy
from .graphed import GraphedEnergyStep

__all__ = ["TorchProteinEnergy", "physics_energy", "GraphedEnergyStep"]
