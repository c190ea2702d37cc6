from .torch_protein_energy import TorchProteinEnergy

__all__ = ["TorchProteinEnergy"]
