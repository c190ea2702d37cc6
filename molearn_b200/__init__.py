"""molearn_b200 - B200-native physics hot path of molearn (TorchProteinEnergy + latent-grid scan).

Importing the package does not load the CUDA extension; the first compute call does, and fails
loudly if ``molearn_b200/_C/libmolearn_b200.so`` is missing (``python -m molearn_b200.build``).
"""
__version__ = "0.1.0"

from .topology import Topology  # noqa: F401
from .loss_functions import TorchProteinEnergy  # noqa: F401

__all__ = ["Topology", "TorchProteinEnergy", "__version__"]
