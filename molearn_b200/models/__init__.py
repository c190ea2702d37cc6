from .folding_autoencoder import AutoEncoder

__all__ = ["AutoEncoder"]
