from .folding_autoencoder import AutoEncoder
from .fused_decoder import FusedDecoder, fuse_decoder

__all__ = ["AutoEncoder", "FusedDecoder", "fuse_decoder"]
