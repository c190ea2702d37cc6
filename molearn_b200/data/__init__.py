from .pdb_reader import read_pdb_frames, read_dcd, select_atoms, CoordinateSet
from .synthetic import tile_system

__all__ = ["read_pdb_frames", "read_dcd", "select_atoms", "CoordinateSet", "tile_system"]
