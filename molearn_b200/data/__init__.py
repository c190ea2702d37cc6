from .pdb_reader import read_pdb_frames, read_dcd, select_atoms, CoordinateSet

__all__ = ["read_pdb_frames", "read_dcd", "select_atoms", "CoordinateSet"]
