from .reducers import rmsd_to_targets

__all__ = ["rmsd_to_targets"]
