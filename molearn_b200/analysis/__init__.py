from .reducers import rmsd_to_targets, GeometryStats, backbone_indices
from .analyser import MolearnAnalysis, shard_range

__all__ = ["MolearnAnalysis", "rmsd_to_targets", "GeometryStats", "backbone_indices", "shard_range"]
