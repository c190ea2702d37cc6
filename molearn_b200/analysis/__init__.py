from .reducers import rmsd_to_targets
from .analyser import MolearnAnalysis, shard_range

__all__ = ["MolearnAnalysis", "rmsd_to_targets", "shard_range"]
