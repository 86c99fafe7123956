"""B200-native radar pillarization hot path for HGSFusion (points -> pillars -> PillarVFE -> BEV canvas)."""
__version__ = "0.1.0"
