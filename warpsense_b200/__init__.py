"""warpsense-b200: B200-native TSDF update + Point-to-TSDF registration."""
__version__ = "0.1.0"
