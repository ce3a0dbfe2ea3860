"""hypha_b200 -- a Blackwell-native DiLoCo training runtime with Hypha's capabilities."""
__version__ = "0.1.0"
