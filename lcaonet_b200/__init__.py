"""lcaonet_b200 — B200-native (sm_100a) implementation of the LCAONet interaction hot path behind the
reference's `LCAONet(...)` / `forward(batch)` API."""
from .keys import GraphKeys
from .model import LCAONet
from .orbitals import ElecInfo

__version__ = "0.1.0"
__all__ = ["LCAONet", "GraphKeys", "ElecInfo"]
