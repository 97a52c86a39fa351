"""vgsim_b200 — B200-native forward simulation + genealogy for the VGsim `Simulator` API."""
from ._interface import Simulator

__version__ = "0.1.0"
__all__ = ["Simulator"]
