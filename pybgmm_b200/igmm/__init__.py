"""Infinite Gaussian mixture samplers on the B200 engine: the classes of pybgmm/igmm/, each a host policy (scan order, count
prior power, constrained re-draws, mask moves) over the same device sweep."""
from .igmm import IGMM
from .crpmm import CRPMM
from .pcrpmm import PCRPMM
from .adapcrpmm import ADAPCRPMM
from .cscrpmm import CSCRPMM
from .subcrpmm import SubCRPMM

__all__ = ["IGMM", "CRPMM", "PCRPMM", "ADAPCRPMM", "CSCRPMM", "SubCRPMM"]
