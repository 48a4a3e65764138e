"""Infinite Gaussian mixture samplers on the B200 engine (mirror of pybgmm/igmm/__init__.py)."""
from .igmm import IGMM
from .crpmm import CRPMM
from .pcrpmm import PCRPMM
from .adapcrpmm import ADAPCRPMM
from .cscrpmm import CSCRPMM
from .subcrpmm import SubCRPMM

__all__ = ["IGMM", "CRPMM", "PCRPMM", "ADAPCRPMM", "CSCRPMM", "SubCRPMM"]
