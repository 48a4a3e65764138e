"""Hyper-parameter structs of the priors the hot path uses (the reference keeps more under pybgmm/prior/: CRP, Wishart, ...
-- those are not on the path): the normal-inverse-Wishart of the components, the Beta prior of SubCRPMM's mask."""
from .betabern import BetaBern
from .niw import NIW

__all__ = ["NIW", "BetaBern"]
