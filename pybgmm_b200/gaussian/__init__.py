from .gaussian_components import GaussianComponents
from .gaussian_components_diag import GaussianComponentsDiag
from .gaussian_components_fixedvar import GaussianComponentsFixedVar, FixedVarPrior

__all__ = ["GaussianComponents", "GaussianComponentsDiag", "GaussianComponentsFixedVar", "FixedVarPrior"]
