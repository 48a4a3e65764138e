from .gaussian_components import GaussianComponents
from .gaussian_components_diag import GaussianComponentsDiag

__all__ = ["GaussianComponents", "GaussianComponentsDiag"]
