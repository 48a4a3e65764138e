"""pybgmm_b200 -- B200-native collapsed-Gibbs sampler for the CRP / pCRP Gaussian mixture model behind the class
surface of junlulocky/PyBGMM (NIW, FixedVarPrior, GaussianComponents{,Diag,FixedVar}, IGMM / CRPMM / PCRPMM / ADAPCRPMM /
CSCRPMM / SubCRPMM).

The package layout mirrors the reference's (`pybgmm.prior`, `pybgmm.gaussian`, `pybgmm.gmm`, `pybgmm.igmm`,
`pybgmm.utils`), so switching is a change of the top-level package name.  All arithmetic of the hot path runs in
libbgmm_b200.so (hand-written sm_100a CUDA); there is no CPU fallback.
"""
from .prior import NIW, BetaBern
from .gaussian import GaussianComponents, GaussianComponentsDiag, GaussianComponentsFixedVar, FixedVarPrior
from .gmm import GMM
from .igmm import IGMM, CRPMM, PCRPMM, ADAPCRPMM, CSCRPMM, SubCRPMM

__all__ = ["NIW", "BetaBern", "FixedVarPrior", "GaussianComponents", "GaussianComponentsDiag", "GaussianComponentsFixedVar", "GMM",
           "IGMM", "CRPMM", "PCRPMM", "ADAPCRPMM", "CSCRPMM", "SubCRPMM"]
__version__ = "0.1.0"
