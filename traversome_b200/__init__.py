"""traversome_b200 -- B200-native maximum-likelihood variant-frequency fitting for Traversome.

Drop-in for ONE hot path of JianjunJin/Traversome (v0.1.7.1): ``ModelFitMaxLike`` +
``PathMultinomialModel.get_like_formula`` and the bootstrap / backward-model-selection loops that
repeat them.  The arithmetic runs in hand-written sm_100a CUDA kernels behind the C ABI of
``include/tvs.h`` (``libtvs.so``); there is no CPU fallback.
"""
__version__ = "0.1.0"

from .utils import Bins, BinInfo, SubPathInfo, LogLikeFormulaInfo, LogLikeFuncInfo, Criterion, aic, bic  # noqa: F401
from .ModelGenerator import PathMultinomialModel  # noqa: F401


def install(module=None, **kw):
    """Swap this package's classes into the reference's orchestrator (see traversome_b200.dropin)."""
    from .dropin import install as _install
    return _install(module, **kw)


def uninstall():
    from .dropin import uninstall as _uninstall
    return _uninstall()
