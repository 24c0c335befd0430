"""
``phylo_utils_b200.likelihood`` - operator mirrors.  The reference's only wired engine is
``numba_likelihood_engine``; importing that name from here yields the CUDA implementation so
that ``from <pkg>.likelihood.numba_likelihood_engine import clv, lnl_node`` keeps working.
"""
import sys

from . import cuda_likelihood_engine
from .cuda_likelihood_engine import clv, lnl_node, lnl_branch, lnl_branch_derivs

from . import legacy
from .legacy import (Leaf, LnlNode, LnlModel, Mixture, GammaMixture, OptWrapper, BranchLengthOptimiser, optimise,
                     brent_optimise)

numba_likelihood_engine = cuda_likelihood_engine
sys.modules[__name__ + ".numba_likelihood_engine"] = cuda_likelihood_engine
