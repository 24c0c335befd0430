"""
phylo_utils_b200 - B200-native tree-likelihood path with the API surface of kgori/phylo_utils.

Everything numerical runs in libphylo_b200.so (hand-written sm_100a CUDA behind the C ABI in
include/phylo_b200.h); importing the package loads that library and fails loudly if it has not
been built - there is no CPU fallback.
"""
from . import _lib

_lib.lib()  # fail at import time, not at first use, when the native library is missing

from . import discrete_gamma            # noqa: E402  (module, like the reference's compiled extension)
from .likelihood import numba_likelihood_engine, cuda_likelihood_engine   # noqa: E402
from . import substitution_models       # noqa: E402
from . import rate_models               # noqa: E402
from . import tree_model                # noqa: E402
from . import traversal, tree, utils, gamma, alignment, optimise   # noqa: E402
from .alignment.alignment import seq_to_partials          # noqa: E402
from .tree_model import TreeModel       # noqa: E402
from .engine import LikelihoodEngine    # noqa: E402

__version__ = "0.1.0"
