"""
ORACLE - TEST INFRASTRUCTURE ONLY.  Runs the UNMODIFIED reference package from /root/reference in the
build container (it cannot travel to the GPU box; its outputs do, as tests/golden/*.npz).

The reference does not import as shipped in this image; four shims make it:
  1. ``phylo_utils/__init__.py`` imports compiled extensions -> register a bare package object whose
     ``__path__`` points at the reference directory, and provide ``phylo_utils.discrete_gamma`` on top of
     the reference's own C file compiled into oracle/_ref (oracle/Makefile);
  2. ``np.int`` was removed from numpy -> alias it to ``int`` before importing;
  3. Biopython is absent -> stub ``Bio``, ``Bio.AlignIO``, ``Bio.Alphabet.IUPAC``;
  4. dendropy is absent -> the reference only duck-types trees; phylo_utils_b200.tree.Tree has the
     required surface.
Nothing here is used by the product.
"""
import os
import sys
import types

import numpy as np

REF_ROOT = "/root/reference"


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "phylo_utils"))


def load_reference():
    """-> the reference's ``phylo_utils`` package (modules import lazily from /root/reference)."""
    if "phylo_utils" in sys.modules and getattr(sys.modules["phylo_utils"], "_is_reference_shim", False):
        return sys.modules["phylo_utils"]
    if not available():
        raise RuntimeError("reference not mounted at " + REF_ROOT)
    if not hasattr(np, "int"):
        np.int = int                                            # shim 2
    bio = types.ModuleType("Bio")                               # shim 3
    bio.AlignIO = types.ModuleType("Bio.AlignIO")
    alpha = types.ModuleType("Bio.Alphabet")
    iupac = types.ModuleType("Bio.Alphabet.IUPAC")
    iupac.IUPACAmbiguousDNA = type("IUPACAmbiguousDNA", (), {"letters": "GATCRYWSMKHBVDN"})
    alpha.IUPAC = iupac
    bio.Alphabet = alpha
    sys.modules.setdefault("Bio", bio)
    sys.modules.setdefault("Bio.AlignIO", bio.AlignIO)
    sys.modules.setdefault("Bio.Alphabet", alpha)
    sys.modules.setdefault("Bio.Alphabet.IUPAC", iupac)

    pkg = types.ModuleType("phylo_utils")                       # shim 1
    pkg.__path__ = [os.path.join(REF_ROOT, "phylo_utils")]
    pkg._is_reference_shim = True
    sys.modules["phylo_utils"] = pkg
    from . import oracle as _oracle
    dg = types.ModuleType("phylo_utils.discrete_gamma")
    dg.discrete_gamma = lambda alpha, ncat, median_rates=False: _oracle.ref_discrete_gamma(alpha, ncat, median_rates)
    sys.modules["phylo_utils.discrete_gamma"] = dg
    pkg.discrete_gamma = dg

    import importlib
    for name in ("substitution_models", "rate_models", "traversal", "tree_model", "gamma"):
        setattr(pkg, name, importlib.import_module("phylo_utils." + name))
    pkg.alignment_module = importlib.import_module("phylo_utils.alignment.alignment")
    pkg.engine = importlib.import_module("phylo_utils.likelihood.numba_likelihood_engine")
    return pkg


class Record(object):
    """Duck-typed alignment record (alignment.py:43-45 reads .seq and .name)."""

    def __init__(self, name, seq):
        self.name, self.seq, self.id = name, seq, name
