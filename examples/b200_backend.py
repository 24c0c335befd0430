"""
The file a maintainer of kgori/phylo_utils would add (as ``phylo_utils/b200_backend.py``) to keep the reference's own
``TreeModel`` and only swap the arithmetic of its likelihood path for libphylo_b200.so - the ctypes stub of
INTEGRATION.md section 2, kept here as real code so that tests/test_integration_stub.py can execute it:

    tm = phylo_utils.tree_model.TreeModel(); tm.set_tree(...); tm.set_alignment(...); tm.set_rate_model(...)
    tm.set_substitution_model(...)                       # the reference object, configured as always
    be = B200Backend(tm)                                 # instead of tm.initialise()                tree_model.py:101-158
    be.compute_partials()                                # instead of tm.compute_partials()          tree_model.py:160-176
    site_lnl = be.compute_likelihood_at_edge(a, b)       # instead of tm.compute_likelihood_at_edge  tree_model.py:200-217

It reads nothing but public attributes of the reference objects (``alignment``, ``names``, ``inverse_index``,
``traversal.{names, postorder_traversal, brlens}``, ``rate_model.{ncat, rates, weights}``,
``substitution_model.{eigen, freqs}``) and imports nothing from phylo_utils_b200: numpy + ctypes + the shared library.
Every call is declared in include/phylo_b200.h next to the reference lines it replaces.
"""
import ctypes
import os

import numpy as np

_d = ctypes.POINTER(ctypes.c_double)
_i = ctypes.POINTER(ctypes.c_int32)
_DEFAULT_LIB = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "phylo_utils_b200", "libphylo_b200.so")


def load_library(path=None):
    lib = ctypes.CDLL(path or os.environ.get("PHB_LIBRARY") or _DEFAULT_LIB)
    lib.phb_create.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_uint,
                               ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.POINTER(ctypes.c_void_p)]
    lib.phb_destroy.argtypes = [ctypes.c_void_p]
    lib.phb_set_tips.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, _d, _i]
    lib.phb_set_model.argtypes = [ctypes.c_void_p, _d, _d, _d, _d, _d, _d]
    lib.phb_set_schedule.argtypes = [ctypes.c_void_p, ctypes.c_int, _i, ctypes.c_int, _i]
    lib.phb_set_edge_lengths.argtypes = [ctypes.c_void_p, _d]
    lib.phb_build_pmatrices.argtypes = [ctypes.c_void_p]
    lib.phb_compute_partials.argtypes = [ctypes.c_void_p, ctypes.c_int]
    lib.phb_root_lnl.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_double, _d, _d, _d, _d]
    lib.phb_last_error.argtypes = [ctypes.c_void_p]
    lib.phb_last_error.restype = ctypes.c_char_p
    return lib


class B200Backend(object):
    """Replaces TreeModel.initialise / compute_partials / compute_likelihood_at_edge (tree_model.py:101-217)."""

    def __init__(self, tm, lib=None, device=0):
        self.tm, self.lib = tm, lib if lib is not None else load_library()
        ntax, npat, A = tm.alignment.shape
        K = tm.rate_model.ncat
        self.ctx = ctypes.c_void_p()
        self._ok(self.lib.phb_create(device, ntax, npat, K, A, 0, None, 0, None, ctypes.byref(self.ctx)))
        # tips: the reference's 0/1 rows -> one code per (taxon, pattern) + the table of distinct rows (tree_model.py:142-148)
        rows, codes = np.unique(tm.alignment.reshape(-1, A), axis=0, return_inverse=True)
        codes = np.ascontiguousarray(np.asarray(codes).reshape(ntax, npat).astype(np.uint8))
        order = sorted(tm.traversal.names, key=lambda name: tm.names[name])          # alignment row order
        tip_nodes = np.array([tm.traversal.names[name] for name in order], dtype=np.int32)
        lut = np.ascontiguousarray(rows, dtype=np.double)
        self._ok(self.lib.phb_set_tips(self.ctx, codes.ctypes.data_as(ctypes.c_void_p), 0, len(lut), lut.ctypes.data_as(_d),
                                       tip_nodes.ctypes.data_as(_i)))
        model, rate = tm.substitution_model, tm.rate_model
        keep = [np.ascontiguousarray(a, dtype=np.double) for a in
                (model.eigen.evecs, model.eigen.evals, model.eigen.ivecs, model.freqs, rate.rates, rate.weights)]
        self._ok(self.lib.phb_set_model(self.ctx, *[a.ctypes.data_as(_d) for a in keep]))
        self.rows = np.ascontiguousarray(tm.traversal.postorder_traversal, dtype=np.int32)      # utils.py:127-134
        self._ok(self.lib.phb_set_schedule(self.ctx, len(self.rows), self.rows.ctypes.data_as(_i), 0, None))

    def _ok(self, status):
        if status:
            message = self.lib.phb_last_error(self.ctx if self.ctx else None)
            raise (ValueError if status in (1, 6) else RuntimeError)(message.decode() if message else "status {}".format(status))

    def compute_partials(self):                                            # tree_model.py:160-176
        br = self.tm.traversal.brlens
        lens = np.array([[br[(p, a)], br[(p, b)]] for p, a, b in self.rows.tolist()], dtype=np.double)
        self._ok(self.lib.phb_set_edge_lengths(self.ctx, lens.ctypes.data_as(_d)))
        self._ok(self.lib.phb_build_pmatrices(self.ctx))                   # replaces the 2(N-2) model.p calls (:168-169)
        self._ok(self.lib.phb_compute_partials(self.ctx, 0))               # replaces the N-2 clv calls (:176)

    def compute_likelihood_at_edge(self, node_a, node_b):                  # tree_model.py:200-217
        try:
            length = self.tm.traversal.brlens[node_a, node_b]
        except KeyError:
            raise ValueError('There is no edge connecting nodes {} and {}'.format(node_a, node_b))   # as tree_model.py:184-187
        total = ctypes.c_double()
        pattern = np.empty(self.tm.alignment.shape[1])
        self._ok(self.lib.phb_root_lnl(self.ctx, int(node_a), int(node_b), float(length), None, ctypes.byref(total),
                                       pattern.ctypes.data_as(_d), None))
        return pattern[self.tm.inverse_index]

    def close(self):
        if self.ctx:
            self.lib.phb_destroy(self.ctx)
            self.ctx = ctypes.c_void_p()
