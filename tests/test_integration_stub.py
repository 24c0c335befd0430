"""
The INTEGRATION.md binding (examples/b200_backend.py) - the file a maintainer of the reference would add - executed.

The reference checkout and a GPU never meet (the reference cannot travel to the GPU box, this container has no GPU),
so the check has two halves that share one recorded call sequence:

  * here, with the reference mounted (marker `reference`): the stub is plugged into the reference's OWN TreeModel
    (loaded unmodified through oracle/ref_shims.py) with a recording stand-in for the library, and every argument it
    would pass over the C ABI - codes, look-up table, tip node ids, eigen-system, schedule, branch lengths, root edge -
    is compared with what the same stub produces from a REPLICA TreeModel assembled from phylo_utils_b200's host objects;
  * on the GPU box (marker `gpu`): the same stub drives the real library from that replica, and its per-site lnL is
    compared with the output of the unmodified reference (tests/golden/).
"""
import ctypes
import importlib.util
import os
import types

import numpy as np
import pytest

import phylo_utils_b200 as phy
from helpers import load, problem, records, tree, assert_lnl_close

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _stub_module():
    spec = importlib.util.spec_from_file_location("b200_backend", os.path.join(ROOT, "examples", "b200_backend.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def replica_tree_model(name):
    """An object with the reference TreeModel's public attributes, assembled from this package's host-side classes."""
    g, tr, codes, lut, sw, ii, names, model, rate = problem(name)
    tm = types.SimpleNamespace()
    tm.alignment = np.ascontiguousarray(lut[codes])           # (ntax, npat, A) 0/1 floats, as alignment_to_numpy returns
    tm.names, tm.inverse_index, tm.siteweights = names, ii, sw
    tm.traversal, tm.rate_model, tm.substitution_model = tr, rate, model
    return g, tm


class RecordingLibrary(object):
    """Stands in for libphylo_b200.so: every phb_* call returns PHB_OK and its array arguments are copied out."""

    def __init__(self, shapes):
        self.calls, self.shapes = [], shapes

    def __getattr__(self, fn):
        if not fn.startswith("phb_"):
            raise AttributeError(fn)

        def call(*args):
            got = []
            spec = self.shapes.get(fn, {})
            for pos, arg in enumerate(args):
                if pos in spec:
                    dtype, count = spec[pos]
                    n = count(self, args) if callable(count) else count
                    addr = ctypes.cast(arg, ctypes.c_void_p).value
                    got.append(np.ctypeslib.as_array((ctypes.c_uint8 * (n * np.dtype(dtype).itemsize)).from_address(addr)).view(dtype).copy())
                elif isinstance(arg, (int, float)):
                    got.append(arg)
            self.calls.append((fn, got))
            if fn == "phb_create":
                self.dims = args[1:5]                          # n_tips, n_patterns, n_cat, n_states
            return 0
        return call


def _recorded(tm):
    n = lambda f: (lambda self, a: f(*self.dims))              # noqa: E731
    shapes = {
        "phb_set_tips": {1: (np.uint8, n(lambda t, s, k, a: t * s)), 4: (np.double, lambda self, a: a[3] * self.dims[3]),
                         5: (np.int32, n(lambda t, s, k, a: t))},
        "phb_set_model": {1: (np.double, n(lambda t, s, k, a: a * a)), 2: (np.double, n(lambda t, s, k, a: a)),
                          3: (np.double, n(lambda t, s, k, a: a * a)), 4: (np.double, n(lambda t, s, k, a: a)),
                          5: (np.double, n(lambda t, s, k, a: k)), 6: (np.double, n(lambda t, s, k, a: k))},
        "phb_set_schedule": {2: (np.int32, lambda self, a: 3 * a[1])},
        "phb_set_edge_lengths": {1: (np.double, n(lambda t, s, k, a: 2 * (t - 2)))},
    }
    lib = RecordingLibrary(shapes)
    be = _stub_module().B200Backend(tm, lib=lib)
    be.compute_partials()
    a, b = tm.traversal.root_edge
    be.compute_likelihood_at_edge(a, b)
    return lib.calls


@pytest.mark.reference
@pytest.mark.parametrize("name", ["cfg1_gtr_g4", "ambig_hky_ig"])
def test_stub_in_the_reference_tree_model_marshals_what_the_replica_marshals(name):
    from oracle import ref_shims
    ref = ref_shims.load_reference()
    g, replica = replica_tree_model(name)
    alphabet = int(g["alphabet"])
    tm = ref.tree_model.TreeModel()                                        # the reference's own class, unmodified
    tm.set_tree(tree(g))
    tm.set_alignment([ref_shims.Record(str(n), bytes(row).decode("ascii")) for n, row in zip(g["names"], g["seqs"])], alphabet)
    tm.set_rate_model(replica.rate_model)
    tm.set_substitution_model(replica.substitution_model)
    ref_calls, rep_calls = _recorded(tm), _recorded(replica)
    assert [c[0] for c in ref_calls] == [c[0] for c in rep_calls] == [
        "phb_create", "phb_set_tips", "phb_set_model", "phb_set_schedule", "phb_set_edge_lengths", "phb_build_pmatrices",
        "phb_compute_partials", "phb_root_lnl"]
    for (fn, a), (_, b) in zip(ref_calls, rep_calls):
        assert len(a) == len(b), fn
        for x, y in zip(a, b):
            assert np.array_equal(np.asarray(x), np.asarray(y)), fn
    # and the reference's own evaluation of that TreeModel is what the golden file holds
    tm.initialise()
    tm.compute_partials()
    assert_lnl_close(tm.compute_likelihood_at_edge(*tm.traversal.root_edge), g["site_lnl"])


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["cfg1_gtr_g4", "ambig_hky_ig", "deep300_gtr_g4", "prot12_lg_g4"])
def test_stub_drives_the_real_library_to_the_reference_result(name):
    g, replica = replica_tree_model(name)
    be = _stub_module().B200Backend(replica)
    be.compute_partials()
    site = be.compute_likelihood_at_edge(*replica.traversal.root_edge)
    assert_lnl_close(site, g["site_lnl"], what=name + " per-site lnL through the INTEGRATION.md binding")
    with pytest.raises(ValueError):
        be.compute_likelihood_at_edge(0, 1)                                # no such edge: the reference's ValueError
    be.close()
