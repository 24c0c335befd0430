"""
GPU parity at sizes beyond the golden files: CUDA path vs the CPU oracle on seeded synthetic inputs,
plus size-independent properties at (a slice of) the headline shape.
"""
import numpy as np
import pytest

import phylo_utils_b200 as phy
from phylo_utils_b200 import _lib
from phylo_utils_b200.tree import random_tree, caterpillar_tree, balanced_tree
from helpers import assert_lnl_close
from oracle import oracle

pytestmark = pytest.mark.gpu


def synthetic(n_taxa, n_pat, n_states, seed, tree_fn=random_tree, gap=0.01):
    rng = np.random.default_rng(seed)
    tree = tree_fn(n_taxa, seed)
    names = [l.taxon.label for l in tree.leaf_node_iter()]
    lut = np.vstack([np.eye(n_states)[::-1], np.ones((1, n_states))])      # lexicographic rank order, gap last
    codes = rng.integers(0, n_states, size=(n_taxa, n_pat)).astype(np.uint8)
    codes[rng.random((n_taxa, n_pat)) < gap] = n_states
    return tree, names, codes, lut


def run_both(tree, names, codes, lut, model, rate, mode="auto", weights=None):
    tm = phy.TreeModel(mode=mode)
    tm.set_tree(tree)
    tm.set_tip_codes(codes, lut, {n: i for i, n in enumerate(names)}, siteweights=weights)
    tm.set_rate_model(rate)
    tm.set_substitution_model(model)
    tm.initialise()
    a, b = tm.traversal.root_edge
    total, pattern = tm._pattern_lnl(a, b)
    tips = {tm.traversal.names[n]: np.ascontiguousarray(lut[codes[i]]) for i, n in enumerate(names)}
    want = oracle.tree_lnl(tm.traversal, tips, model.p, model.freqs, rate.rates, rate.weights)
    return tm, total, pattern, want


@pytest.mark.parametrize("tree_fn,n_taxa,n_pat,mode", [
    (random_tree, 200, 20000, "tile"),        # several tiles per CTA, tile mode with on-chip reuse
    (random_tree, 200, 20000, "level"),
    (caterpillar_tree, 150, 5000, "tile"),    # depth = n-2, every row chains on the previous one
    (balanced_tree, 128, 3000, "level"),
    (random_tree, 64, 33, "tile"),            # ragged: fewer patterns than one tile
    (random_tree, 5, 1, "level"),             # a single pattern
])
def test_dna_gtr_gamma_vs_oracle(tree_fn, n_taxa, n_pat, mode):
    tree, names, codes, lut = synthetic(n_taxa, n_pat, 4, seed=n_taxa + n_pat, tree_fn=tree_fn)
    model = phy.substitution_models.GTR([6., 5., 4., 3., 2., 1.], [0.1, 0.2, 0.3, 0.4])
    rate = phy.rate_models.GammaRateModel(4, 0.5)
    w = np.random.default_rng(1).integers(1, 5, size=n_pat)
    tm, total, pattern, want = run_both(tree, names, codes, lut, model, rate, mode, weights=w)
    assert_lnl_close(pattern, want)
    assert_lnl_close(total, float(np.dot(want, w)))


@pytest.mark.parametrize("K,rate", [(1, lambda: phy.rate_models.UniformRateModel()),
                                    (2, lambda: phy.rate_models.InvariantSitesModel(0.2)),
                                    (5, lambda: phy.rate_models.InvariantGammaModel(0.1, 4, 0.9)),
                                    (8, lambda: phy.rate_models.GammaRateModel(8, 0.4))])
def test_dna_other_category_counts(K, rate):
    tree, names, codes, lut = synthetic(60, 4000, 4, seed=K)
    model = phy.substitution_models.HKY85(2.0, [0.3, 0.2, 0.2, 0.3])
    for mode in ("tile", "level"):
        _, total, pattern, want = run_both(tree, names, codes, lut, model, rate(), mode)
        assert_lnl_close(pattern, want)


def test_protein_lg_gamma_vs_oracle():
    tree, names, codes, lut = synthetic(50, 2000, 20, seed=3)
    for mode in ("tile", "level"):
        _, total, pattern, want = run_both(tree, names, codes, lut, phy.substitution_models.LG(),
                                           phy.rate_models.GammaRateModel(4, 0.7), mode)
        assert_lnl_close(pattern, want)
        assert_lnl_close(total, want.sum())


def test_codon_gy94_gamma_vs_oracle():
    rng = np.random.default_rng(4)
    from phylo_utils_b200.substitution_models.codon import f3x4
    model = phy.substitution_models.GY94(2.0, 0.2, f3x4(rng.dirichlet(np.ones(4) * 5, size=3)))
    tree, names, codes, lut = synthetic(20, 500, 61, seed=4)
    for mode in ("tile", "level"):
        _, total, pattern, want = run_both(tree, names, codes, lut, model, phy.rate_models.GammaRateModel(4, 0.5), mode)
        assert_lnl_close(pattern, want)


def test_binary_states():
    tree, names, codes, lut = synthetic(30, 700, 2, seed=5)

    class Binary(phy.substitution_models.abstract.Model):
        _size = 2

        def __init__(self):
            from phylo_utils_b200.substitution_models.utils import compute_q_matrix, get_eigen
            self._freqs = np.array([0.3, 0.7])
            self._q_mtx = compute_q_matrix(np.array([[0., 1.], [1., 0.]]), self._freqs)
            self.eigen = phy.substitution_models.abstract.Eigen(*get_eigen(self._q_mtx, self._freqs))
    _, total, pattern, want = run_both(tree, names, codes, lut, Binary(), phy.rate_models.GammaRateModel(4, 1.0))
    assert_lnl_close(pattern, want)


def test_repeated_blocks_property_at_scale():
    """
    Size-independent property: an alignment made of R copies of a block of patterns has
    per-pattern lnL periodic in the block and total = R x block total.  Run at 1000 taxa with a
    pattern count large enough for many tiles per CTA; the block itself is checked against the oracle.
    """
    n_taxa, block, reps = 1000, 512, 64
    tree, names, codes, lut = synthetic(n_taxa, block, 4, seed=11)
    big = np.ascontiguousarray(np.tile(codes, (1, reps)))
    model = phy.substitution_models.GTR([6., 5., 4., 3., 2., 1.], [0.1, 0.2, 0.3, 0.4])
    rate = phy.rate_models.GammaRateModel(4, 0.5)
    tm = phy.TreeModel(mode="tile")
    tm.set_tree(tree)
    tm.set_tip_codes(big, lut, {n: i for i, n in enumerate(names)})
    tm.set_rate_model(rate)
    tm.set_substitution_model(model)
    tm.initialise()
    a, b = tm.traversal.root_edge
    total, pattern = tm._pattern_lnl(a, b)
    tips = {tm.traversal.names[n]: np.ascontiguousarray(lut[codes[i]]) for i, n in enumerate(names)}
    want = oracle.tree_lnl(tm.traversal, tips, model.p, model.freqs, rate.rates, rate.weights)
    assert np.array_equal(pattern.reshape(reps, block), np.tile(pattern[:block], (reps, 1)))   # bitwise periodic
    assert_lnl_close(pattern[:block], want)
    assert_lnl_close(total, reps * want.sum())


@pytest.mark.parametrize("tree_fn,n_taxa,n_pat", [(random_tree, 300, 40000), (caterpillar_tree, 200, 10000),
                                                  (balanced_tree, 256, 9000), (random_tree, 64, 33), (random_tree, 7, 1)])
def test_resident_kernels_vs_oracle(tree_fn, n_taxa, n_pat):
    tree, names, codes, lut = synthetic(n_taxa, n_pat, 4, seed=n_taxa * 7 + n_pat, tree_fn=tree_fn)
    model = phy.substitution_models.GTR([6., 5., 4., 3., 2., 1.], [0.1, 0.2, 0.3, 0.4])
    rate = phy.rate_models.GammaRateModel(4, 0.5)
    w = np.random.default_rng(2).integers(1, 4, size=n_pat)
    want = None
    for store in (False, True):
        tm = phy.TreeModel(mode="resident", store_partials=store)
        tm.set_tree(tree)
        tm.set_tip_codes(codes, lut, {n: i for i, n in enumerate(names)}, siteweights=w)
        tm.set_rate_model(rate)
        tm.set_substitution_model(model)
        tm.initialise()
        total, pattern = tm._pattern_lnl(*tm.traversal.root_edge)
        if want is None:
            tips = {tm.traversal.names[n]: np.ascontiguousarray(lut[codes[i]]) for i, n in enumerate(names)}
            want = oracle.tree_lnl(tm.traversal, tips, model.p, model.freqs, rate.rates, rate.weights)
        assert_lnl_close(pattern, want)
        assert_lnl_close(total, float(np.dot(want, w)))


@pytest.mark.parametrize("K,rate", [(1, lambda: phy.rate_models.UniformRateModel()),
                                    (2, lambda: phy.rate_models.InvariantSitesModel(0.2)),
                                    (8, lambda: phy.rate_models.GammaRateModel(8, 0.4))])
def test_resident_other_category_counts(K, rate):
    tree, names, codes, lut = synthetic(90, 6000, 4, seed=40 + K)
    model = phy.substitution_models.HKY85(2.0, [0.3, 0.2, 0.2, 0.3])
    tm = phy.TreeModel(store_partials=False)
    tm.set_tree(tree)
    tm.set_tip_codes(codes, lut, {n: i for i, n in enumerate(names)})
    r = rate()
    tm.set_rate_model(r)
    tm.set_substitution_model(model)
    tm.initialise()
    total, pattern = tm._pattern_lnl(*tm.traversal.root_edge)
    tips = {tm.traversal.names[n]: np.ascontiguousarray(lut[codes[i]]) for i, n in enumerate(names)}
    want = oracle.tree_lnl(tm.traversal, tips, model.p, model.freqs, r.rates, r.weights)
    assert_lnl_close(pattern, want)


def test_pipelined_evaluation_from_host_codes_matches_resident():
    import torch
    tree, names, codes, lut = synthetic(120, 70000, 4, seed=77)
    model = phy.substitution_models.GTR([6., 5., 4., 3., 2., 1.], [0.1, 0.2, 0.3, 0.4])
    rate = phy.rate_models.GammaRateModel(4, 0.5)
    tm = phy.TreeModel(store_partials=False)
    tm.set_tree(tree)
    tm.set_tip_codes(codes, lut, {n: i for i, n in enumerate(names)})
    tm.set_rate_model(rate)
    tm.set_substitution_model(model)
    tm.initialise()
    a, b = tm.traversal.root_edge
    length = tm.traversal.brlens[(a, b)]
    total, pattern = tm.engine.lnl_resident(a, b, length, want_pattern=True)
    # a different alignment arrives on the host (pinned): evaluate it with copy/compute overlap
    rng = np.random.default_rng(5)
    other = rng.integers(0, 5, size=codes.shape).astype(np.uint8)
    pinned = torch.from_numpy(other).pin_memory().numpy()
    for chunks in (1, 3, 8, 32):
        t2, p2 = tm.engine.lnl_from_host(pinned, a, b, length, n_chunks=chunks, want_pattern=True)
        tips = {tm.traversal.names[n]: np.ascontiguousarray(lut[other[i]]) for i, n in enumerate(names)}
        want = oracle.tree_lnl(tm.traversal, tips, model.p, model.freqs, rate.rates, rate.weights)
        assert_lnl_close(p2, want)
        assert_lnl_close(t2, want.sum())
    # and the original alignment again, through the same path
    t3, p3 = tm.engine.lnl_from_host(codes, a, b, length, want_pattern=True)
    assert np.array_equal(p3, pattern) and t3 == total
    # two codes per byte: same bits out, half the bytes in
    packed = torch.from_numpy(phy.LikelihoodEngine.pack_codes(other)).pin_memory().numpy()
    assert packed.shape == (other.shape[0], other.shape[1] // 2)
    for chunks in (1, 5, 16):
        t4, p4 = tm.engine.lnl_from_host(packed, a, b, length, n_chunks=chunks, want_pattern=True, packed=True)
        assert np.array_equal(p4, p2)
        assert_lnl_close(t4, t2)
    # the device now holds packed codes: the resident evaluation keeps working, tip readers ask for set_tips
    t5, p5 = tm.engine.lnl_resident(a, b, length, want_pattern=True)
    assert np.array_equal(p5, p2)
    with pytest.raises(RuntimeError):
        tm.engine.get_partials(tm.traversal.names[names[0]])
    # three bits per code in two planes (look-up tables of at most 8 rows): same bits out, 3/8 of the bytes in
    low, high = (torch.from_numpy(x).pin_memory().numpy() for x in phy.LikelihoodEngine.split_codes(other))
    assert low.shape == (other.shape[0], other.shape[1] // 4) and high.shape == (other.shape[0], other.shape[1] // 8)
    for chunks in (1, 5, 16):
        t6, p6 = tm.engine.lnl_from_host_split((low, high), a, b, length, n_chunks=chunks, want_pattern=True)
        assert np.array_equal(p6, p2)
        assert_lnl_close(t6, t2)
    t7, p7 = tm.engine.lnl_resident(a, b, length, want_pattern=True)       # the device keeps the split codes
    assert np.array_equal(p7, p2)
    assert tm.lnl_from_host_codes((low, high)) == t6                       # the TreeModel-level call
    tm.engine.set_tips(codes, lut, tm._tip_rows()[1])                      # back to one byte per code
    t8, p8 = tm.engine.lnl_resident(a, b, length, want_pattern=True)
    assert np.array_equal(p8, pattern)


def test_two_host_fed_evaluations_in_flight():
    """lnl_from_host_submit: alignments arrive one after the other on the host; the copy of one runs under the walk of the
    one before.  Every result must be the bits the immediate call returns for that alignment, in order, whatever the mix of
    code formats, and the two code slots / result words must not bleed into each other."""
    import torch
    tree, names, codes, lut = synthetic(90, 50021, 4, seed=31)
    model = phy.substitution_models.GTR([6., 5., 4., 3., 2., 1.], [0.1, 0.2, 0.3, 0.4])
    rate = phy.rate_models.GammaRateModel(4, 0.5)
    tm = phy.TreeModel(store_partials=False)
    tm.set_tree(tree)
    tm.set_tip_codes(codes, lut, {n: i for i, n in enumerate(names)})
    tm.set_rate_model(rate)
    tm.set_substitution_model(model)
    tm.initialise()
    rng = np.random.default_rng(9)
    aligns = [rng.integers(0, 5, size=codes.shape).astype(np.uint8) for _ in range(5)]
    pin = lambda x: torch.from_numpy(x).pin_memory().numpy()        # noqa: E731
    nibbles = [pin(phy.LikelihoodEngine.pack_codes(a)) for a in aligns]
    planes = [tuple(pin(x) for x in phy.LikelihoodEngine.split_codes(a)) for a in aligns]
    want = [tm.lnl_from_host_codes(n) for n in nibbles]
    assert len(set(want)) == 5
    # strictly alternating submit / result, depth two
    got, pending = [], None
    for i in range(5):
        nxt = tm.lnl_from_host_submit(planes[i] if i % 2 else nibbles[i], n_chunks=(3, 0, 16)[i % 3])
        if pending is not None:
            got.append(pending.result())
        pending = nxt
    got.append(pending.result())
    assert got == want
    # results asked for late and out of order; a third submission forces the first slot's value out in time
    h0 = tm.lnl_from_host_submit(planes[0])
    h1 = tm.lnl_from_host_submit(planes[1])
    h2 = tm.lnl_from_host_submit(nibbles[2])
    assert (h2.result(), h0.result(), h1.result()) == (want[2], want[0], want[1])
    # branch lengths may change between submissions
    key = sorted(tm.traversal.brlens.keys())[4]
    old = tm.traversal.brlens[key]
    ha = tm.lnl_from_host_submit(planes[3])
    tm.traversal.brlens[key] = old * 2.5
    tm.compute_partials()
    hb = tm.lnl_from_host_submit(planes[3])
    assert ha.result() == want[3] and hb.result() == tm.lnl_from_host_codes(planes[3]) != want[3]
    tm.traversal.brlens[key] = old
    tm.compute_partials()
    # afterwards the device holds the last alignment: the resident evaluation works on it
    a, b = tm.traversal.root_edge
    assert tm.lnl_from_host_submit(nibbles[4]).result() == want[4]
    assert tm.engine.lnl_resident(a, b, tm.traversal.brlens[(a, b)])[0] == want[4]
    with pytest.raises(ValueError):
        tm.engine.host_fed_submit(aligns[0], a, b, 0.1)             # one byte per code: no room for two slots


@pytest.mark.parametrize("n_pat", [1, 2, 63, 64, 65, 127, 4097])
def test_packed_codes_ragged_pattern_counts(n_pat):
    tree, names, codes, lut = synthetic(33, n_pat, 4, seed=900 + n_pat)
    model = phy.substitution_models.GTR([6., 5., 4., 3., 2., 1.], [0.1, 0.2, 0.3, 0.4])
    rate = phy.rate_models.GammaRateModel(4, 0.5)
    w = np.random.default_rng(3).integers(1, 4, size=n_pat)
    tm = phy.TreeModel(store_partials=False)
    tm.set_tree(tree)
    tm.set_tip_codes(codes, lut, {n: i for i, n in enumerate(names)}, siteweights=w)
    tm.set_rate_model(rate)
    tm.set_substitution_model(model)
    tm.initialise()
    a, b = tm.traversal.root_edge
    length = tm.traversal.brlens[(a, b)]
    tips = {tm.traversal.names[n]: np.ascontiguousarray(lut[codes[i]]) for i, n in enumerate(names)}
    want = oracle.tree_lnl(tm.traversal, tips, model.p, model.freqs, rate.rates, rate.weights)
    for packed in (False, True):
        src = phy.LikelihoodEngine.pack_codes(codes) if packed else codes
        t, p = tm.engine.lnl_from_host(src, a, b, length, want_pattern=True, packed=packed)
        assert_lnl_close(p, want)
        assert_lnl_close(t, float(np.dot(want, w)))
    t, p = tm.engine.lnl_from_host_split(phy.LikelihoodEngine.split_codes(codes), a, b, length, want_pattern=True)
    assert_lnl_close(p, want)
    assert_lnl_close(t, float(np.dot(want, w)))


def test_four_patterns_per_lane_knob_gives_the_same_answer(monkeypatch):
    """PHB_PAIR_PPT=4 selects the 128-pattern-tile flavour of the lnL-only kernel (tuning knob, K = 4 only)."""
    tree, names, codes, lut = synthetic(150, 30011, 4, seed=4242)
    model = phy.substitution_models.GTR([6., 5., 4., 3., 2., 1.], [0.1, 0.2, 0.3, 0.4])
    rate = phy.rate_models.GammaRateModel(4, 0.5)
    w = np.random.default_rng(8).integers(1, 4, size=codes.shape[1])
    tm = phy.TreeModel(store_partials=False)
    tm.set_tree(tree)
    tm.set_tip_codes(codes, lut, {n: i for i, n in enumerate(names)}, siteweights=w)
    tm.set_rate_model(rate)
    tm.set_substitution_model(model)
    tm.initialise()
    a, b = tm.traversal.root_edge
    length = tm.traversal.brlens[(a, b)]
    monkeypatch.setenv("PHB_PAIR_PPT", "2")
    _lib.lib().phb_reload_tuning()
    t2, p2 = tm.engine.lnl_resident(a, b, length, want_pattern=True)
    monkeypatch.setenv("PHB_PAIR_PPT", "4")
    _lib.lib().phb_reload_tuning()
    t4, p4 = tm.engine.lnl_resident(a, b, length, want_pattern=True)
    packed = phy.LikelihoodEngine.pack_codes(codes)
    t4p, p4p = tm.engine.lnl_from_host(packed, a, b, length, n_chunks=7, want_pattern=True, packed=True)
    t4s, p4s = tm.engine.lnl_from_host_split(phy.LikelihoodEngine.split_codes(codes), a, b, length, n_chunks=7, want_pattern=True)
    assert np.array_equal(p4s, p4p)
    monkeypatch.delenv("PHB_PAIR_PPT")
    _lib.lib().phb_reload_tuning()
    assert_lnl_close(p4, p2)
    assert np.array_equal(p4p, p4)
    assert_lnl_close(t4, t2)
    tips = {tm.traversal.names[n]: np.ascontiguousarray(lut[codes[i]]) for i, n in enumerate(names)}
    want = oracle.tree_lnl(tm.traversal, tips, model.p, model.freqs, rate.rates, rate.weights)
    assert_lnl_close(p4, want)


def test_headline_shape_full_size_properties():
    """
    BASELINE configs[1] at its full size - 1000 taxa x 1,000,000 patterns, GTR+G4 - through the lnL-only path the
    benchmark times.  The alignment is a 500-pattern block repeated 2000 times, so (size-independent properties):
    per-pattern lnL is bitwise periodic in the block, the total is 2000 x the block total, the block itself matches the
    oracle to 1e-10, and the evaluation fed from packed host codes returns the same bits as the resident one.
    """
    n_taxa, block, reps = 1000, 500, 2000
    tree, names, codes, lut = synthetic(n_taxa, block, 4, seed=2)
    big = np.ascontiguousarray(np.tile(codes, (1, reps)))
    assert big.shape == (1000, 1000000)
    model = phy.substitution_models.GTR([6., 5., 4., 3., 2., 1.], [0.1, 0.2, 0.3, 0.4])
    rate = phy.rate_models.GammaRateModel(4, 0.5)
    tm = phy.TreeModel(store_partials=False)
    tm.set_tree(tree)
    tm.set_tip_codes(big, lut, {n: i for i, n in enumerate(names)})
    tm.set_rate_model(rate)
    tm.set_substitution_model(model)
    tm.initialise()
    a, b = tm.traversal.root_edge
    length = tm.traversal.brlens[(a, b)]
    total, pattern = tm.engine.lnl_resident(a, b, length, want_pattern=True)
    tips = {tm.traversal.names[n]: np.ascontiguousarray(lut[codes[i]]) for i, n in enumerate(names)}
    want = oracle.tree_lnl(tm.traversal, tips, model.p, model.freqs, rate.rates, rate.weights)
    assert np.array_equal(pattern.reshape(reps, block), np.tile(pattern[:block], (reps, 1)))
    assert_lnl_close(pattern[:block], want)
    assert_lnl_close(total, reps * want.sum())
    packed = phy.LikelihoodEngine.pack_codes(big)
    t2, p2 = tm.engine.lnl_from_host(packed, a, b, length, n_chunks=32, want_pattern=True, packed=True)
    assert np.array_equal(p2, pattern) and t2 == total


@pytest.mark.parametrize("shape", ["cfg5_shard", "cfg3", "cfg4"])
def test_derivative_configs_full_size_properties(shape):
    """
    BASELINE configs 5 (one GPU's shard: 2000 taxa x 62,500 patterns, GTR+G4), 3 (500 taxa x 100,000 patterns, LG+G4) and
    4 (100 taxa x 50,000 codons, GY94+G4, 61 states on the FP64 tensor cores) at full size through the derivative path - post-order pass, pre-order pass, all edges in one launch.  The alignment
    is a small block repeated, so (size-independent properties): the total is reps x the block total, which the oracle
    gives to 1e-10; EVERY edge reproduces that total (pulley principle); derivatives at other trial lengths agree between
    the first pass and the passes that read the per-edge sum tables; a Newton sweep does not lower lnL.
    """
    from phylo_utils_b200.optimise import edge_nodes, optimise_branch_lengths
    if shape == "cfg5_shard":
        n_taxa, block, reps, n_states = 2000, 250, 250, 4
        model = phy.substitution_models.GTR([6., 5., 4., 3., 2., 1.], [0.1, 0.2, 0.3, 0.4])
    elif shape == "cfg3":
        n_taxa, block, reps, n_states = 500, 200, 500, 20
        model = phy.substitution_models.LG()
    else:
        from phylo_utils_b200.substitution_models.codon import f3x4
        n_taxa, block, reps, n_states = 100, 125, 400, 61
        model = phy.substitution_models.GY94(2.0, 0.2, f3x4(np.random.default_rng(4).dirichlet(np.ones(4) * 5, size=3)))
    tree, names, codes, lut = synthetic(n_taxa, block, n_states, seed=5)
    big = np.ascontiguousarray(np.tile(codes, (1, reps)))
    rate = phy.rate_models.GammaRateModel(4, 0.5)
    tm = phy.TreeModel(up_partials=True)
    tm.set_tree(tree)
    tm.set_tip_codes(big, lut, {n: i for i, n in enumerate(names)})
    tm.set_rate_model(rate)
    tm.set_substitution_model(model)
    tm.initialise()
    tips = {tm.traversal.names[n]: np.ascontiguousarray(lut[codes[i]]) for i, n in enumerate(names)}
    want = reps * oracle.tree_lnl(tm.traversal, tips, model.p, model.freqs, rate.rates, rate.weights).sum()
    total = tm.lnl()
    assert_lnl_close(total, want)
    if shape == "cfg4":
        # per-pattern values: bitwise periodic in the block, the block itself against the oracle
        pattern = tm.compute_likelihood_at_edge(*tm.traversal.root_edge)
        assert pattern.shape == (50000,)
        assert np.array_equal(pattern.reshape(reps, block), np.tile(pattern[:block], (reps, 1)))
        assert_lnl_close(pattern[:block], oracle.tree_lnl(tm.traversal, tips, model.p, model.freqs, rate.rates, rate.weights))
    tm.compute_up_partials()
    nodes = edge_nodes(tm.traversal)
    lengths = tm.lengths_above(nodes)
    first = tm.edge_derivatives(nodes, lengths)
    assert np.all(np.isfinite(first))
    assert np.all(np.abs(first[:, 0] - want) <= 1e-10 * abs(want))
    again = tm.edge_derivatives(nodes, lengths)                      # from the sum tables
    # two summation orders of the same terms.  At 61 states the derivative sums cancel heavily (244 terms of both signs
    # per pattern): both passes sit 1e-10 .. 4e-9 from the composed oracle (tools/diag_codon_derivs.py), hence the
    # derivative tolerance of DESIGN.md section 5 (1e-8) here instead of 1e-12
    rtol = 1e-8 if shape == "cfg4" else 1e-12
    assert np.allclose(first, again, rtol=rtol, atol=1e-6)
    other = tm.edge_derivatives(nodes, lengths * 1.5)
    tm.compute_up_partials()
    assert np.allclose(tm.edge_derivatives(nodes, lengths * 1.5), other, rtol=rtol, atol=1e-6)   # first pass again
    res = optimise_branch_lengths(tm, max_sweeps=1, inner_iterations=2, tol=0.0)
    assert res["lnl"] >= total


@pytest.mark.parametrize("tree_fn,n_taxa,n_pat,mode", [
    (random_tree, 300, 6000, "auto"),         # deep enough that every pattern is rescaled many times on the way up
    (caterpillar_tree, 200, 3000, "tile"),
    (random_tree, 300, 20000, "resident"),    # the stored operand-resident walk (one binary exponent per pattern)
])
def test_per_category_lnl_vs_oracle_where_rescaling_is_routine(tree_fn, n_taxa, n_pat, mode):
    """The engine keeps ONE binary exponent per pattern, the reference one natural-log scaler per (site,
    category) (numba_likelihood_engine.py:37-44).  Per-category root values (`lnl_node`, :82-87) must still agree wherever
    the reference's value is representable next to the pattern's best category: checked here at sizes where the 2^-128
    threshold fires at most levels of the tree, not only on the 10-taxon golden cases."""
    tree, names, codes, lut = synthetic(n_taxa, n_pat, 4, seed=7 * n_taxa + n_pat, tree_fn=tree_fn)
    model = phy.substitution_models.GTR([6., 5., 4., 3., 2., 1.], [0.1, 0.2, 0.3, 0.4])
    rate = phy.rate_models.GammaRateModel(4, 0.5)
    tm = phy.TreeModel(mode=mode)
    tm.set_tree(tree)
    tm.set_tip_codes(codes, lut, {n: i for i, n in enumerate(names)})
    tm.set_rate_model(rate)
    tm.set_substitution_model(model)
    tm.initialise()
    a, b = tm.traversal.root_edge
    length = tm.traversal.brlens[(a, b)]
    _, pattern, cat = tm.engine.root_lnl(a, b, length, want_pattern=True, want_cat=True, root_pmats=tm._root_pmats(length))
    assert tm.scale.min() < -128                      # the rescaling branch ran, repeatedly
    tips = {tm.traversal.names[n]: np.ascontiguousarray(lut[codes[i]]) for i, n in enumerate(names)}
    _, ot = oracle.tree_lnl(tm.traversal, tips, model.p, model.freqs, rate.rates, rate.weights, return_tree=True)
    root_pm = np.stack([model.p(0, rate.rates), model.p(length, rate.rates)])
    want_pattern, want_cat = ot.likelihood_at_edge(a, b, root_pm, model.freqs, rate.weights, want_cat=True)
    assert_lnl_close(pattern, want_pattern)
    # One exponent per pattern: the rescaling fires when the pattern's largest entry drops below 2^-128, and a product of
    # two such operands can reach 2^-256 before it is rescaled - so a category more than ~2^-766 (530 nats) below the
    # pattern's best one can run into the denormal range or flush to zero (documented quirk; its weight in the
    # mixture is < 1e-230).  Everything within 500 nats must match to 1e-10; beyond that: -inf, or the value to within a nat.
    spread = want_cat.max(axis=1, keepdims=True) - want_cat
    near = spread < 500
    assert near.mean() > 0.7 and spread.max() > 200           # three categories in full, the slowest one hundreds of nats down
    assert_lnl_close(cat[near], want_cat[near], what="per-category lnL")
    far = ~near
    assert np.all(np.isneginf(cat[far]) | (np.abs(cat[far] - want_cat[far]) <= 1.0))
    assert_lnl_close(oracle.mix_categories(cat, rate.weights), pattern, rtol=1e-13)


@pytest.mark.parametrize("n_states,K", [(20, 2), (20, 5), (61, 2)])
def test_dmma_kernels_with_a_run_time_category_count(n_states, K):
    """The FP64 tensor-core kernels are specialised for K = 4 (category count as a template parameter, clv_mma.cu); every other
    count goes through the instantiation that reads K at run time - post-order walk in both launch modes against the oracle,
    and the pre-order pass + derivative passes through the pulley principle (lnL from any edge = lnL at the root)."""
    if n_states == 20:
        model = phy.substitution_models.WAG()
        tree, names, codes, lut = synthetic(40, 1500, 20, seed=11 + K)
    else:
        from phylo_utils_b200.substitution_models.codon import f3x4
        model = phy.substitution_models.GY94(2.0, 0.3, f3x4(np.random.default_rng(5).dirichlet(np.ones(4) * 5, size=3)))
        tree, names, codes, lut = synthetic(16, 300, 61, seed=13)
    rate = phy.rate_models.GammaRateModel(K, 0.6)
    for mode in ("tile", "level"):
        _, total, pattern, want = run_both(tree, names, codes, lut, model, rate, mode)
        assert_lnl_close(pattern, want)
    tm = phy.TreeModel(up_partials=True)
    tm.set_tree(tree)
    tm.set_tip_codes(codes, lut, {n: i for i, n in enumerate(names)})
    tm.set_rate_model(rate)
    tm.set_substitution_model(model)
    tm.initialise()
    root = tm.lnl()
    tm.compute_up_partials()
    nodes = [n for n in range(2 * len(names) - 2) if n != tm.traversal.root_edge[1]][::3][:12]
    for _ in range(2):                      # first pass (writes the sum tables) and a later one (reads them)
        d = tm.edge_derivatives(nodes)
        assert_lnl_close(d[:, 0], np.full(len(nodes), root), rtol=1e-9, what="lnL across an edge")
