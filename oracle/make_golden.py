"""
ORACLE - TEST INFRASTRUCTURE ONLY.  Writes tests/golden/*.npz by running the UNMODIFIED reference
(/root/reference, numba engine driven through its own TreeModel) under the shims of
oracle/ref_shims.py.  Run in the build container only:

    python -m oracle.make_golden

Every case is seeded; inputs (newick, sequences) are stored next to the reference's outputs so
the tests can rebuild the same problem with phylo_utils_b200 and with the C oracle anywhere.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_shims  # noqa: E402
from phylo_utils_b200.tree import random_tree, caterpillar_tree, parse_newick  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
DNA_CHARS = np.array(list("ACGT"))
DNA_AMBIG = np.array(list("ACGTRYMKWSBDHVN-acgtn"))
AA_CHARS = np.array(list("ARNDCQEGHILKMFPSTWYV"))


def seq_matrix(strings):
    return np.frombuffer("".join(strings).encode("ascii"), dtype=np.uint8).reshape(len(strings), -1).copy()


def random_alignment(rng, names, nsite, chars, gap_frac=0.0, gap="-"):
    seqs = []
    for _ in names:
        s = rng.choice(chars, size=nsite)
        if gap_frac > 0:
            s = np.where(rng.random(nsite) < gap_frac, gap, s)
        seqs.append("".join(s))
    return seqs


def run_reference(ref, tree, names, seqs, alphabet, model, rate_model, ascbias=False, keep_nodes=False):
    aln = [ref_shims.Record(n, s) for n, s in zip(names, seqs)]
    tm = ref.tree_model.TreeModel()
    tm.set_tree(tree)
    tm.set_alignment(aln, alphabet)
    tm.set_rate_model(rate_model)
    tm.set_substitution_model(model)
    if ascbias:
        tm.set_ascertainment_bias_correction()
    tm.initialise()
    a, b = tm.traversal.root_edge
    site_lnl = tm.compute_likelihood_at_edge(a, b)
    n_dummy = tm.alignment.shape[2] if ascbias else 0
    out = dict(
        newick=np.array(tree.as_newick()),
        names=np.array(names),
        seqs=seq_matrix(seqs),
        alphabet=np.array(alphabet),
        rates=np.asarray(rate_model.rates, dtype=np.double),
        cat_weights=np.asarray(rate_model.weights, dtype=np.double),
        freqs=np.asarray(model.freqs, dtype=np.double),
        postorder=np.asarray(tm.traversal.postorder_traversal, dtype=np.int64),
        optimising=np.asarray(tm.traversal.optimising_traversal, dtype=np.int64),
        root_edge=np.asarray(tm.traversal.root_edge, dtype=np.int64),
        brlen_keys=np.asarray(sorted(tm.traversal.brlens.keys()), dtype=np.int64),
        brlen_vals=np.asarray([tm.traversal.brlens[k] for k in sorted(tm.traversal.brlens.keys())]),
        tip_names=np.array(sorted(tm.traversal.names, key=tm.traversal.names.get)),
        tip_nodes=np.asarray(sorted(tm.traversal.names.values()), dtype=np.int64),
        patterns=np.asarray(tm.alignment, dtype=np.double),
        siteweights=np.asarray(tm.siteweights, dtype=np.int64),
        inverse_index=np.asarray(tm.inverse_index, dtype=np.int64).reshape(-1),
        site_lnl=site_lnl,
        total_lnl=np.array(site_lnl.sum()),
        root_partials=tm.root_partials.copy(),
        root_scale=tm.root_scale.copy(),
        n_dummy=np.array(n_dummy),
    )
    if model.eigen is not None:
        out.update(evecs=np.ascontiguousarray(model.eigen.evecs), evals=np.ascontiguousarray(model.eigen.evals),
                   ivecs=np.ascontiguousarray(model.eigen.ivecs))
    # per-category values before mixing, recomputed with the reference operators
    cat = ref.engine.lnl_node(model.freqs, tm.root_partials, tm.root_scale)
    out["cat_lnl"] = cat
    if keep_nodes:
        out["partials"] = tm.partials.copy()
        out["scale"] = tm.scale.copy()
    else:
        # log-domain summary per node: log(max_i partial) + scale - what must agree whatever the scaling scheme
        with np.errstate(divide="ignore"):
            out["node_logmax"] = np.log(tm.partials.max(axis=3)) + tm.scale
    return out, tm


def save(name, **arrays):
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **arrays)
    print("wrote {} ({:.1f} kB)".format(path, os.path.getsize(path) / 1e3))


def leaf_names(tree):
    return [lf.taxon.label for lf in tree.leaf_node_iter()]


def main():
    os.makedirs(OUT, exist_ok=True)
    ref = ref_shims.load_reference()
    sm, rmods = ref.substitution_models, ref.rate_models

    # ---- config 1: 10 taxa x 1000 sites, GTR+G4 and JC (as GTR defaults), seed 1 --------------------
    rng = np.random.default_rng(1)
    tree = random_tree(10, 1)
    names = leaf_names(tree)
    seqs = random_alignment(rng, names, 1000, DNA_CHARS)
    gtr = sm.GTR([6., 5., 4., 3., 2., 1.], [0.1, 0.2, 0.3, 0.4])
    out, _ = run_reference(ref, tree, names, seqs, 0, gtr, rmods.GammaRateModel(4, 0.5), keep_nodes=True)
    save("cfg1_gtr_g4", **out)
    out, _ = run_reference(ref, tree, names, seqs, 0, sm.GTR(), rmods.GammaRateModel(4, 0.5))
    save("cfg1_jc_g4", **out)
    out, _ = run_reference(ref, tree, names, seqs, 0, gtr, rmods.UniformRateModel())
    save("cfg1_gtr_uniform", **out)

    # ---- ambiguity codes, duplicated columns, +I+G (5 categories, one of rate 0) -----------------------
    rng = np.random.default_rng(2)
    tree = random_tree(7, 2)
    names = leaf_names(tree)
    base = random_alignment(rng, names, 60, DNA_AMBIG)
    cols = rng.integers(0, 60, size=400)
    seqs = ["".join(np.array(list(s))[cols]) for s in base]
    hky = sm.HKY85(2.5, [0.3, 0.2, 0.15, 0.35])
    out, _ = run_reference(ref, tree, names, seqs, 0, hky, rmods.InvariantGammaModel(0.2, 4, 0.7), keep_nodes=True)
    save("ambig_hky_ig", **out)
    out, _ = run_reference(ref, tree, names, seqs, 0, sm.TN93(2.0, 3.0, 1.0, [0.25, 0.2, 0.3, 0.25]),
                           rmods.InvariantSitesModel(0.3))
    save("ambig_tn93_inv", **out)

    # ---- deep tree: scaling threshold is crossed many times --------------------------------------------
    rng = np.random.default_rng(3)
    tree = random_tree(300, 3)
    names = leaf_names(tree)
    seqs = random_alignment(rng, names, 150, DNA_CHARS, gap_frac=0.01)
    out, _ = run_reference(ref, tree, names, seqs, 0, gtr, rmods.GammaRateModel(4, 0.5))
    save("deep300_gtr_g4", **out)
    tree = caterpillar_tree(120, 4)
    names = leaf_names(tree)
    seqs = random_alignment(rng, names, 100, DNA_CHARS)
    out, _ = run_reference(ref, tree, names, seqs, 0, sm.K80(2.0), rmods.GammaRateModel(4, 1.3))
    save("ladder120_k80_g4", **out)

    # ---- protein (config 3 scaled down) ------------------------------------------------------------------
    rng = np.random.default_rng(4)
    tree = random_tree(12, 5)
    names = leaf_names(tree)
    seqs = random_alignment(rng, names, 300, AA_CHARS, gap_frac=0.02)
    out, _ = run_reference(ref, tree, names, seqs, 1, sm.LG(), rmods.GammaRateModel(4, 0.8), keep_nodes=True)
    save("prot12_lg_g4", **out)
    out, _ = run_reference(ref, tree, names, seqs, 1, sm.WAG(), rmods.GammaRateModel(4, 0.8))
    save("prot12_wag_g4", **out)
    tree = random_tree(150, 6)
    names = leaf_names(tree)
    seqs = random_alignment(rng, names, 64, AA_CHARS)
    out, _ = run_reference(ref, tree, names, seqs, 1, sm.JTT(), rmods.GammaRateModel(4, 0.6))
    save("prot150_jtt_g4", **out)

    # ---- non-reversible model (P through the Taylor expm) -----------------------------------------------
    rng = np.random.default_rng(5)
    tree = random_tree(8, 7)
    names = leaf_names(tree)
    seqs = random_alignment(rng, names, 200, DNA_CHARS)
    unrest = sm.Unrest(rates=[[0., 1., 2., 3.], [4., 0., 5., 6.], [7., 8., 0., 9.], [10., 11., 12., 0.]])
    out, _ = run_reference(ref, tree, names, seqs, 0, unrest, rmods.GammaRateModel(4, 0.5))
    save("nonrev_unrest_g4", **out)

    # ---- Lewis ascertainment-bias correction --------------------------------------------------------------
    rng = np.random.default_rng(6)
    tree = random_tree(6, 8)
    names = leaf_names(tree)
    seqs = random_alignment(rng, names, 120, DNA_CHARS)
    arr = np.array([list(s) for s in seqs])
    variable = [j for j in range(arr.shape[1]) if len(set(arr[:, j])) > 1]
    seqs = ["".join(arr[i, variable]) for i in range(arr.shape[0])]
    out, _ = run_reference(ref, tree, names, seqs, 0, gtr, rmods.UniformRateModel(), ascbias=True)
    save("ascbias_gtr_uniform", **out)
    # with K > 1 the reference pools the dummy-pattern likelihoods of all categories unweighted
    # (tree_model.py:213), which only stays below 1 on long trees
    tree = random_tree(6, 8, min_len=0.6, max_len=1.5)
    out, _ = run_reference(ref, tree, names, seqs, 0, gtr, rmods.GammaRateModel(4, 2.0), ascbias=True)
    assert np.all(np.isfinite(out["site_lnl"])), "asc-bias golden case must be finite"
    save("ascbias_gtr_g4", **out)

    # ---- engine-level vectors at A = 61 (no codon model exists in the reference; the engine is generic) ---
    rng = np.random.default_rng(7)
    K, A, S = 4, 61, 40

    def stochastic():
        m = rng.random((K, A, A)) ** 4
        return m / m.sum(axis=2, keepdims=True)
    p1, p2 = stochastic(), stochastic()
    c1 = rng.random((S, K, A)) * np.exp(-rng.uniform(0, 200, size=(S, K, 1)))
    c2 = rng.random((S, K, A)) * np.exp(-rng.uniform(0, 200, size=(S, K, 1)))
    sa = -rng.uniform(0, 50, size=(S, K))
    sb = -rng.uniform(0, 50, size=(S, K))
    sp = np.zeros((S, K))
    o = ref.engine.clv(p1, p2, c1, c2, sa, sb, sp)
    pi = rng.random(A)
    pi /= pi.sum()
    ln = ref.engine.lnl_node(pi, o, sp)
    save("engine_a61", p1=p1, p2=p2, clv1=c1, clv2=c2, sa=sa, sb=sb, out=o, out_scale=sp, pi=pi, lnl_node=ln)

    # ---- derivative primitives --------------------------------------------------------------------------
    rng = np.random.default_rng(8)
    S = 50
    t, r = 0.17, 1.0
    probs = np.stack([gtr.p(t), gtr.dp_dt(t), gtr.d2p_dt2(t)])
    pa = rng.random((S, 4))
    pb = rng.random((S, 4))
    sa1 = -rng.uniform(0, 30, size=S)
    sb1 = -rng.uniform(0, 30, size=S)
    d = np.stack([ref.engine.lnl_branch_derivs(probs, gtr.freqs, pa[i], pb[i], sa1[i:i + 1], sb1[i:i + 1])
                  for i in range(S)])
    l0 = np.array([ref.engine.lnl_branch(probs[0], gtr.freqs, pa[i], pb[i], sa1[i:i + 1], sb1[i:i + 1])
                   for i in range(S)])
    save("engine_branch", probs=probs, pi=gtr.freqs, pa=pa, pb=pb, sa=sa1, sb=sb1, derivs=d, lnl=l0, t=np.array(t))

    # ---- known answer from the reference's own test-suite: K80(2.) pair (tests/test_likelihood.py:30-49) ---
    k80 = sm.K80(2.)
    cvec = np.array([[[0., 1., 0., 0.]]])          # C
    tvec = np.array([[[0., 0., 0., 1.]]])          # T  (a transition pair)
    sc = np.zeros((1, 1))
    part = ref.engine.clv(k80.p(0.1)[None], k80.p(0.2)[None], cvec, tvec, np.zeros((1, 1)), np.zeros((1, 1)), sc)
    lnl = ref.engine.lnl_node(k80.freqs, part, sc)
    save("k80_pair", partials=part, lnl=lnl, p01=k80.p(0.1), p02=k80.p(0.2), freqs=k80.freqs)

    # ---- model-level vectors: Q, freqs, P / dP / d2P for every model ---------------------------------------
    rates4 = ref.rate_models.GammaRateModel(4, 0.5).rates
    f4 = [0.1, 0.2, 0.3, 0.4]
    models = dict(
        JC69=sm.JC69(), K80=sm.K80(1.5), F81=sm.F81(f4), F84=sm.F84(1.5, f4), HKY85=sm.HKY85(1.5, f4),
        TN93=sm.TN93(2.5, 2.4, freqs=f4), GTR=gtr, Strsym=sm.Strsym([1., 2., 3., 4., 5., 6.]), Unrest=unrest,
        WAG=sm.WAG(), LG=sm.LG(), JTT=sm.JTT(), Dayhoff=sm.Dayhoff())
    arrays = {}
    for name, m in models.items():
        arrays[name + "_q"] = np.asarray(m.q())
        arrays[name + "_freqs"] = np.asarray(m.freqs)
        if name == "JC69":
            arrays[name + "_p"] = np.stack([ref.substitution_models.abstract.Model.p(m, 0.23, rates4)])[0]
            arrays[name + "_p_closed"] = m.p(0.23)
        else:
            arrays[name + "_p"] = m.p(0.23, rates4)
        arrays[name + "_dp"] = m.dp_dt(0.23, rates4)
        arrays[name + "_d2p"] = m.d2p_dt2(0.23, rates4)
    arrays["rates"] = rates4
    arrays["t"] = np.array(0.23)
    save("models", **arrays)

    # ---- discrete gamma: native (PAML C) and scipy flavours ---------------------------------------------------
    alphas = np.array([0.05, 0.1, 0.3, 0.5, 1.0, 2.0, 5.0, 17.3, 50.0])
    ncats = np.array([2, 4, 5, 8])
    native = {}
    for a in alphas:
        for k in ncats:
            native["native_a{}_k{}".format(a, k)] = ref.discrete_gamma.discrete_gamma(float(a), int(k))
            native["median_a{}_k{}".format(a, k)] = ref.discrete_gamma.discrete_gamma(float(a), int(k), True)
            native["scipy_a{}_k{}".format(a, k)] = ref.gamma.discrete_gamma(int(k), float(a))
    save("gamma", alphas=alphas, ncats=ncats, **native)

    # ---- traversal tables for assorted shapes ---------------------------------------------------------------
    trav = {}
    shapes = {"rand5": random_tree(5, 11), "rand23": random_tree(23, 12), "ladder9": caterpillar_tree(9, 13),
              "trifurcating": parse_newick("(a:0.1,b:0.2,(c:0.3,d:0.4):0.5);"),
              "polytomy": parse_newick("((a:1,b:2,c:3,d:4):0.5,(e:1,f:1):0.25,g:2);"),
              "rooted_pair_first": parse_newick("((a:0.1,b:0.2):0.05,c:0.3);")}
    for key, tr in shapes.items():
        clone = ref.tree_model.deepcopy_tree(tr)
        t = ref.traversal.Traversal(clone)
        trav[key + "_newick"] = np.array(tr.as_newick())
        trav[key + "_postorder"] = np.asarray(t.postorder_traversal, dtype=np.int64)
        trav[key + "_optimising"] = np.asarray(t.optimising_traversal, dtype=np.int64)
        trav[key + "_root_edge"] = np.asarray(t.root_edge, dtype=np.int64)
        keys = sorted(t.brlens.keys())
        trav[key + "_brlen_keys"] = np.asarray(keys, dtype=np.int64)
        trav[key + "_brlen_vals"] = np.asarray([t.brlens[k] for k in keys])
        trav[key + "_tip_names"] = np.array(sorted(t.names, key=t.names.get))
        trav[key + "_tip_nodes"] = np.asarray(sorted(t.names.values()), dtype=np.int64)
    trav["keys"] = np.array(sorted(shapes))
    save("traversal", **trav)

    # ---- seq_to_partials for every character of every charmap --------------------------------------------------
    am = ref.alignment_module
    dna_chars = "".join(sorted(am.dna_charmap))
    prot_chars = "".join(sorted(am.protein_charmap))
    bin_chars = "".join(sorted(am.binary_charmap))
    save("charmaps", dna_chars=np.array(dna_chars), dna=am.seq_to_partials(dna_chars, 0),
         protein_chars=np.array(prot_chars), protein=am.seq_to_partials(prot_chars, 1),
         binary_chars=np.array(bin_chars), binary=am.seq_to_partials(bin_chars, 2))


if __name__ == "__main__":
    main()
