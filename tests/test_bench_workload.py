"""bench.py's two arms must evaluate the same problem: the pure-numpy workload replica used by the CPU / reference arm
(no phylo_utils_b200 import) against the package's own tree, traversal and model objects - on the CPU oracle."""
import os
import subprocess
import sys

import numpy as np

import phylo_utils_b200 as phy
from oracle import oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def test_workload_replica_matches_the_package_objects():
    n_taxa, n_pat, seed = 57, 300, 2
    p_matrices, prune, threads = bench.cpu_eval_factory(n_taxa, n_pat, seed)
    replica = prune(p_matrices())
    tree = phy.tree.random_tree(n_taxa, seed)
    names = [lf.taxon.label for lf in tree.leaf_node_iter()]
    trav = phy.traversal.Traversal(phy.utils.deepcopy_tree(tree))
    codes, lut = bench.make_codes(n_taxa, 0, n_pat, seed), bench.dna_lut()
    model = phy.substitution_models.GTR(bench.GTR_RATES, bench.GTR_FREQS)
    rate = phy.rate_models.GammaRateModel(bench.NCAT, bench.ALPHA)
    tips = {trav.names[n]: np.ascontiguousarray(lut[codes[i]]) for i, n in enumerate(names)}
    want = oracle.tree_lnl(trav, tips, model.p, model.freqs, rate.rates, rate.weights).sum()
    assert abs(replica - want) <= 1e-11 * abs(want)
    assert threads >= 1


def test_code_blocks_do_not_depend_on_the_sharding():
    full = bench.make_codes(5, 0, 3 * bench.CODE_BLOCK, 7)
    lo, hi = bench.CODE_BLOCK - 17, 2 * bench.CODE_BLOCK + 5
    assert np.array_equal(bench.make_codes(5, lo, hi, 7), full[:, lo:hi])
    assert set(np.unique(full)) == {0, 1, 2, 3, 4}


def test_reference_arm_runs_without_the_package_and_prints_one_line():
    # OMP_NUM_THREADS=1 is what torchrun exports; the arm must ignore it and must not load libphylo_b200.so
    env = dict(os.environ, OMP_NUM_THREADS="1")
    code = ("import sys, json, runpy; sys.argv = ['bench.py', '--impl', 'reference', '--taxa', '40', '--patterns', '4000', "
            "'--cpu-patterns', '400', '--steps', '2', '--warmup', '1']; runpy.run_path('bench.py', run_name='__main__'); "
            "assert not any(m.startswith('phylo_utils_b200') for m in sys.modules), 'reference arm imported the package'")
    out = subprocess.run([sys.executable, "-c", code], cwd=ROOT, env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert out.returncode == 0, out.stderr
    import json
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["scaling"] == "strong" and line["gpu_launches"] == 0
    assert line["cpu_baseline"]["kind"] == "port" and line["value"] > 0
    assert line["cpu_baseline"]["cores"] == len(os.sched_getaffinity(0))
    assert line["e2e"]["h2d_bytes_per_step"] == 0
