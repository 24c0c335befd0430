"""
Branch-length optimisation driver (BASELINE config 5).

The reference ships scalar Brent/dbrent optimisers (src/optimisation.pyx) and a one-edge-at-a-time
re-rooting order (utils.py:137-188) that nothing consumes.  On a GPU the natural unit is a SWEEP:

    1. post-order pass  (down partials for the current lengths)
    2. pre-order pass   (up partials: everything outside each subtree)
    3. for ALL edges at once, a few Newton-Raphson iterations on the edge's own length with both end
       partials held fixed - each iteration is one batched derivative launch (f, f', f'' per edge)
    4. all edges move together (Jacobi); if the joint move lowers lnL it is halved until it does not

Each accepted sweep increases lnL monotonically; the loop stops when the gain falls under ``tol``.
"""
import numpy as np

__all__ = ["edge_nodes", "optimise_branch_lengths"]

MIN_BRANCH_LENGTH = 1.0 / 2 ** 16      # same floor as the reference (substitution_models/abstract.py:8)
MAX_BRANCH_LENGTH = 20.0


def edge_nodes(traversal):
    """One node id per edge of the unrooted tree: the node below the edge; the root edge is owned by root_edge[0]."""
    a, b = traversal.root_edge
    n_nodes = 2 * len(traversal.names) - 2
    return np.asarray([n for n in range(n_nodes) if n != b], dtype=np.int32)


def _get_lengths(br, keys):
    return br.gather(br.slots(keys))


def _set_lengths(br, keys, lengths):
    br.scatter(br.slots(keys), np.asarray(lengths, dtype=np.double))


def newton_step(t, d1, d2, lo=MIN_BRANCH_LENGTH, hi=MAX_BRANCH_LENGTH):
    """Safeguarded Newton-Raphson update of a vector of branch lengths."""
    t = np.asarray(t, dtype=np.double)
    concave = d2 < 0
    with np.errstate(divide="ignore", invalid="ignore"):
        raw = np.where(concave, t - d1 / d2, np.where(d1 > 0, t * 2.0, t * 0.5))
    raw = np.where(np.isfinite(raw), raw, t)
    # never move by more than a factor of four in one iteration
    return np.clip(np.clip(raw, t * 0.25, t * 4.0 + 1e-3), lo, hi)


def optimise_branch_lengths(tm, max_sweeps=20, inner_iterations=3, tol=1e-4, verbose=False):
    """
    Maximise lnL over all branch lengths of ``tm`` (a TreeModel or ShardedTreeModel built with
    ``up_partials=True`` and initialised).  Returns a dict with the lnL trace.
    """
    local = getattr(tm, "local", tm)
    nodes = edge_nodes(tm.traversal)
    br = local.traversal.brlens
    slots = br.slots(local.edge_keys(nodes))   # resolved once: the sweeps run next to millisecond kernels
    lengths = br.gather(slots)
    lnl = tm.lnl()
    trace = [lnl]
    derivative_launches = 0
    for sweep in range(max_sweeps):
        tm.compute_up_partials()
        trial = lengths.copy()
        for _ in range(inner_iterations):
            d = tm.edge_derivatives(nodes, trial)
            derivative_launches += 1
            trial = newton_step(trial, d[:, 1], d[:, 2])
        step = trial - lengths
        alpha, accepted = 1.0, False
        while alpha > 1e-3:
            br.scatter(slots, lengths + alpha * step)
            tm.compute_partials()
            new_lnl = tm.lnl()
            if new_lnl >= lnl:
                accepted = True
                break
            alpha *= 0.5
        if not accepted:
            br.scatter(slots, lengths)
            tm.compute_partials()
            break
        gain = new_lnl - lnl
        lengths = lengths + alpha * step
        lnl = new_lnl
        trace.append(lnl)
        if verbose:
            print("sweep {:2d}  lnL = {:.6f}  gain = {:.3e}  step scale = {}".format(sweep + 1, lnl, gain, alpha))
        if gain < tol:
            break
    return {"lnl": lnl, "trace": trace, "sweeps": len(trace) - 1, "lengths": lengths, "nodes": nodes,
            "derivative_launches": derivative_launches}
