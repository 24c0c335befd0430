"""
Branch-length optimisation driver (BASELINE config 5).

The reference ships scalar Brent/dbrent optimisers (src/optimisation.pyx) and a one-edge-at-a-time
re-rooting order (utils.py:137-188) that nothing consumes.  On a GPU the natural unit is a SWEEP:

    1. post-order pass  (down partials for the current lengths)
    2. pre-order pass   (up partials: everything outside each subtree)
    3. for ALL edges at once, a few Newton-Raphson iterations on the edge's own length with both end
       partials held fixed - each iteration is one batched derivative launch (f, f', f'' per edge);
       edges whose curve is not concave where Newton stands (f'' >= 0) leave the Newton track and are finished
       by a bracketing search on f' (``optimisation.maximise_bracketed``, the batched counterpart of the
       reference's dbrent, src/optimisation.pyx:179-306) - again one launch per step, over those edges only
    4. all edges move together (Jacobi); if the joint move lowers lnL it is halved until it does not

``optimise_by_rerooting`` is the other driver: the reference's own one-edge-at-a-time order
(Traversal.optimising_traversal, utils.py:137-188) executed on the device - every step re-roots the partials in
place (one pruning row) and maximises one edge exactly (Gauss-Seidel), which needs no up partials at all.

Each accepted sweep increases lnL monotonically; the loop stops when the gain falls under ``tol``.
"""
import numpy as np

from .optimisation import maximise_bracketed

__all__ = ["edge_nodes", "optimise_branch_lengths", "optimise_by_rerooting"]

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


def optimise_branch_lengths(tm, max_sweeps=20, inner_iterations=3, tol=1e-4, verbose=False, bracket_fallback=True):
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
    derivative_launches = fallback_edges = 0
    for sweep in range(max_sweeps):
        tm.compute_up_partials()
        trial = lengths.copy()
        unsafe = np.zeros(len(nodes), dtype=bool)
        for _ in range(inner_iterations):
            d = tm.edge_derivatives(nodes, trial)
            derivative_launches += 1
            unsafe |= ~(d[:, 2] < 0) | ~np.isfinite(d).all(axis=1)
            trial = newton_step(trial, d[:, 1], d[:, 2])
        if bracket_fallback and unsafe.any():
            # not concave somewhere along the Newton track: maximise those edges' own curves inside [lo, hi] instead
            sub = np.flatnonzero(unsafe)

            def curve(t, idx):
                return tm.edge_derivatives(nodes[sub[idx]], t)
            best, _, _, evals = maximise_bracketed(curve, MIN_BRANCH_LENGTH, MAX_BRANCH_LENGTH, lengths[sub], tol=1e-6)
            trial[sub] = best
            derivative_launches += evals
            fallback_edges += int(sub.size)
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
            "derivative_launches": derivative_launches, "fallback_edges": fallback_edges}


def _maximise_edge(curve, t0, lo=MIN_BRANCH_LENGTH, hi=MAX_BRANCH_LENGTH, newton_iterations=6, tol=1e-7):
    """Maximum of one edge's lnL curve.  ``curve(t)`` -> (lnL, d1, d2) rows for a vector of trial lengths (one launch).
    Newton-Raphson while the curve is concave and lnL does not drop; otherwise (or if Newton did not settle) the
    bracketing search.  Returns (length, lnL, evaluations)."""
    t = float(np.clip(t0, lo, hi))
    f0, d1, d2 = curve(np.array([t]))[0]
    best_t, best_f, evals, settled = t, f0, 1, False
    for _ in range(newton_iterations):
        if not (d2 < 0) or not np.isfinite(d1):
            break
        nxt = float(np.clip(t - d1 / d2, max(lo, t * 0.1), min(hi, t * 10.0 + 1e-3)))
        if abs(nxt - t) <= tol * max(t, 1e-3):
            settled = True
            break
        f, d1, d2 = curve(np.array([nxt]))[0]
        evals += 1
        if not np.isfinite(f) or f < best_f - 1e-9 * abs(best_f):
            break                                         # Newton overshot: the bracket takes over
        t = nxt
        if f >= best_f:
            best_t, best_f = t, f
    if not settled:
        x, fx, _, n = maximise_bracketed(lambda tt, idx: curve(tt), lo, hi, np.array([best_t]), tol=tol)
        evals += n
        if fx[0] >= best_f:
            best_t, best_f = float(x[0]), float(fx[0])
    return best_t, best_f, evals


def optimise_by_rerooting(tm, max_sweeps=5, tol=1e-4, verbose=False):
    """
    Branch-length optimisation in the reference's own order: ``Traversal.optimising_traversal`` (utils.py:137-188), the
    re-rooting sweep the reference builds but never consumes.  Row by row:

        [-1, -1, -1, L, R]        maximise the root edge L-R
        [PAR, SIB, GPA, NOD, PAR] re-point PAR's partial at NOD (one pruning row from SIB and GPA, on the device:
                                  ``phb_update_node``), then maximise edge NOD-PAR exactly (``phb_branch_derivatives``)
        [NOD, CH1, CH2, -1, -1]   point NOD's partial back at the root, with its children's new lengths

    One edge moves at a time and every later step sees it (Gauss-Seidel), so lnL never decreases; no pre-order pass and
    no up partials are needed.  ``tm``: an initialised single-GPU TreeModel with stored partials.  It is the sequential
    algorithm (3N-5 dependent steps of a few launches each); ``optimise_branch_lengths`` is the batched one.
    """
    eng, trav = tm.engine, tm.traversal
    br = trav.brlens
    table = np.asarray(trav.optimising_traversal)
    tm.compute_partials()
    lnl = tm.lnl()
    trace, evaluations, updates = [lnl], 0, 0
    for sweep in range(max_sweeps):
        for par, c1, c2, x, y in table.tolist():
            if par >= 0:
                eng.update_node(par, c1, br[(par, c1)], c2, br[(par, c2)])
                updates += 1
            if x >= 0:
                best, _, n = _maximise_edge(lambda t: eng.branch_derivatives(x, y, t), br[(x, y)])
                br[(x, y)] = best
                evaluations += n
        # every internal node points at the root edge again and was rebuilt with the new lengths
        a, b = trav.root_edge
        new_lnl = eng.root_lnl(a, b, br[(a, b)])[0]
        gain = new_lnl - lnl
        lnl = new_lnl
        trace.append(lnl)
        if verbose:
            print("re-rooting sweep {:2d}  lnL = {:.6f}  gain = {:.3e}".format(sweep + 1, lnl, gain))
        if gain < tol:
            break
    return {"lnl": lnl, "trace": trace, "sweeps": len(trace) - 1, "edge_evaluations": evaluations, "node_updates": updates}
