/*
 * ORACLE - TEST INFRASTRUCTURE ONLY.  Not part of the product: only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this.  The product path
 * (phylo_utils_b200/) never links, imports or calls it.
 *
 * Plain-C restatement of the reference's CPU likelihood engine, function by function, in the
 * reference's own data layout (partials [node][site][cat][state] fp64 with tips replicated over
 * categories, natural-log scalers per (site, cat)):
 *
 *   oracle_clv               <- clv                numba_likelihood_engine.py:10-46
 *   oracle_lnl_node          <- lnl_node           numba_likelihood_engine.py:82-87
 *   oracle_lnl_branch        <- lnl_branch /       numba_likelihood_engine.py:60-79
 *                               lnl_branch_derivs  numba_likelihood_engine.py:49-57
 *   oracle_compute_partials  <- TreeModel.compute_partials               tree_model.py:160-176
 *   oracle_likelihood_at_edge<- compute_partials_at_edge + compute_likelihood_at_edge  tree_model.py:178-217
 *
 * Parity is PINNED: tests/test_oracle.py checks these against outputs of the unmodified reference
 * (numba engine driven through TreeModel under the shims in oracle/ref_shims.py) stored in
 * tests/golden/, and against the known answers in the reference's own tests
 * (tests/test_likelihood.py:30-49).
 *
 * Threading mirrors the reference: `clv` is parallel over sites (numba target='parallel',
 * numba_likelihood_engine.py:13) -> OpenMP parallel-for over sites; everything else is serial in the
 * reference but is given the same site-parallel loop here so that the CPU baseline is not penalised.
 */
#include <math.h>
#include <stddef.h>
#include <stdlib.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define SCALE_THRESHOLD (1.0 / 340282366920938463463374607431768211456.0) /* 1 / 2^128, engine.py:7 */

int oracle_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* one (site, all categories) update; numba_likelihood_engine.py:34-44 */
static void clv_site(int K, int A, const double* p1, const double* p2, const double* c1, const double* c2,
                     const double* sa, const double* sb, double* s_out, double* out) {
    for (int cat = 0; cat < K; ++cat) {
        const double* P1 = p1 + (size_t)cat * A * A;
        const double* P2 = p2 + (size_t)cat * A * A;
        const double* a = c1 + (size_t)cat * A;
        const double* b = c2 + (size_t)cat * A;
        double* o = out + (size_t)cat * A;
        double m = -INFINITY;
        for (int i = 0; i < A; ++i) {
            double x = 0.0, y = 0.0;
            for (int j = 0; j < A; ++j) {
                x += P1[i * A + j] * a[j]; /* np.dot(p1[cat], clv1[cat]) */
                y += P2[i * A + j] * b[j]; /* np.dot(p2[cat], clv2[cat]) */
            }
            o[i] = x * y;
            if (o[i] > m) m = o[i];
        }
        if (m < SCALE_THRESHOLD && m > 0) {
            s_out[cat] = sa[cat] + sb[cat] + log(m);
            for (int i = 0; i < A; ++i) o[i] /= m;
        } else {
            s_out[cat] = sa[cat] + sb[cat];
        }
    }
}

void oracle_clv(long S, int K, int A, const double* p1, const double* p2, const double* clv1, const double* clv2,
                const double* scaler_a, const double* scaler_b, double* cml_scaler, double* out, int n_threads) {
    (void)n_threads;
#pragma omp parallel for schedule(static) num_threads(n_threads > 0 ? n_threads : oracle_max_threads())
    for (long s = 0; s < S; ++s)
        clv_site(K, A, p1, p2, clv1 + (size_t)s * K * A, clv2 + (size_t)s * K * A, scaler_a + (size_t)s * K,
                 scaler_b + (size_t)s * K, cml_scaler + (size_t)s * K, out + (size_t)s * K * A);
}

void oracle_lnl_node(long S, int K, int A, const double* pi, const double* partials, const double* scale, double* out,
                     int n_threads) {
    (void)n_threads;
#pragma omp parallel for schedule(static) num_threads(n_threads > 0 ? n_threads : oracle_max_threads())
    for (long s = 0; s < S; ++s)
        for (int cat = 0; cat < K; ++cat) {
            const double* v = partials + ((size_t)s * K + cat) * A;
            double f = 0.0;
            for (int i = 0; i < A; ++i) f += v[i] * pi[i];
            out[(size_t)s * K + cat] = f > 0 ? log(f) + scale[(size_t)s * K + cat] : -INFINITY;
        }
}

/* nd = 0: lnl_branch (probs[A][A], out[S]);  nd = 2: lnl_branch_derivs (probs[3][A][A], out[S][3]) */
void oracle_lnl_branch(long S, int A, int nd, const double* probs, const double* pi, const double* pa, const double* pb,
                       const double* sa, const double* sb, double* out) {
    for (long s = 0; s < S; ++s) {
        const double* a = pa + (size_t)s * A;
        const double* b = pb + (size_t)s * A;
        double f[3] = {0, 0, 0};
        for (int d = 0; d <= nd; ++d) {
            const double* P = probs + (size_t)d * A * A;
            double acc = 0.0;
            for (int i = 0; i < A; ++i) {
                double x = 0.0;
                for (int j = 0; j < A; ++j) x += P[i * A + j] * a[j];
                acc += x * b[i] * pi[i]; /* np.sum(np.dot(probs, a) * b * pi) */
            }
            f[d] = acc;
        }
        double* o = out + (size_t)s * (nd + 1);
        o[0] = log(f[0]) + sa[s] + sb[s];
        if (nd == 2) {
            o[1] = f[1] / f[0];
            o[2] = ((f[2] * f[0]) - (f[1] * f[1])) / (f[0] * f[0]);
        }
    }
}

/*
 * tree_model.py:160-176.  partials [n_nodes][S][K][A], scale [n_nodes][S][K] (tips pre-filled as in
 * initialise, :142-148); rows [n_rows][3] = PAR, CH1, CH2; pmats [n_rows][2][K][A][A] = the two
 * model.p(brlen, rates) results of each row (:168-169).
 */
void oracle_compute_partials(int n_rows, const long* rows, const double* pmats, double* partials, double* scale,
                             long S, int K, int A, int n_threads) {
    const size_t node = (size_t)S * K * A, snode = (size_t)S * K, blk = (size_t)K * A * A;
    for (int r = 0; r < n_rows; ++r) {
        const long par = rows[3 * r], c1 = rows[3 * r + 1], c2 = rows[3 * r + 2];
        oracle_clv(S, K, A, pmats + (size_t)(2 * r) * blk, pmats + (size_t)(2 * r + 1) * blk, partials + c1 * node,
                   partials + c2 * node, scale + c1 * snode, scale + c2 * snode, scale + par * snode,
                   partials + par * node, n_threads);
    }
}

/*
 * tree_model.py:178-217 without the ascertainment branch: root partials on edge (a, b) from
 * root_pmats [2][K][A][A] = { p(0, rates), p(length, rates) }, lnl_node, then
 * logsumexp over categories of (lnl + log w).  pattern_lnl [S]; cat_lnl [S][K] (may be NULL);
 * root_partials [S][K][A] and root_scale [S][K] are caller-provided work arrays, as in TreeModel.
 */
void oracle_likelihood_at_edge(long a, long b, const double* root_pmats, const double* partials, const double* scale,
                               const double* freqs, const double* cat_weights, long S, int K, int A,
                               double* root_partials, double* root_scale, double* cat_lnl, double* pattern_lnl,
                               int n_threads) {
    const size_t node = (size_t)S * K * A, snode = (size_t)S * K, blk = (size_t)K * A * A;
    double* tmp = cat_lnl ? cat_lnl : (double*)malloc(snode * sizeof(double));
    oracle_clv(S, K, A, root_pmats, root_pmats + blk, partials + a * node, partials + b * node, scale + a * snode,
               scale + b * snode, root_scale, root_partials, n_threads);
    oracle_lnl_node(S, K, A, freqs, root_partials, root_scale, tmp, n_threads);
    for (long s = 0; s < S; ++s) { /* scipy.special.logsumexp(x + log w, axis=1) */
        double mx = -INFINITY;
        for (int k = 0; k < K; ++k) {
            const double v = tmp[(size_t)s * K + k] + log(cat_weights[k]);
            if (v > mx) mx = v;
        }
        if (!isfinite(mx)) {
            pattern_lnl[s] = mx;
            continue;
        }
        double acc = 0.0;
        for (int k = 0; k < K; ++k) acc += exp(tmp[(size_t)s * K + k] + log(cat_weights[k]) - mx);
        pattern_lnl[s] = log(acc) + mx;
    }
    if (!cat_lnl) free(tmp);
}
