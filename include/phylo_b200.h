/*
 * phylo_b200 - C ABI of the B200 tree-likelihood engine.
 *
 * This is the drop-in boundary for the likelihood hot path of kgori/phylo_utils.  The
 * reference has no FFI of its own on this path: the path sits behind a Python module
 * import (phylo_utils/tree_model.py:1) of four numba gufuncs and is driven by TreeModel.
 * Each entry point below names the reference interface (file:line, relative to the
 * reference root) it stands in for.  INTEGRATION.md shows the ctypes stub a maintainer of
 * the reference would add.
 *
 * Conventions
 *   - every function returns an int status (PHB_OK == 0); no C++ exception crosses the ABI
 *   - host pointers are borrowed for the duration of the call only
 *   - all matrices are row-major (C order) double precision unless stated otherwise
 *   - one phb_ctx <-> one GPU <-> one stream; a ctx is not thread-safe
 *   - there is NO CPU fallback: without an sm_100 device phb_create fails with PHB_ERR_NO_DEVICE
 *
 * Layout in HBM (per context; S patterns, K rate categories, A states, N tips)
 *   tips       uint8  [N][S]            state-set code per tip and pattern (+ double lut[n_codes][A])
 *   partials   double [N-2][S][K][A]    one block per internal node, pattern-major, state innermost
 *   scalers    int32  [N-2][S]          cumulative binary exponent per pattern:
 *                                       true partial = stored partial * 2^scaler
 *   P          double [2(N-2)+2][K][A][A]
 * Node ids are the reference's (Traversal numbering, tips and internal nodes share 0..2N-3).
 */
#ifndef PHYLO_B200_H
#define PHYLO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default) /* the library is built with -fvisibility=hidden; only this ABI is exported */
#endif

#define PHB_VERSION 100

enum phb_status {
    PHB_OK = 0,
    PHB_ERR_INVALID = 1,     /* bad argument                          -> ValueError   */
    PHB_ERR_CUDA = 2,        /* CUDA runtime / launch failure         -> RuntimeError */
    PHB_ERR_NO_DEVICE = 3,   /* no usable sm_100 GPU                  -> RuntimeError */
    PHB_ERR_STATE = 4,       /* call out of order (e.g. no schedule)  -> RuntimeError */
    PHB_ERR_NOMEM = 5,       /* workspace too small / allocation      -> MemoryError  */
    PHB_ERR_UNSUPPORTED = 6  /* shape outside what the kernels cover  -> ValueError   */
};

/* phb_create flags */
#define PHB_FLAG_UP_PARTIALS 0x1u /* also reserve the pre-order ("up") partials used by edge derivatives */
#define PHB_FLAG_NO_PARTIALS 0x2u /* lnL-only context: no per-node partial storage (phb_lnl_resident only) */

/* phb_compute_partials modes */
#define PHB_MODE_AUTO 0
#define PHB_MODE_TILE 1  /* one launch: every CTA owns a pattern tile and walks the whole schedule */
#define PHB_MODE_LEVEL 2 /* one launch per tree level (all rows of a level are independent)        */
#define PHB_MODE_RESIDENT 3 /* 4-state only: one launch, operands stay on chip, blocks streamed out   */

typedef struct phb_ctx phb_ctx;

/* ---- library ------------------------------------------------------------------------- */
int phb_version(void);
const char* phb_status_name(int status);
/* message of the last failure on this ctx, or (ctx == NULL) of the last failed phb_create /
 * context-free call on the calling thread */
const char* phb_last_error(const phb_ctx* ctx);
/* The PHB_* developer switches (A/B measurements; csrc/common.cuh `Tuning`) are read from the environment once, on
 * first use; this re-reads them (tests and tuning scripts that change a switch inside one process). */
int phb_reload_tuning(void);
/* number of kernel launches issued by this ctx since creation (bench.py's gpu_launches) */
int64_t phb_launch_count(const phb_ctx* ctx);

/* ---- host-only numerics --------------------------------------------------------------- */
/* Yang (1994) discrete gamma; stands in for phylo_utils.discrete_gamma.discrete_gamma
 * (src/discrete_gamma.pyx:30-47 -> src/c_discrete_gamma.c:285-321 DiscreteGamma).
 * rates[ncat], weights[ncat] are outputs; mean rate is alpha/beta. */
int phb_discrete_gamma(double alpha, double beta, int ncat, int use_median, double* rates, double* weights);

/* ---- site-pattern compression (context-free) --------------------------------------------
 * Device form of alignment_to_numpy's np.unique(axis=1, return_inverse, return_counts)
 * (alignment/alignment.py:40-57) fused with the charmap look-up of seq_to_partials (:26-37).
 * data[n_tips][n_sites]: state-set codes (byte_table == NULL), or raw characters that byte_table[256] maps to
 * codes (255 = not in the charmap: PHB_ERR_INVALID, flat index of the first offender in *bad_index_out -
 * the reference raises KeyError there).  Codes must be ranks of the charmap's 0/1 rows in lexicographic order
 * (charmaps.CodeBook) for the pattern order to equal the reference's.
 * Outputs (host, caller-allocated for the worst case n_patterns == n_sites): patterns_out[n_tips][n_patterns]
 * (dense, row length = *n_patterns_out), weights_out[n_patterns] (siteweights), inverse_out[n_sites]. */
int phb_compress_patterns(int device, const uint8_t* data, const uint8_t* byte_table, int n_tips, int64_t n_sites,
                          uint8_t* patterns_out, int64_t* weights_out, int64_t* inverse_out, int64_t* n_patterns_out,
                          int64_t* bad_index_out);

/* ---- context --------------------------------------------------------------------------
 * Replaces the array allocation of TreeModel.initialise (phylo_utils/tree_model.py:101-132).
 * `workspace` is device memory owned by the caller (e.g. a torch uint8 tensor's data_ptr) of at
 * least phb_workspace_bytes(); pass NULL to let the library cudaMalloc it.  `stream` is a
 * cudaStream_t (NULL = default stream). */
size_t phb_workspace_bytes(int n_tips, int64_t n_patterns, int n_cat, int n_states, unsigned flags);
int phb_create(int device, int n_tips, int64_t n_patterns, int n_cat, int n_states, unsigned flags, void* workspace,
               size_t workspace_bytes, void* stream, phb_ctx** out);
int phb_destroy(phb_ctx* ctx);
int phb_sync(phb_ctx* ctx);

/* ---- inputs ---------------------------------------------------------------------------- */
/* Tip data: replaces the K-fold replicated fp64 tip copy of TreeModel.initialise
 * (tree_model.py:142-148).  codes[n_tips][n_patterns] indexes lut[n_codes][n_states]
 * (rows are usually 0/1 state sets but any non-negative doubles are allowed).
 * tip_nodes[n_tips] gives the node id of each codes row (Traversal.names ∘ TreeModel.names).
 * `codes_on_device` != 0: codes is a device pointer (copied device-to-device into the pitched workspace buffer). */
int phb_set_tips(phb_ctx* ctx, const uint8_t* codes, int codes_on_device, int n_codes, const double* lut,
                 const int32_t* tip_nodes);
/* pattern multiplicities (alignment.py:48-51 `siteweights`); NULL = all ones */
int phb_set_pattern_weights(phb_ctx* ctx, const int64_t* weights);
/* Eigen-decomposed rate matrix + mixture: model.eigen.{evecs,evals,ivecs}, model.freqs
 * (substitution_models/abstract.py:88-122) and rate_model.{rates,weights} (rate_models.py:4-13).
 * evecs/ivecs are [A][A] row-major (pass np.ascontiguousarray(ivecs): the reference keeps ivecs in F order). */
int phb_set_model(phb_ctx* ctx, const double* evecs, const double* evals, const double* ivecs, const double* freqs,
                  const double* rates, const double* cat_weights);
/* Frequencies / mixture only, for models whose P matrices are supplied ready-made */
int phb_set_mixture(phb_ctx* ctx, const double* freqs, const double* rates, const double* cat_weights);
/* Schedule: rows[n_rows][3] = {PAR, CH1, CH2} in an order where children precede parents
 * (Traversal.postorder_traversal, utils.py:127-134, or any re-ordering of it).
 * level_offsets[n_levels+1] (optional, may be NULL) marks groups of mutually independent rows
 * for PHB_MODE_LEVEL. */
int phb_set_schedule(phb_ctx* ctx, int n_rows, const int32_t* rows, int n_levels, const int32_t* level_offsets);
/* lengths[n_rows][2]: branch lengths PAR-CH1, PAR-CH2 of every row (Traversal.brlens lookups at
 * tree_model.py:166-167) */
int phb_set_edge_lengths(phb_ctx* ctx, const double* lengths);

/* ---- transition matrices ---------------------------------------------------------------- */
/* P(t r_k) = V diag(exp(lambda t r_k)) V^-1 for every row x child x category in one launch;
 * replaces the 2(N-2) host calls of Model.p at tree_model.py:168-169 (abstract.py:49-59). */
int phb_build_pmatrices(phb_ctx* ctx);
/* Ready-made P[n_rows][2][K][A][A] from the host (non-reversible models that go through expm,
 * abstract.py:173-192) */
int phb_set_pmatrices(phb_ctx* ctx, const double* pmats);
/* debugging / parity: copy back P of (row, child) -> out[K][A][A] */
int phb_get_pmatrix(phb_ctx* ctx, int row, int child, double* out);

/* ---- the hot path ------------------------------------------------------------------------ */
/* Post-order pruning over all rows: TreeModel.compute_partials (tree_model.py:160-176), each row
 * being one `clv` gufunc call (likelihood/numba_likelihood_engine.py:10-46). */
int phb_compute_partials(phb_ctx* ctx, int mode);
/* Virtual root on edge (node_a, node_b) of the given length, per-category root likelihoods, mixture,
 * log and weighted sum: TreeModel.compute_partials_at_edge + compute_likelihood_at_edge
 * (tree_model.py:178-217) with lnl_node (numba_likelihood_engine.py:82-87) and the logsumexp mix.
 * root_pmats: optional host P[2][K][A][A] (P for a, P for b); NULL = built on device as P(0), P(length).
 * total: sum_p weight_p * lnl_p.  pattern_lnl[S] (host, optional): per-pattern log-likelihood.
 * cat_lnl[S][K] (host, optional): per-pattern per-category log-likelihood before mixing (lnl_node's output). */
int phb_root_lnl(phb_ctx* ctx, int node_a, int node_b, double length, const double* root_pmats, double* total,
                 double* pattern_lnl, double* cat_lnl);
/* Whole evaluation without storing per-node partials: P build + pruning + root + reduction in
 * pattern-tile resident kernels; same result as phb_build_pmatrices + phb_compute_partials +
 * phb_root_lnl.  Works on contexts created with or without PHB_FLAG_NO_PARTIALS. */
int phb_lnl_resident(phb_ctx* ctx, int node_a, int node_b, double length, double* total, double* pattern_lnl);

/* The same evaluation starting from HOST tip codes (same layout and look-up table as the last phb_set_tips):
 * the pattern axis is cut into n_chunks pieces (0 = default) and the host->device copy of piece i+1 overlaps
 * the pruning of piece i.  Pass pinned memory for the copies to be truly asynchronous.  Replaces
 * "TreeModel.set_alignment + initialise + compute_likelihood_at_edge" for a changed alignment. */
int phb_lnl_from_host(phb_ctx* ctx, const uint8_t* codes, int n_chunks, int node_a, int node_b, double length,
                      double* total, double* pattern_lnl);

/* phb_lnl_from_host with the tip codes packed two per byte: packed_codes[n_tips][(n_patterns + 1) / 2], pattern 2j in
 * the low nibble of byte j, pattern 2j+1 in the high nibble (look-up tables of at most 16 rows, i.e. every
 * nucleotide alphabet including the IUPAC ambiguity codes of alignment/charmaps.py:2-20).  Half the bytes cross
 * PCIe; the kernel reads the nibbles directly.  Afterwards the device holds PACKED codes: phb_lnl_resident keeps
 * working, everything that reads tips otherwise returns PHB_ERR_STATE until the next phb_set_tips. */
int phb_lnl_from_host_packed(phb_ctx* ctx, const uint8_t* packed_codes, int n_chunks, int node_a, int node_b,
                             double length, double* total, double* pattern_lnl);
/* host helper: codes[n_tips][n_patterns] (values < 16) -> out[n_tips][(n_patterns + 1) / 2] as described above */
int phb_pack_codes(const uint8_t* codes, int n_tips, int64_t n_patterns, uint8_t* out);

/* phb_lnl_from_host for look-up tables of at most 8 rows (any alignment without partial ambiguity codes: the four
 * nucleotides, the gap / N row, up to three more) and a reversible model: 3 bits per code, split into a plane of 2-bit
 * values low_plane[n_tips][(n_patterns + 3) / 4] (pattern s in bits 2 (s % 4) .. of byte s / 4) and a plane of high
 * bits high_plane[n_tips][(n_patterns + 7) / 8] (pattern s in bit s % 8 of byte s / 8) - 3/8 of a byte per code over
 * PCIe, where the host-to-device link bounds the evaluation once several GPUs of a box are fed at the same time.
 * Afterwards the device holds split codes (as after phb_lnl_from_host_packed). phb_split_codes builds the planes. */
int phb_lnl_from_host_split(phb_ctx* ctx, const uint8_t* low_plane, const uint8_t* high_plane, int n_chunks, int node_a,
                            int node_b, double length, double* total, double* pattern_lnl);
int phb_split_codes(const uint8_t* codes, int n_tips, int64_t n_patterns, uint8_t* low_plane, uint8_t* high_plane);

/* ---- read-back for parity tests (TreeModel.partials / .scale / .root_partials attributes) --- */
/* out[S][K][A]; tips are expanded from their codes */
int phb_get_partials(phb_ctx* ctx, int node, double* out);
/* out[S][K] natural-log scalers, i.e. exponent * ln 2 repeated over categories */
int phb_get_scalers(phb_ctx* ctx, int node, double* out);
/* partials / scalers of the last phb_root_lnl virtual root */
int phb_get_root_partials(phb_ctx* ctx, double* out_partials, double* out_scalers);

/* ---- derivatives ------------------------------------------------------------------------- */
/* Pre-order pass: for every non-root-child node the partial of everything outside its subtree
 * ("up" partial), the operand lnl_branch_derivs needs at the far end of each edge; device-side
 * equivalent of walking Traversal.optimising_traversal (utils.py:137-188) without re-rooting in place.
 * (node_a, node_b, length) is the root edge the down partials were computed for
 * (Traversal.root_edge).  Needs PHB_FLAG_UP_PARTIALS and a reversible model.
 * What the pass leaves in the "up" storage is private to the library and only consumed by
 * phb_edge_derivatives: up partials, or - where a kernel can form it on the way - the per-edge sum table
 * s_km = (V^-1 down)_m (V^T (pi * up))_m from which every later derivative pass reads one block per edge. */
int phb_compute_up_partials(phb_ctx* ctx, int node_a, int node_b, double length);
/* For each listed node (edge above it; for a root child: the root edge), at the given trial length:
 * out[i] = { lnL, d lnL / dt, d2 lnL / dt2 } summed over patterns with their weights, Gamma mixture
 * composed as in SURVEY.md 8(a) row a12 from lnl_branch_derivs (numba_likelihood_engine.py:49-57).
 * chain_rule != 0 applies the r_k, r_k^2 factors (true d/dt); 0 reproduces Model.dp_dt / d2p_dt2's
 * convention (abstract.py:61-77). */
int phb_edge_derivatives(phb_ctx* ctx, int n_edges, const int32_t* nodes, const double* lengths, int chain_rule,
                         double* out);

/* ---- re-rooting in place -------------------------------------------------------------------
 * The reference's one-edge-at-a-time optimisation order (Traversal.optimising_traversal, utils.py:137-188; rows
 * [PAR, SIB, GPA, NOD, PAR] and [NOD, CH1, CH2, -1, -1]) re-computes one node's partial from two neighbours so that it
 * faces the edge about to be optimised.  phb_update_node is that single `clv` call on the device
 * (numba_likelihood_engine.py:10-46 with P(len_a), P(len_b) built on the device): node's block := combine(child_a over
 * len_a, child_b over len_b), children being tips or internal nodes as they currently stand.
 * phb_branch_derivatives evaluates { lnL, dlnL/dt, d2lnL/dt2 } ACROSS the edge (node_a, node_b) at n trial lengths in
 * one launch, from the two nodes' current partials (lnl_branch_derivs, :49-57, composed over the Gamma mixture): the
 * line search of one edge.  Neither needs the pre-order pass; both invalidate what it left. */
int phb_update_node(phb_ctx* ctx, int node, int child_a, double len_a, int child_b, double len_b);
int phb_branch_derivatives(phb_ctx* ctx, int node_a, int node_b, int n_lengths, const double* lengths, int chain_rule,
                           double* out);

/* ---- stream-ordered forms (multi-GPU drivers) ------------------------------------------------
 * The reference is one process and sums per-site lnL on the host (bin/phy.py:146).  With the site patterns sharded
 * over several GPUs the only exchange is the sum of the per-shard scalars, and it must not cost a
 * host round trip per GPU: these calls do the same work as phb_lnl_resident / phb_root_lnl /
 * phb_lnl_from_host_packed / phb_edge_derivatives but only ENQUEUE it on the context's stream and leave the sums in
 * the context's device result buffer - lnL at [0], or { lnL, d1, d2 } of edge i at [3 i .. 3 i + 2] - so that a
 * collective (ncclAllReduce on the same stream, or torch.distributed on a tensor view of the buffer) can follow
 * directly.  phb_result_fetch copies the first n doubles back and synchronises; it (or phb_sync) also completes a
 * host-fed evaluation (reports a chunk of tip codes that never arrived). */
int phb_lnl_resident_async(phb_ctx* ctx, int node_a, int node_b, double length);
int phb_root_lnl_async(phb_ctx* ctx, int node_a, int node_b, double length);
int phb_lnl_from_host_packed_async(phb_ctx* ctx, const uint8_t* packed_codes, int n_chunks, int node_a, int node_b,
                                   double length);
int phb_lnl_from_host_split_async(phb_ctx* ctx, const uint8_t* low_plane, const uint8_t* high_plane, int n_chunks,
                                  int node_a, int node_b, double length);
int phb_edge_derivatives_async(phb_ctx* ctx, int n_edges, const int32_t* nodes, const double* lengths, int chain_rule);
/* Pipelined host-fed evaluations (many alignments over one tree - bootstrap replicates, simulated data sets): up to TWO
 * in flight.  phb_lnl_from_host_submit enqueues the copy of a new alignment (two codes per byte when high_plane is
 * NULL, else the split 3-bit planes) into the code slot that is free and its walk behind whatever the stream holds;
 * the copy engine works under the walk of the evaluation before.  The sum lands in device result[*slot_out] (an
 * all-reduce may follow on the stream); phb_result_post enqueues its copy to the host, phb_result_wait blocks for
 * that one evaluation only.  Post evaluation i before submitting i + 2 (its slot and result word are reused). */
int phb_lnl_from_host_submit(phb_ctx* ctx, const uint8_t* codes, const uint8_t* high_plane, int n_chunks, int node_a,
                             int node_b, double length, int* slot_out);
int phb_result_post(phb_ctx* ctx, int slot);
int phb_result_wait(phb_ctx* ctx, int slot, double* out);
/* device address and capacity (in doubles) of the result buffer; it lies inside the caller's workspace when one
 * was given to phb_create */
int phb_device_result(phb_ctx* ctx, void** device_ptr, int64_t* capacity_doubles);
int phb_result_fetch(phb_ctx* ctx, int n, double* out);

/* ---- the scalar sum over the ranks of ONE box inside the reduction kernel -----------------------------------------
 * (no counterpart in the reference: it is one process, bin/phy.py:146.)  At 2 ms per evaluation on eight GPUs a
 * collective-library call behind every evaluation is 2-3 % of the step.  Instead every rank owns a small exchange
 * buffer (phb_peer_buffer returns its CUDA IPC handle, PHB_PEER_HANDLE_BYTES bytes), maps the buffers of all ranks
 * (phb_peer_connect: `handles` = world handles in rank order, exchanged by the caller - e.g. one all_gather at set-up)
 * and, when phb_peer_sum_next was called right before a stream-ordered scalar-lnL entry point (phb_lnl_resident_async,
 * phb_root_lnl_async, phb_lnl_from_host_packed_async / _split_async / _submit), the kernel that reduces the per-CTA
 * sums also stores the rank's total into every peer's buffer (NVLink), waits for the peers' totals and adds them in
 * rank order: result[0] (or result[slot]) then holds the GLOBAL lnL, bit-identical on every rank.  Every rank must
 * make the same sequence of such calls.  The wait is bounded (about a minute); a rank that never arrives yields NaN. */
#define PHB_PEER_HANDLE_BYTES 64
int phb_peer_buffer(phb_ctx* ctx, void* handle_out);
int phb_peer_connect(phb_ctx* ctx, int rank, int world, const void* handles);
int phb_peer_sum_next(phb_ctx* ctx);

/* ---- stand-alone operators (host arrays in, host arrays out; reference-exact semantics) ------ */
/* clv gufunc (numba_likelihood_engine.py:10-46): per-(site,category) natural-log scalers, rescale by the
 * category maximum when 0 < max < 2^-128.  p1,p2 [K][A][A]; clv1,clv2,out [S][K][A]; scalers [S][K]. */
int phb_op_clv(int device, int64_t S, int K, int A, const double* p1, const double* p2, const double* clv1,
               const double* clv2, const double* scaler_a, const double* scaler_b, double* cml_scaler, double* out);
/* lnl_node (:82-87): pi[A], partials[S][K][A], scale[S][K] -> out[S][K] */
int phb_op_lnl_node(int device, int64_t S, int K, int A, const double* pi, const double* partials,
                    const double* scale, double* out);
/* lnl_branch (:60-79) with n_derivs = 0 (probs[1][A][A], out[S][1]) and lnl_branch_derivs (:49-57) with
 * n_derivs = 2 (probs[3][A][A], out[S][3]); partials_a/b [S][A], scale_a/b [S] */
int phb_op_lnl_branch(int device, int64_t S, int A, int n_derivs, const double* probs, const double* pi,
                      const double* partials_a, const double* partials_b, const double* scale_a,
                      const double* scale_b, double* out);
/* Model.p / dp_dt / d2p_dt2 (abstract.py:49-77) for a batch of scaled times: out[n][A][A] =
 * V diag(lambda^order exp(lambda t_i)) V^-1 */
int phb_op_pmatrices(int device, int A, int n, const double* evecs, const double* evals, const double* ivecs,
                     const double* times, int order, double* out);

/* Measured fp64 throughput of the device in TFLOP/s: kind 0 = vector pipe (independent DFMA chains on every SM),
 * kind 1 = fp64 tensor pipe (DMMA m8n8k4 chains).  Measurement aid: the denominators of the compute-bound rooflines
 * bench.py reports (no counterpart in the reference). */
int phb_op_fp64_peak(int device, int kind, double* tflops);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* PHYLO_B200_H */
