// Shared declarations of the phylo_b200 engine: context, schedule rows, launch helpers.
#pragma once

#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

#include "../../include/phylo_b200.h"

namespace phb {

// Partials smaller than this get rescaled (same threshold as the reference's SCALE_THRESHOLD,
// likelihood/numba_likelihood_engine.py:7); rescaling multiplies by an exact power of two.
constexpr double kScaleThreshold = 2.938735877055718769921841343056e-39;  // 2^-128
// Device-side invariant checks of the walks (slot indices, tile ranges, code-row bounds, producer-before-consumer order):
// compiled in by `python phylo_utils_b200/csrc/build.py --checks` (-DPHB_DEVICE_CHECKS -> libphylo_b200_checks.so), nothing
// in the shipped build.  compute-sanitizer is closed on this pool; the parity tests are run once per round against the
// checked build instead (PHB_LIBRARY=.../libphylo_b200_checks.so python -m pytest tests -m gpu).
#ifdef PHB_DEVICE_CHECKS
#include <cassert>
#define PHB_DCHECK(cond) assert(cond)
#else
#define PHB_DCHECK(cond) ((void)0)
#endif

constexpr int kScaleThresholdHi = 0x37F00000;                              // high word of 2^-128
constexpr double kLn2 = 0.693147180559945309417232121458;

// One row of the pruning schedule as the kernels see it.
// kind: where a child's partial comes from.  SRC_PREV = "the block written by the row just before
// this one" (src still holds its slot): tile-resident kernels read it from on-chip storage,
// level-order kernels treat it as SRC_GLOBAL.  The host canonicalises every row so that
// rank(kind[0]) <= rank(kind[1]) with TIP < PREV < GLOBAL (children commute), which leaves five
// row shapes: TT, TP, TG, PG, GG.
enum : int32_t { SRC_GLOBAL = 0, SRC_TIP = 1, SRC_PREV = 2, SRC_SUMTABLE = 3 };   // SRC_SUMTABLE: derivative edges only
struct __align__(16) OpRow {
    int32_t dst;      // internal slot written by this row
    int32_t src[2];   // tip row or internal slot of each child
    int32_t kind[2];  // SRC_*
    int32_t pidx[2];  // index of each child's P matrix block ([K][A][A]) in the P buffer
    int32_t pad;
};
static_assert(sizeof(OpRow) == 32, "OpRow must stay 32 bytes");

struct Ctx;

// Developer switches (A/B measurements, what-if builds).  The environment is read ONCE, on first use, into this
// struct (api.cu); nothing on an evaluation's path calls getenv.  Defaults = the shipped configuration.
struct Tuning {
    bool disable_mma = false;        // PHB_DISABLE_MMA: generic kernels instead of the FP64 tensor-core ones
    bool disable_tiptab = false;     // PHB_DISABLE_TIPTAB: DMMA rows multiply tip operands instead of reading P.lut rows
    bool up_two_rows = false;        // PHB_UP_TWO_ROWS: pre-order pass as two pruning rows per parent
    bool up_plain = false;           // PHB_UP_PLAIN: pre-order walk writes up partials, not per-edge sum tables
    bool deriv_no_st = false;        // PHB_DERIV_NO_ST: DMMA derivative pass leaves no sum tables
    bool deriv_matrix_form = false;  // PHB_DERIV_MATRIX_FORM: matrix-form derivative kernels
    bool compress_timing = false;    // PHB_COMPRESS_TIMING: per-phase wall clock of phb_compress_patterns on stderr
    int pair_ctas = 0;               // PHB_PAIR_CTAS: cap on resident warps per SM of the pair kernel
    int pair_ppt = 0;                // PHB_PAIR_PPT: patterns per lane of the lnL-only pair kernel (0 = choose)
    bool pair_one_warp_ctas = false; // PHB_PAIR_ONE_WARP_CTAS: the lnL-only walk always as 1-warp CTAs (12 per SM), never one CTA of 12 warps per SM
    int pair_cta_rounds = 0;         // PHB_PAIR_CTA_ROUNDS: rounds of tiles from which on the one-CTA-per-SM form is used (0 = 6; -1 = always)
    int pair_stagger = 0;            // PHB_PAIR_STAGGER: start offset between the warps of the one-CTA-per-SM form, ns (experiment)
    int pair_grid = 0;               // PHB_PAIR_GRID: 0 = choose, 1 = every resident warp, 2 = equal tiles per warp
    bool pair_full_p = false;        // PHB_PAIR_FULL_P: the lnL-only walk reads full P blocks even for reversible models
    int up_ppt = 0;                  // PHB_UP_PPT: patterns per lane of the pre-order walk
    int up_warps = 0;                // PHB_UP_WARPS: cap on resident warps per SM of the pre-order walk
    int tile_want = 0;               // PHB_TILE_WANT: tiles per SM the streaming tile walk asks for
    int mma_variant = 0;             // PHB_MMA_VARIANT: alternative DMMA tile shapes
};
const Tuning& tuning();

// launch bookkeeping ------------------------------------------------------------------------
#define PHB_CUDA(ctx, expr)                                                                     \
    do {                                                                                        \
        cudaError_t _e = (expr);                                                                \
        if (_e != cudaSuccess) return (ctx)->fail_cuda(_e, #expr, __FILE__, __LINE__);          \
    } while (0)

#define PHB_REQUIRE(ctx, cond, code, msg)                                                       \
    do {                                                                                        \
        if (!(cond)) return (ctx)->fail((code), (msg));                                         \
    } while (0)

void set_thread_error(const std::string& msg);
const char* thread_error();

struct Ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    int n_tips = 0, K = 0, A = 0, n_codes = 0;
    int64_t S = 0;
    unsigned flags = 0;
    int n_nodes = 0, n_internal = 0, sm_count = 148;
    size_t smem_optin = 0;

    // workspace
    bool owns_ws = false;
    uint8_t* ws = nullptr;
    size_t ws_bytes = 0;

    // device views (all inside ws unless noted)
    // tip codes, one row per tip with a 128-byte multiple pitch (rows padded with zeros) so that tiles can
    // be fetched with aligned 16-byte asynchronous copies and may over-read up to the pitch
    const uint8_t* d_codes = nullptr;
    uint8_t* d_codes_ws = nullptr;
    size_t code_pitch = 0;
    // true after phb_lnl_from_host_packed: the buffer holds two 4-bit codes per byte (rows of code_pitch / 2
    // bytes), which only the pair kernel reads; every other consumer asks for phb_set_tips first
    bool codes_packed = false;
    int codes_mode = 0;                // 0 one byte per code, 1 nibbles, 2 split 3-bit planes (pair_common.cuh CODES_*)
    double* d_lut = nullptr;           // [256][A]
    double* d_weights = nullptr;       // [S]
    double* d_clv = nullptr;           // [n_internal][S][K][A]
    int32_t* d_scale = nullptr;        // [n_internal][S]
    // "up" (pre-order) partials live in the SAME block array as the down partials, right behind them:
    // block n_internal + node_id.  One array, one kernel family.
    double* d_up = nullptr;            // = d_clv + n_internal * stride, [n_nodes][S][K][A] (optional)
    int32_t* d_up_scale = nullptr;     // = d_scale + n_internal * S
    OpRow* d_up_rows = nullptr;        // [2 * max_rows]
    std::vector<OpRow> up_rows;
    std::vector<int32_t> up_levels;
    int root_block = 0;                // index of the virtual-root block in the shared block array
    double* d_root_clv = nullptr;      // [S][K][A] = block root_block
    int32_t* d_root_scale = nullptr;   // [S]
    double* d_pmats = nullptr;         // [2*max_rows + 2][K][A][A]
    double* d_dmats = nullptr;         // derivative scratch: [3][edges per launch][K][A][A], or the sum-table coefficients
    size_t dmats_doubles = 0;
    void* d_edges = nullptr;           // [n_nodes] edge operand descriptors of a derivative launch
    // 4-state models: T[m][k][code][:] = P[m][k] . lut[code] for every P block m, so that a tip operand is
    // a 32-byte table look-up instead of a matrix-vector product (codes padded to kTipTabCodes rows)
    double* d_tiptab = nullptr;
    // 4-state reversible models: R[m][k] = packed upper triangle of diag(pi) P[m][k] (10 doubles; symmetric by detailed
    // balance), what the lnL-only walk reads instead of P (clv_dna_pair.cu, SYM); rebuilt with the tip tables
    double* d_rmats = nullptr;
    bool reversible = false;           // pi_i q_ij == pi_j q_ji for the eigen-system and frequencies of phb_set_model
    // 61-state models: zero-padded, 16-byte aligned staging images of every P block and tip table for the DMMA
    // kernels ([mat][K][2][pimg_rows][pimg_pitch], clv_mma.cu), rebuilt with the matrices
    double* d_pimg = nullptr;
    int pimg_rows = 0, pimg_pitch = 0;
    double* d_model = nullptr;         // evecs | evals | ivecs | freqs | rates | catw
    double* d_lengths = nullptr;       // [2*max_rows + 2]
    OpRow* d_rows = nullptr;           // [max_rows]
    void* d_res_rows = nullptr;        // [max_rows + 1] 16-byte descriptors of the resident kernel
    // what d_res_rows currently holds: a plan is a function of the schedule, the tip layout and the root edge only, so
    // repeated evaluations (new branch lengths, same tree) skip planning and upload
    int64_t sched_gen = 0;             // bumped whenever the schedule or the tip layout changes
    struct {
        int kind = 0;                  // 0 = nothing cached, 1 = pair lnL-only plan, 2 = pair store plan
        int root_a = -1, root_b = -1;
        int64_t gen = -1;
        int n_steps = 0, n_slots = 0;
        bool sym = false;              // the descriptors point at the symmetric P blocks (d_rmats)
    } res_cache;
    double h_root_two[2] = {0.0, 0.0}; // P(0), P(root length): source of the asynchronous copy behind the row lengths
    OpRow h_spare_row{};               // the one-row schedule of phb_update_node (same reason)
    int resident_slots = 0;            // parked blocks the last resident launch needed
    int resident_warps = 0;            // warps per SM of the last resident launch
    unsigned char* d_scratch = nullptr;  // L2-resident parking area of the lnL-only resident kernel
    size_t scratch_bytes = 0;
    size_t smem_per_sm = 0;
    // host->device pipelining of phb_lnl_from_host
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t chunk_events[32] = {};
    cudaEvent_t start_event = nullptr;
    // single-launch pipelining (clv_dna_pair.cu): per-chunk arrival flags + error word on the device, the epoch the
    // copy engine stamps them with (pinned host word), and whether an evaluation's error word is still unread
    int* d_flags = nullptr;            // [2][kMaxFlagChunks + 1]: one set per code slot
    int* d_flags_cur = nullptr;        // the set the next pair-kernel launch waits on
    // pipelined host-fed evaluations (phb_lnl_from_host_submit): two code slots in the tip-code buffer (packed formats
    // need at most half of it), so that the copy of evaluation i+1 runs under the walk of evaluation i
    cudaEvent_t slot_done[2] = {};     // the walk that read slot s has finished (the copy stream waits for it)
    cudaEvent_t result_event[2] = {};  // result s (and its error word) has reached h_results
    cudaEvent_t copies_done[2] = {};   // every copy of the evaluation in slot s (and its flags) has executed
    double* h_results = nullptr;       // pinned: [2] sums, then [2] error words (as doubles' worth of ints)
    int next_slot = 0;
    int* h_epoch = nullptr;
    int flag_epoch = 0;
    bool pipelined_pending = false;
    // sum of a scalar result over the ranks of one box INSIDE the reduction kernel (phb_peer_*): every rank owns a small
    // exchange buffer, has the peers' buffers mapped (CUDA IPC over NVLink) and writes its value into all of them
    struct PeerLink {
        void* own = nullptr;            // [2 parities][kMaxPeers] cells {value, sequence number} in this device's memory
        void* cells[16] = {};           // rank r's buffer as this process sees it (own for r == rank)
        int rank = 0, world = 0;
        unsigned long long epoch = 0;   // exchanges done; every rank runs the same sequence of them
        bool connected = false;
        bool armed = false;             // phb_peer_sum_next: the next scalar-lnL entry point uses the exchange
        bool use_now = false;           // ... and this is that entry point
    } peer;
    double* d_pattern_lnl = nullptr;   // [S]
    double* d_cat_lnl = nullptr;       // [S][K]
    double* d_partial_sums = nullptr;  // [kPartialCap]
    double* d_result = nullptr;        // [result_doubles] = max(4 * kMaxEdgeBatch, 6 * n_tips): lnL, or 3 sums per edge
    size_t result_doubles = 0;

    // host mirrors
    std::vector<int32_t> node_tip;     // node id -> tip row, or -1
    std::vector<int32_t> node_slot;    // node id -> internal slot, or -1
    std::vector<int32_t> node_parent;  // node id -> parent node id (-1 for root children)
    std::vector<int32_t> node_row;     // node id -> schedule row that computes it (-1 for tips)
    std::vector<OpRow> rows;
    std::vector<int32_t> rows_raw;     // PAR, CH1, CH2 as given
    std::vector<int32_t> level_offsets;
    std::vector<double> lengths;       // [n_rows][2]
    std::vector<double> h_evecs, h_ivecs, h_freqs;   // host copies of the eigenvectors / frequencies (sum-table derivatives)
    bool have_tips = false, have_model = false, have_mixture = false, have_schedule = false;
    bool have_lengths = false, have_pmats = false, have_partials = false, have_up = false;
    bool have_root = false;
    std::vector<char> st_ready;        // per node: its up block already holds the edge's sum table (DMMA derivative path)
    bool up_sumtable = false;          // the up blocks hold per-edge sum tables (up_dna_pair.cu), not up partials
    bool resident_partials = false;    // the last post-order pass was the operand-resident walk (the pre-order pass follows suit)
    int root_a = -1, root_b = -1;
    double root_len = 0;

    std::string err;
    int64_t launches = 0;

    int fail(int code, const std::string& msg) {
        err = msg;
        return code;
    }
    int fail_cuda(cudaError_t e, const char* what, const char* file, int line) {
        char buf[512];
        snprintf(buf, sizeof buf, "CUDA error %d (%s) at %s:%d: %s", (int)e, cudaGetErrorString(e), file, line, what);
        err = buf;
        return PHB_ERR_CUDA;
    }
    int max_rows() const { return n_internal > 0 ? n_internal : 1; }
    int n_rows() const { return (int)rows.size(); }
    size_t clv_stride() const { return (size_t)S * K * A; }  // doubles per node
    double* model_evecs() const { return d_model; }
    double* model_evals() const { return d_model + (size_t)A * A; }
    double* model_ivecs() const { return d_model + (size_t)A * A + A; }
    double* model_freqs() const { return d_model + 2 * (size_t)A * A + A; }
    double* model_rates() const { return d_model + 2 * (size_t)A * A + 2 * A; }
    double* model_catw() const { return d_model + 2 * (size_t)A * A + 2 * A + K; }
};

constexpr int kMaxReduceBlocks = 4096;
constexpr int kPartialCap = 65536;   // doubles in the block-sum buffer
constexpr int kMaxEdgeBatch = 64;
constexpr int kMaxPeers = 16;
constexpr int kMaxChunks = 32;
constexpr int kMaxFlagChunks = 255;
constexpr int kTipTabCodes = 16;     // tip tables cover look-up tables of up to 16 rows (IUPAC DNA has 15)

// kernel families (each returns a phb_status) ---------------------------------------------------
// pmatrix.cu
int launch_build_pmatrices(Ctx* c, const double* d_lengths, int n_mats, double* d_out, int order, int chain_rule);
int launch_tip_tables(Ctx* c, int first_mat, int n_mats);
// rows per category block of the tip tables: 8 or 16 for 4-state models, a multiple of 8 up to 64 otherwise
inline int tip_table_rows(const Ctx* c) {
    if (c->A == 4) return c->n_codes <= 8 ? 8 : kTipTabCodes;
    return (c->n_codes + 7) / 8 * 8;
}
inline bool tip_tables_usable(const Ctx* c) {
    if (c->d_tiptab == nullptr || !c->have_tips) return false;
    return c->A == 4 ? c->n_codes <= kTipTabCodes : c->n_codes <= 64;
}
cudaError_t launch_pmatrix_raw(cudaStream_t stream, const double* evecs, const double* evals, const double* ivecs,
                               const double* rates, const double* d_lengths, double* d_out, int A, int K, int n_mats,
                               int order, int chain_rule);
// clv_dna.cu  (A == 4, K in {1,2,4,8})
// A set of rows to execute: device array, count and (for PHB_MODE_LEVEL) offsets of independent groups.
struct RowSet {
    const OpRow* d_rows;
    int n_rows;
    const std::vector<int32_t>* levels;  // may be null / empty in tile mode
};
bool dna_supported(const Ctx* c);
int dna_run_rows(Ctx* c, const RowSet& rs, int mode);
int dna_root(Ctx* c, int a, int b, bool want_cat, bool store_root);
// clv_dna_pair.cu: lnL-only walk, two patterns per lane (the default lnL-only path)
int dna_pair_lnl(Ctx* c, int root_a, int root_b);
int dna_pair_store(Ctx* c);   // all partials stored; PHB_ERR_UNSUPPORTED (no message) when the shape is not covered
// slot < 0: immediate form (slot 0, ordered behind everything queued); slot 0 / 1: pipelined form, sum -> d_result[slot]
int dna_pair_from_host(Ctx* c, const uint8_t* codes_host, const uint8_t* codes_hi_host, int mode, int n_chunks, int root_a,
                       int root_b, int slot = -1);
// up_dna_pair.cu: pre-order pass as one operand-resident walk; PHB_ERR_UNSUPPORTED (no message) when not covered
int dna_up_walk(Ctx* c, int node_a, int node_b);
// clv_generic.cu (any A <= 64, any K <= 16)
int generic_run_rows(Ctx* c, const RowSet& rs, int mode);
// clv_mma.cu (A == 20 or 61, FP64 tensor cores)
bool mma_supported(const Ctx* c);
int mma_run_rows(Ctx* c, const RowSet& rs, int mode);
int launch_mma_images(Ctx* c, int first_mat, int n_mats);   // after every (re)build of P blocks / tip tables
int mma_run_parent_rows(Ctx* c, const std::vector<int32_t>& parents, const std::vector<int32_t>& levels);   // pre-order pass
// picks the kernel family for this context's shape
int run_rows(Ctx* c, const RowSet& rs, int mode);
int generic_root(Ctx* c, int a, int b, bool want_cat, bool store_root);
// derivs.cu
int launch_up_partials(Ctx* c, int node_a, int node_b);
int launch_edge_derivatives(Ctx* c, int n_edges, const int32_t* nodes, const double* lengths, int chain_rule,
                            double* out, const int32_t* far_nodes = nullptr);
// max over a byte array (in clv_generic.cu); synchronises the stream
int launch_max_code(Ctx* c, const uint8_t* d_codes, size_t n, int* worst);
// reduce (in clv_generic.cu): sums n_parts partial sums (stride 1) into d_result[0..n_out)
int launch_final_reduce(Ctx* c, const double* d_parts, int n_parts, int n_out, double* d_out);

// device helpers ---------------------------------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ void ld256(const double* p, double (&v)[4]) {
    asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];"
                 : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[3])
                 : "l"(p)
                 : "memory");
}
__device__ __forceinline__ void ld256_nc(const double* p, double (&v)[4]) {
    asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];"
                 : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[3])
                 : "l"(p));
}
__device__ __forceinline__ void st256(double* p, const double (&v)[4]) {
    asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(v[0]), "d"(v[1]), "d"(v[2]), "d"(v[3])
                 : "memory");
}
// 2^e as a double, e in [-1022, 1023]
__device__ __forceinline__ double pow2i(int e) { return __hiloint2double((e + 1023) << 20, 0); }

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
#endif

}  // namespace phb

struct phb_ctx : phb::Ctx {};
