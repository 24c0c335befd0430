// C ABI of the engine: context life cycle, input upload, schedule resolution, orchestration.
// Every function maps to a piece of TreeModel (see include/phylo_b200.h for the file:line map).
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <new>

#include "common.cuh"

namespace phb {

static thread_local std::string g_thread_error;
void set_thread_error(const std::string& msg) { g_thread_error = msg; }

static Tuning read_tuning() {
    {
        Tuning v;
        auto flag = [](const char* name) { return getenv(name) != nullptr; };
        auto num = [](const char* name) { const char* s = getenv(name); return s ? atoi(s) : 0; };
        v.disable_mma = flag("PHB_DISABLE_MMA");
        v.disable_tiptab = flag("PHB_DISABLE_TIPTAB");
        v.up_two_rows = flag("PHB_UP_TWO_ROWS");
        v.up_plain = flag("PHB_UP_PLAIN");
        v.deriv_no_st = flag("PHB_DERIV_NO_ST");
        v.deriv_matrix_form = flag("PHB_DERIV_MATRIX_FORM");
        v.compress_timing = flag("PHB_COMPRESS_TIMING");
        v.pair_ctas = num("PHB_PAIR_CTAS");
        v.pair_ppt = num("PHB_PAIR_PPT");
        v.pair_grid = num("PHB_PAIR_GRID");
        v.pair_one_warp_ctas = num("PHB_PAIR_ONE_WARP_CTAS") != 0;
        v.pair_cta_rounds = num("PHB_PAIR_CTA_ROUNDS");
        v.pair_stagger = num("PHB_PAIR_STAGGER");
        v.pair_full_p = flag("PHB_PAIR_FULL_P");
        v.up_ppt = num("PHB_UP_PPT");
        v.up_warps = num("PHB_UP_WARPS");
        v.tile_want = num("PHB_TILE_WANT");
        v.mma_variant = num("PHB_MMA_VARIANT");
        return v;
    }
}

static Tuning& tuning_slot() {
    static Tuning t = read_tuning();
    return t;
}

const Tuning& tuning() { return tuning_slot(); }
const char* thread_error() { return g_thread_error.c_str(); }

namespace {

constexpr size_t kAlign = 256;
inline size_t align_up(size_t v) { return (v + kAlign - 1) / kAlign * kAlign; }

struct Plan {
    size_t codes, lut, weights, clv, scale, up, up_scale, up_rows, root_clv, root_scale, pmats, dmats, model, lengths, rows, res_rows, scratch, scratch_size, tiptab, rmats, pimg, pimg_size, flags, edges, dmats_doubles,
        pattern_lnl, cat_lnl, partial, result, total;
    int root_block;
};

inline size_t code_pitch_for(int64_t S) { return ((size_t)S + 127) / 128 * 128; }

Plan make_plan(int n_tips, int64_t S, int K, int A, unsigned flags) {
    Plan p;
    size_t off = 0;
    auto take = [&](size_t bytes) {
        size_t at = off;
        off += align_up(bytes ? bytes : 1);
        return at;
    };
    const size_t n_int = n_tips > 2 ? (size_t)(n_tips - 2) : 0;
    const size_t max_rows = n_int > 0 ? n_int : 1;
    const bool store = !(flags & PHB_FLAG_NO_PARTIALS);
    const bool up = store && (flags & PHB_FLAG_UP_PARTIALS);
    const size_t node_doubles = (size_t)S * K * A;
    p.codes = take((size_t)n_tips * code_pitch_for(S));
    p.lut = take(256 * (size_t)A * 8);
    p.weights = take((size_t)S * 8);
    // down partials [n_int blocks] immediately followed by up partials [n_nodes blocks] (if requested)
    const size_t n_nodes = 2 * (size_t)n_tips - 2;
    const size_t n_blocks = n_int + (up ? n_nodes : 0) + 1;   // + the virtual-root block
    p.clv = take(store ? n_blocks * node_doubles * 8 : 0);
    p.scale = take(store ? n_blocks * (size_t)S * 4 : 0);
    p.up = p.clv + n_int * node_doubles * 8;
    p.up_scale = p.scale + n_int * (size_t)S * 4;
    p.up_rows = take(up ? 2 * max_rows * sizeof(OpRow) : 0);
    p.root_clv = p.clv + (n_blocks - 1) * node_doubles * 8;
    p.root_scale = p.scale + (n_blocks - 1) * (size_t)S * 4;
    p.root_block = (int)(n_blocks - 1);
    p.pmats = take((2 * max_rows + 2) * (size_t)K * A * A * 8);
    // derivative matrices for a whole launch of edges: every edge of the tree when that costs <= 16 MB, never fewer than 64
    {
        const size_t per_edge = 3 * (size_t)K * A * A, n_edges = 2 * (size_t)n_tips - 2;
        const size_t want = std::max<size_t>(kMaxEdgeBatch, std::min<size_t>(n_edges, ((size_t)16 << 20) / (per_edge * 8)));
        p.dmats_doubles = want * per_edge;
        p.dmats = take(p.dmats_doubles * 8);
        p.edges = take(n_edges * 16);
    }
    p.tiptab = take(A == 4 ? (2 * max_rows + 2) * (size_t)K * kTipTabCodes * 32
                           : ((A == 20 || A == 61) ? (2 * max_rows + 2) * (size_t)K * 64 * A * 8 : 0));
    // 4 states: packed symmetric P blocks for the lnL-only walk (+ one copy round of slack behind the last block)
    p.rmats = take(A == 4 ? (2 * max_rows + 2) * (size_t)K * 80 + 1024 : 0);
    // 61 states: padded staging images (64 rows x 68 doubles) of every P block and tip table for the DMMA kernels
    p.pimg_size = A == 61 ? (2 * max_rows + 2) * (size_t)K * 2 * 64 * 68 * 8 : 0;
    p.pimg = take(p.pimg_size);
    p.model = take((2 * (size_t)A * A + 2 * A + 2 * K) * 8);
    p.lengths = take((2 * max_rows + 2 + 2 * (size_t)n_tips) * 8);   // rows, root, + trial lengths of a derivative launch
    p.rows = take((max_rows + 1) * sizeof(OpRow));   // + the root pseudo-row
    p.res_rows = take((max_rows + 1) * 16);
    // parking area of the lnL-only resident kernel (4-state models): 160 SMs x 16 warps x 15 blocks
    p.scratch_size = A == 4 ? (size_t)160 * 16 * 15 * ((size_t)K * 1024 + 128) : 0;
    p.scratch = take(p.scratch_size);
    p.flags = take(2 * (kMaxFlagChunks + 1) * sizeof(int));
    p.pattern_lnl = take((size_t)S * 8);
    p.cat_lnl = take((size_t)S * K * 8);
    p.partial = take((size_t)kPartialCap * 8);
    p.result = take(std::max<size_t>((size_t)kMaxEdgeBatch * 4, 6 * (size_t)n_tips) * 8);   // 3 sums per edge
    p.total = off;
    return p;
}

bool shape_ok(int n_tips, int64_t S, int K, int A, std::string* why) {
    if (n_tips < 2) { *why = "need at least two tips"; return false; }
    if (S < 1) { *why = "need at least one pattern"; return false; }
    if (K < 1 || K > 16) { *why = "number of rate categories must be in 1..16"; return false; }
    if (A < 2 || A > 64) { *why = "number of states must be in 2..64"; return false; }
    return true;
}

}  // namespace

// operand-resident post-order pass with all blocks stored (clv_dna_pair.cu)
int resident_store(Ctx* c) { return dna_pair_store(c); }

int run_rows(Ctx* c, const RowSet& rs, int mode) {
    if (dna_supported(c)) return dna_run_rows(c, rs, mode);
    if (mma_supported(c) && !tuning().disable_mma) return mma_run_rows(c, rs, mode);
    return generic_run_rows(c, rs, mode);
}

namespace {

int activate(Ctx* c) {
    PHB_CUDA(c, cudaSetDevice(c->device));
    return PHB_OK;
}

// Turn the raw (PAR, CH1, CH2) rows into kernel rows; validates dependencies.
int resolve_schedule(Ctx* c) {
    const int n_rows = (int)c->rows_raw.size() / 3;
    PHB_REQUIRE(c, c->have_tips, PHB_ERR_STATE, "phb_set_tips must be called before the schedule can be resolved");
    c->node_slot.assign(c->n_nodes, -1);
    c->node_row.assign(c->n_nodes, -1);
    c->node_parent.assign(c->n_nodes, -1);
    c->rows.assign(n_rows, OpRow{});
    auto rank = [](int kind) { return kind == SRC_TIP ? 0 : (kind == SRC_PREV ? 1 : 2); };
    for (int r = 0; r < n_rows; ++r) {
        const int par = c->rows_raw[3 * r], ch[2] = {c->rows_raw[3 * r + 1], c->rows_raw[3 * r + 2]};
        PHB_REQUIRE(c, par >= 0 && par < c->n_nodes, PHB_ERR_INVALID, "schedule: parent node id out of range");
        PHB_REQUIRE(c, c->node_tip[par] < 0, PHB_ERR_INVALID, "schedule: a tip appears as a parent");
        PHB_REQUIRE(c, c->node_row[par] < 0, PHB_ERR_INVALID, "schedule: a node is computed twice");
        OpRow row{};
        row.dst = r;
        for (int i = 0; i < 2; ++i) {
            PHB_REQUIRE(c, ch[i] >= 0 && ch[i] < c->n_nodes && ch[i] != par, PHB_ERR_INVALID,
                        "schedule: child node id out of range");
            PHB_REQUIRE(c, c->node_parent[ch[i]] < 0, PHB_ERR_INVALID, "schedule: a node has two parents");
            c->node_parent[ch[i]] = par;
            if (c->node_tip[ch[i]] >= 0) {
                row.kind[i] = SRC_TIP;
                row.src[i] = c->node_tip[ch[i]];
            } else {
                PHB_REQUIRE(c, c->node_row[ch[i]] >= 0, PHB_ERR_INVALID,
                            "schedule: a child is used before the row that computes it");
                row.src[i] = c->node_slot[ch[i]];
                row.kind[i] = (c->node_row[ch[i]] == r - 1) ? SRC_PREV : SRC_GLOBAL;
            }
            row.pidx[i] = 2 * r + i;
        }
        PHB_REQUIRE(c, ch[0] != ch[1], PHB_ERR_INVALID, "schedule: both children are the same node");
        if (rank(row.kind[0]) > rank(row.kind[1])) {
            std::swap(row.kind[0], row.kind[1]);
            std::swap(row.src[0], row.src[1]);
            std::swap(row.pidx[0], row.pidx[1]);
        }
        c->rows[r] = row;
        c->node_slot[par] = r;
        c->node_row[par] = r;
    }
    if (!c->level_offsets.empty()) {
        const int n_levels = (int)c->level_offsets.size() - 1;
        PHB_REQUIRE(c, c->level_offsets.front() == 0 && c->level_offsets.back() == n_rows, PHB_ERR_INVALID,
                    "schedule: level offsets must start at 0 and end at n_rows");
        std::vector<int> level_of_row(n_rows, 0);
        for (int l = 0; l < n_levels; ++l) {
            PHB_REQUIRE(c, c->level_offsets[l] <= c->level_offsets[l + 1], PHB_ERR_INVALID,
                        "schedule: level offsets must be non-decreasing");
            for (int r = c->level_offsets[l]; r < c->level_offsets[l + 1]; ++r) level_of_row[r] = l;
        }
        for (int r = 0; r < n_rows; ++r)
            for (int i = 0; i < 2; ++i) {
                const int child = c->rows_raw[3 * r + 1 + i];
                if (c->node_tip[child] < 0)
                    PHB_REQUIRE(c, level_of_row[c->node_row[child]] < level_of_row[r], PHB_ERR_INVALID,
                                "schedule: a row depends on a row of the same or a later level");
            }
    }
    if (n_rows > 0)
        PHB_CUDA(c, cudaMemcpyAsync(c->d_rows, c->rows.data(), (size_t)n_rows * sizeof(OpRow),
                                    cudaMemcpyHostToDevice, c->stream));
    c->have_schedule = true;
    c->have_partials = false;
    c->have_up = false;
    c->sched_gen++;
    return PHB_OK;
}

int node_operand_ok(Ctx* c, int node) {
    PHB_REQUIRE(c, node >= 0 && node < c->n_nodes, PHB_ERR_INVALID, "node id out of range");
    if (c->node_tip[node] >= 0) return PHB_OK;
    PHB_REQUIRE(c, c->have_schedule && c->node_slot[node] >= 0, PHB_ERR_INVALID,
                "node is neither a tip nor computed by the schedule");
    return PHB_OK;
}

}  // namespace
}  // namespace phb

using namespace phb;

extern "C" {

int phb_version(void) { return PHB_VERSION; }

int phb_reload_tuning(void) {
    tuning_slot() = read_tuning();
    return PHB_OK;
}

const char* phb_status_name(int status) {
    switch (status) {
        case PHB_OK: return "PHB_OK";
        case PHB_ERR_INVALID: return "PHB_ERR_INVALID";
        case PHB_ERR_CUDA: return "PHB_ERR_CUDA";
        case PHB_ERR_NO_DEVICE: return "PHB_ERR_NO_DEVICE";
        case PHB_ERR_STATE: return "PHB_ERR_STATE";
        case PHB_ERR_NOMEM: return "PHB_ERR_NOMEM";
        case PHB_ERR_UNSUPPORTED: return "PHB_ERR_UNSUPPORTED";
    }
    return "PHB_ERR_UNKNOWN";
}

const char* phb_last_error(const phb_ctx* ctx) { return ctx ? ctx->err.c_str() : thread_error(); }

int64_t phb_launch_count(const phb_ctx* ctx) { return ctx ? ctx->launches : 0; }

size_t phb_workspace_bytes(int n_tips, int64_t n_patterns, int n_cat, int n_states, unsigned flags) {
    std::string why;
    if (!shape_ok(n_tips, n_patterns, n_cat, n_states, &why)) return 0;
    return make_plan(n_tips, n_patterns, n_cat, n_states, flags).total;
}

int phb_create(int device, int n_tips, int64_t n_patterns, int n_cat, int n_states, unsigned flags, void* workspace,
               size_t workspace_bytes, void* stream, phb_ctx** out) {
    if (out == nullptr) {
        set_thread_error("phb_create: out is NULL");
        return PHB_ERR_INVALID;
    }
    *out = nullptr;
    std::string why;
    if (!shape_ok(n_tips, n_patterns, n_cat, n_states, &why)) {
        set_thread_error("phb_create: " + why);
        return PHB_ERR_INVALID;
    }
    int n_dev = 0;
    cudaError_t e = cudaGetDeviceCount(&n_dev);
    if (e != cudaSuccess || n_dev == 0) {
        set_thread_error(std::string("phb_create: no CUDA device available (") + cudaGetErrorString(e) +
                         "); this engine has no CPU fallback");
        cudaGetLastError();
        return PHB_ERR_NO_DEVICE;
    }
    if (device < 0 || device >= n_dev) {
        set_thread_error("phb_create: device index out of range");
        return PHB_ERR_INVALID;
    }
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess || (e = cudaSetDevice(device)) != cudaSuccess) {
        set_thread_error(std::string("phb_create: ") + cudaGetErrorString(e));
        return PHB_ERR_CUDA;
    }
    if (prop.major < 10) {
        set_thread_error("phb_create: device is not sm_100 or newer (kernels are built for sm_100a only)");
        return PHB_ERR_NO_DEVICE;
    }
    phb_ctx* c = new (std::nothrow) phb_ctx();
    if (!c) {
        set_thread_error("phb_create: out of host memory");
        return PHB_ERR_NOMEM;
    }
    c->device = device;
    c->stream = (cudaStream_t)stream;
    c->n_tips = n_tips;
    c->S = n_patterns;
    c->K = n_cat;
    c->A = n_states;
    c->flags = flags;
    c->n_nodes = 2 * n_tips - 2;
    c->n_internal = n_tips > 2 ? n_tips - 2 : 0;
    c->sm_count = prop.multiProcessorCount;
    c->smem_optin = prop.sharedMemPerBlockOptin;
    c->smem_per_sm = prop.sharedMemPerMultiprocessor;
    const Plan p = make_plan(n_tips, n_patterns, n_cat, n_states, flags);
    if (workspace != nullptr) {
        if (workspace_bytes < p.total || ((uintptr_t)workspace % kAlign) != 0) {
            set_thread_error("phb_create: workspace too small or not 256-byte aligned");
            delete c;
            return PHB_ERR_NOMEM;
        }
        c->ws = (uint8_t*)workspace;
        c->owns_ws = false;
    } else {
        void* ptr = nullptr;
        if ((e = cudaMalloc(&ptr, p.total)) != cudaSuccess) {
            set_thread_error(std::string("phb_create: cudaMalloc of workspace failed: ") + cudaGetErrorString(e));
            cudaGetLastError();
            delete c;
            return PHB_ERR_NOMEM;
        }
        c->ws = (uint8_t*)ptr;
        c->owns_ws = true;
    }
    c->ws_bytes = p.total;
    uint8_t* w = c->ws;
    c->d_codes_ws = w + p.codes;
    c->d_codes = c->d_codes_ws;
    c->code_pitch = code_pitch_for(n_patterns);
    c->d_lut = (double*)(w + p.lut);
    c->d_weights = nullptr;  // all ones until phb_set_pattern_weights
    const bool store = !(flags & PHB_FLAG_NO_PARTIALS);
    c->d_clv = store ? (double*)(w + p.clv) : nullptr;
    c->d_scale = store ? (int32_t*)(w + p.scale) : nullptr;
    const bool up = store && (flags & PHB_FLAG_UP_PARTIALS);
    c->d_up = up ? (double*)(w + p.up) : nullptr;
    c->d_up_scale = up ? (int32_t*)(w + p.up_scale) : nullptr;
    c->d_up_rows = up ? (OpRow*)(w + p.up_rows) : nullptr;
    c->d_root_clv = store ? (double*)(w + p.root_clv) : nullptr;
    c->d_root_scale = store ? (int32_t*)(w + p.root_scale) : nullptr;
    c->root_block = p.root_block;
    c->d_pmats = (double*)(w + p.pmats);
    c->d_dmats = (double*)(w + p.dmats);
    c->dmats_doubles = p.dmats_doubles;
    c->d_edges = (void*)(w + p.edges);
    c->d_tiptab = (n_states == 4 || n_states == 20 || n_states == 61) ? (double*)(w + p.tiptab) : nullptr;
    c->d_rmats = n_states == 4 ? (double*)(w + p.rmats) : nullptr;
    c->d_pimg = p.pimg_size ? (double*)(w + p.pimg) : nullptr;
    c->pimg_rows = p.pimg_size ? 64 : 0;
    c->pimg_pitch = p.pimg_size ? 68 : 0;
    c->d_model = (double*)(w + p.model);
    c->d_lengths = (double*)(w + p.lengths);
    c->d_rows = (OpRow*)(w + p.rows);
    c->d_res_rows = (void*)(w + p.res_rows);
    c->d_scratch = p.scratch_size ? w + p.scratch : nullptr;
    c->scratch_bytes = p.scratch_size;
    c->d_flags = (int*)(w + p.flags);
    c->d_pattern_lnl = (double*)(w + p.pattern_lnl);
    c->d_cat_lnl = (double*)(w + p.cat_lnl);
    c->d_partial_sums = (double*)(w + p.partial);
    c->d_result = (double*)(w + p.result);
    c->result_doubles = std::max<size_t>((size_t)kMaxEdgeBatch * 4, 6 * (size_t)n_tips);
    c->node_tip.assign(c->n_nodes, -1);
    *out = c;
    return PHB_OK;
}

static void peer_close(phb_ctx* c) {
    for (int r = 0; r < kMaxPeers; ++r) {
        if (c->peer.cells[r] != nullptr && c->peer.cells[r] != c->peer.own) cudaIpcCloseMemHandle(c->peer.cells[r]);
        c->peer.cells[r] = nullptr;
    }
    c->peer.connected = false;
}

int phb_destroy(phb_ctx* c) {
    if (!c) return PHB_OK;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    if (c->copy_stream) {
        cudaStreamSynchronize(c->copy_stream);
        for (int i = 0; i < kMaxChunks; ++i) cudaEventDestroy(c->chunk_events[i]);
        cudaEventDestroy(c->start_event);
        cudaStreamDestroy(c->copy_stream);
    }
    if (c->copy_stream)
        for (int i = 0; i < 2; ++i) {
            if (c->slot_done[i]) cudaEventDestroy(c->slot_done[i]);
            if (c->result_event[i]) cudaEventDestroy(c->result_event[i]);
            if (c->copies_done[i]) cudaEventDestroy(c->copies_done[i]);
        }
    peer_close(c);
    if (c->peer.own) cudaFree(c->peer.own);
    if (c->h_epoch) cudaFreeHost(c->h_epoch);
    if (c->h_results) cudaFreeHost(c->h_results);
    if (c->owns_ws && c->ws) cudaFree(c->ws);
    delete c;
    return PHB_OK;
}

// Wait for the context's stream; if a host-fed evaluation (phb_lnl_from_host*) is in flight, also retire its copy
// stream and report a chunk of tip codes that never arrived.
static int finish_stream(phb_ctx* c) {
    int late = 0;
    if (c->pipelined_pending)
        PHB_CUDA(c, cudaMemcpyAsync(&late, c->d_flags + kMaxFlagChunks, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    PHB_CUDA(c, cudaStreamSynchronize(c->stream));
    if (c->pipelined_pending) {
        c->pipelined_pending = false;
        PHB_CUDA(c, cudaStreamSynchronize(c->copy_stream));
        if (late) {
            cudaMemsetAsync(c->d_flags + kMaxFlagChunks, 0, sizeof(int), c->stream);
            return c->fail(PHB_ERR_CUDA, "phb_lnl_from_host: a chunk of tip codes never arrived on the device");
        }
    }
    return PHB_OK;
}

int phb_sync(phb_ctx* c) {
    if (!c) return PHB_ERR_INVALID;
    int st = activate(c);
    if (st) return st;
    return finish_stream(c);
}

int phb_set_tips(phb_ctx* c, const uint8_t* codes, int codes_on_device, int n_codes, const double* lut,
                 const int32_t* tip_nodes) {
    if (!c) return PHB_ERR_INVALID;
    int st = activate(c);
    if (st) return st;
    PHB_REQUIRE(c, codes && lut && tip_nodes, PHB_ERR_INVALID, "phb_set_tips: NULL argument");
    PHB_REQUIRE(c, n_codes >= 1 && n_codes <= 256, PHB_ERR_INVALID, "phb_set_tips: n_codes must be in 1..256");
    std::vector<int32_t> node_tip(c->n_nodes, -1);
    for (int t = 0; t < c->n_tips; ++t) {
        const int node = tip_nodes[t];
        PHB_REQUIRE(c, node >= 0 && node < c->n_nodes, PHB_ERR_INVALID, "phb_set_tips: tip node id out of range");
        PHB_REQUIRE(c, node_tip[node] < 0, PHB_ERR_INVALID, "phb_set_tips: two tips share a node id");
        node_tip[node] = t;
    }
    for (int i = 0; i < n_codes * c->A; ++i)
        PHB_REQUIRE(c, lut[i] >= 0.0 && std::isfinite(lut[i]), PHB_ERR_INVALID,
                    "phb_set_tips: look-up table entries must be finite and non-negative");
    // host or device source, always copied into the pitched workspace buffer (padding stays zero)
    const size_t n_code_bytes = (size_t)c->n_tips * c->code_pitch;
    if (c->code_pitch != (size_t)c->S && (!c->have_tips || c->codes_packed))
        PHB_CUDA(c, cudaMemsetAsync(c->d_codes_ws, 0, n_code_bytes, c->stream));
    c->codes_packed = false;
    c->codes_mode = 0;
    PHB_CUDA(c, cudaMemcpy2DAsync(c->d_codes_ws, c->code_pitch, codes, (size_t)c->S, (size_t)c->S, (size_t)c->n_tips,
                                  codes_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, c->stream));
    c->d_codes = c->d_codes_ws;
    // The look-up table always has 256 rows on the device (unused rows are zero), so a stray code can
    // never index out of bounds; it is still an input error, detected on the device in one pass.
    {
        int worst = 0;
        int st2 = launch_max_code(c, c->d_codes, n_code_bytes, &worst);
        if (st2) return st2;
        PHB_REQUIRE(c, worst < n_codes, PHB_ERR_INVALID, "phb_set_tips: a code is >= n_codes");
    }
    std::vector<double> full(256 * (size_t)c->A, 0.0);
    std::memcpy(full.data(), lut, (size_t)n_codes * c->A * sizeof(double));
    PHB_CUDA(c, cudaMemcpyAsync(c->d_lut, full.data(), full.size() * sizeof(double), cudaMemcpyHostToDevice,
                                c->stream));
    PHB_CUDA(c, cudaStreamSynchronize(c->stream));  // `full` is a stack-lifetime staging buffer
    const bool remap = node_tip != c->node_tip;
    c->node_tip.swap(node_tip);
    c->n_codes = n_codes;
    c->have_tips = true;
    c->have_pmats = false;   // the P.lut tip tables depend on the look-up table and on n_codes: rebuild with the matrices
    c->have_partials = false;
    c->have_up = false;
    c->sched_gen++;   // the tip-table geometry the cached plans refer to may have changed
    if (remap && !c->rows_raw.empty()) return resolve_schedule(c);
    return PHB_OK;
}

int phb_set_pattern_weights(phb_ctx* c, const int64_t* weights) {
    if (!c) return PHB_ERR_INVALID;
    int st = activate(c);
    if (st) return st;
    if (weights == nullptr) {
        c->d_weights = nullptr;
        return PHB_OK;
    }
    std::vector<double> w((size_t)c->S);
    for (int64_t i = 0; i < c->S; ++i) {
        PHB_REQUIRE(c, weights[i] >= 0, PHB_ERR_INVALID, "phb_set_pattern_weights: negative weight");
        w[i] = (double)weights[i];
    }
    const Plan p = make_plan(c->n_tips, c->S, c->K, c->A, c->flags);
    c->d_weights = (double*)(c->ws + p.weights);
    PHB_CUDA(c, cudaMemcpyAsync(c->d_weights, w.data(), w.size() * sizeof(double), cudaMemcpyHostToDevice,
                                c->stream));
    PHB_CUDA(c, cudaStreamSynchronize(c->stream));
    return PHB_OK;
}

static int upload_mixture(phb_ctx* c, const double* freqs, const double* rates, const double* cat_weights) {
    const int A = c->A, K = c->K;
    std::vector<double> buf(2 * (size_t)A + 2 * K, 0.0);  // evals | freqs | rates | catw  (evals untouched here)
    for (int i = 0; i < A; ++i) {
        PHB_REQUIRE(c, std::isfinite(freqs[i]), PHB_ERR_INVALID, "model: non-finite state frequency");
        buf[i] = freqs[i];
    }
    for (int k = 0; k < K; ++k) {
        PHB_REQUIRE(c, rates[k] >= 0 && std::isfinite(rates[k]), PHB_ERR_INVALID, "model: category rates must be >= 0");
        PHB_REQUIRE(c, cat_weights[k] >= 0 && std::isfinite(cat_weights[k]), PHB_ERR_INVALID,
                    "model: category weights must be >= 0");
        buf[A + k] = rates[k];
        buf[A + K + k] = cat_weights[k];
    }
    c->h_freqs.assign(freqs, freqs + A);
    PHB_CUDA(c, cudaMemcpyAsync(c->model_freqs(), buf.data(), ((size_t)A + 2 * K) * sizeof(double),
                                cudaMemcpyHostToDevice, c->stream));
    PHB_CUDA(c, cudaStreamSynchronize(c->stream));
    c->have_mixture = true;
    return PHB_OK;
}

int phb_set_model(phb_ctx* c, const double* evecs, const double* evals, const double* ivecs, const double* freqs,
                  const double* rates, const double* cat_weights) {
    if (!c) return PHB_ERR_INVALID;
    int st = activate(c);
    if (st) return st;
    PHB_REQUIRE(c, evecs && evals && ivecs && freqs && rates && cat_weights, PHB_ERR_INVALID,
                "phb_set_model: NULL argument");
    const size_t AA = (size_t)c->A * c->A;
    std::vector<double> buf(2 * AA + c->A);
    std::memcpy(buf.data(), evecs, AA * 8);
    std::memcpy(buf.data() + AA, evals, (size_t)c->A * 8);
    std::memcpy(buf.data() + AA + c->A, ivecs, AA * 8);
    for (double v : buf) PHB_REQUIRE(c, std::isfinite(v), PHB_ERR_INVALID, "phb_set_model: non-finite eigen-system");
    PHB_CUDA(c, cudaMemcpyAsync(c->d_model, buf.data(), buf.size() * 8, cudaMemcpyHostToDevice, c->stream));
    PHB_CUDA(c, cudaStreamSynchronize(c->stream));
    st = upload_mixture(c, freqs, rates, cat_weights);
    if (st) return st;
    c->h_evecs.assign(evecs, evecs + AA);
    c->h_ivecs.assign(ivecs, ivecs + AA);
    {
        // detailed balance of Q = V diag(lambda) V^-1 with respect to the given frequencies: pi_i q_ij == pi_j q_ji.
        // Every model the reference's TreeModel can drive through an eigen-system has it; a caller of the C ABI need not.
        const int A = c->A;
        bool rev = true;
        double scale = 0.0;
        std::vector<double> q(AA, 0.0);
        for (int i = 0; i < A; ++i)
            for (int j = 0; j < A; ++j) {
                double v = 0.0;
                for (int m = 0; m < A; ++m) v += evecs[i * A + m] * evals[m] * ivecs[m * A + j];
                q[i * A + j] = v;
                scale = std::max(scale, std::fabs(v));
            }
        for (int i = 0; i < A && rev; ++i) {
            if (!(freqs[i] > 0)) rev = false;
            for (int j = i + 1; j < A && rev; ++j)
                if (std::fabs(freqs[i] * q[i * A + j] - freqs[j] * q[j * A + i]) > 1e-12 * scale) rev = false;
        }
        if (rev != c->reversible) c->res_cache.kind = 0;   // the cached walk descriptors point at the other kind of block
        c->reversible = rev;
    }
    c->have_model = true;
    c->have_pmats = false;
    c->have_partials = false;
    c->have_up = false;
    return PHB_OK;
}

int phb_set_mixture(phb_ctx* c, const double* freqs, const double* rates, const double* cat_weights) {
    if (!c) return PHB_ERR_INVALID;
    int st = activate(c);
    if (st) return st;
    PHB_REQUIRE(c, freqs && rates && cat_weights, PHB_ERR_INVALID, "phb_set_mixture: NULL argument");
    return upload_mixture(c, freqs, rates, cat_weights);
}

int phb_set_schedule(phb_ctx* c, int n_rows, const int32_t* rows, int n_levels, const int32_t* level_offsets) {
    if (!c) return PHB_ERR_INVALID;
    int st = activate(c);
    if (st) return st;
    PHB_REQUIRE(c, n_rows >= 0 && n_rows <= c->max_rows() && (n_rows == 0 || rows), PHB_ERR_INVALID,
                "phb_set_schedule: row count must be in 0..n_tips-2");
    PHB_REQUIRE(c, n_rows <= c->n_internal, PHB_ERR_INVALID, "phb_set_schedule: more rows than internal nodes");
    c->rows_raw.assign(rows, rows + 3 * (size_t)n_rows);
    c->level_offsets.clear();
    if (level_offsets != nullptr && n_levels > 0) c->level_offsets.assign(level_offsets, level_offsets + n_levels + 1);
    c->have_lengths = false;
    c->have_pmats = false;
    c->have_schedule = false;
    if (c->have_tips) return resolve_schedule(c);
    return PHB_OK;
}

int phb_set_edge_lengths(phb_ctx* c, const double* lengths) {
    if (!c) return PHB_ERR_INVALID;
    int st = activate(c);
    if (st) return st;
    const size_t n = c->rows_raw.size() / 3 * 2;
    PHB_REQUIRE(c, n == 0 || lengths, PHB_ERR_INVALID, "phb_set_edge_lengths: NULL lengths");
    for (size_t i = 0; i < n; ++i)
        PHB_REQUIRE(c, lengths[i] >= 0 && std::isfinite(lengths[i]), PHB_ERR_INVALID,
                    "phb_set_edge_lengths: branch lengths must be finite and >= 0");
    c->lengths.assign(lengths, lengths + n);
    if (n) {
        PHB_CUDA(c, cudaMemcpyAsync(c->d_lengths, c->lengths.data(), n * 8, cudaMemcpyHostToDevice, c->stream));
    }
    c->have_lengths = true;
    c->have_pmats = false;
    c->have_partials = false;
    c->have_up = false;
    return PHB_OK;
}

int phb_build_pmatrices(phb_ctx* c) {
    if (!c) return PHB_ERR_INVALID;
    int st = activate(c);
    if (st) return st;
    PHB_REQUIRE(c, c->have_model, PHB_ERR_STATE, "phb_build_pmatrices: no eigen-system set (phb_set_model)");
    PHB_REQUIRE(c, c->have_lengths, PHB_ERR_STATE, "phb_build_pmatrices: no edge lengths set");
    st = launch_build_pmatrices(c, c->d_lengths, (int)c->lengths.size(), c->d_pmats, 0, 0);
    if (st) return st;
    st = launch_tip_tables(c, 0, (int)c->lengths.size());
    if (st) return st;
    c->have_pmats = true;
    c->have_partials = false;
    c->have_up = false;
    return PHB_OK;
}

int phb_set_pmatrices(phb_ctx* c, const double* pmats) {
    if (!c) return PHB_ERR_INVALID;
    int st = activate(c);
    if (st) return st;
    const size_t n = c->rows_raw.size() / 3 * 2 * (size_t)c->K * c->A * c->A;
    PHB_REQUIRE(c, n == 0 || pmats, PHB_ERR_INVALID, "phb_set_pmatrices: NULL matrices");
    if (n) {
        PHB_CUDA(c, cudaMemcpyAsync(c->d_pmats, pmats, n * 8, cudaMemcpyHostToDevice, c->stream));
        PHB_CUDA(c, cudaStreamSynchronize(c->stream));
        st = launch_tip_tables(c, 0, (int)(c->rows_raw.size() / 3 * 2));
        if (st) return st;
    }
    c->have_pmats = true;
    c->have_partials = false;
    c->have_up = false;
    return PHB_OK;
}

int phb_get_pmatrix(phb_ctx* c, int row, int child, double* out) {
    if (!c) return PHB_ERR_INVALID;
    int st = activate(c);
    if (st) return st;
    PHB_REQUIRE(c, c->have_pmats, PHB_ERR_STATE, "phb_get_pmatrix: matrices have not been built");
    PHB_REQUIRE(c, out && row >= 0 && row < (int)c->rows_raw.size() / 3 && (child == 0 || child == 1), PHB_ERR_INVALID,
                "phb_get_pmatrix: bad row / child");
    const size_t blk = (size_t)c->K * c->A * c->A;
    PHB_CUDA(c, cudaMemcpyAsync(out, c->d_pmats + (size_t)(2 * row + child) * blk, blk * 8, cudaMemcpyDeviceToHost,
                                c->stream));
    PHB_CUDA(c, cudaStreamSynchronize(c->stream));
    return PHB_OK;
}

int phb_compute_partials(phb_ctx* c, int mode) {
    if (!c) return PHB_ERR_INVALID;
    int st = activate(c);
    if (st) return st;
    PHB_REQUIRE(c, !(c->flags & PHB_FLAG_NO_PARTIALS), PHB_ERR_STATE,
                "phb_compute_partials: context was created without partial storage");
    PHB_REQUIRE(c, c->have_tips, PHB_ERR_STATE, "phb_compute_partials: no tip data");
    PHB_REQUIRE(c, !c->codes_packed, PHB_ERR_STATE,
                "phb_compute_partials: the device holds packed codes (phb_lnl_from_host_packed); call phb_set_tips");
    PHB_REQUIRE(c, c->have_schedule, PHB_ERR_STATE, "phb_compute_partials: no schedule");
    PHB_REQUIRE(c, c->have_pmats, PHB_ERR_STATE, "phb_compute_partials: transition matrices not built");
    if (c->n_rows() == 0) {  // two-tip tree: nothing to prune
        c->have_partials = true;
        return PHB_OK;
    }
    if (mode == PHB_MODE_AUTO) {
        if (!c->level_offsets.empty()) {
            mode = PHB_MODE_LEVEL;
        } else if (dna_supported(c) && c->S >= 16384) {
            // operand-resident walk with streamed stores; needs a post-order schedule - if the caller's row order
            // is not one, fall back to the plain tile walk
            st = resident_store(c);
            if (st == PHB_OK) {
                c->have_partials = true;
                c->have_up = false;
                c->resident_partials = true;
                return PHB_OK;
            }
            if (st != PHB_ERR_UNSUPPORTED) return st;
            mode = PHB_MODE_TILE;
        } else {
            mode = PHB_MODE_TILE;
        }
    }
    PHB_REQUIRE(c, mode == PHB_MODE_TILE || mode == PHB_MODE_LEVEL || mode == PHB_MODE_RESIDENT, PHB_ERR_INVALID,
                "phb_compute_partials: bad mode");
    if (mode == PHB_MODE_RESIDENT) {
        PHB_REQUIRE(c, dna_supported(c), PHB_ERR_UNSUPPORTED,
                    "phb_compute_partials: resident mode covers 4-state models with K in {1,2,4,8}");
        st = resident_store(c);
        if (st) return st;
        c->have_partials = true;
        c->have_up = false;
        c->resident_partials = true;
        return PHB_OK;
    }
    PHB_REQUIRE(c, mode != PHB_MODE_LEVEL || !c->level_offsets.empty(), PHB_ERR_STATE,
                "phb_compute_partials: level mode needs level offsets in the schedule");
    const RowSet rs{c->d_rows, c->n_rows(), &c->level_offsets};
    st = run_rows(c, rs, mode);
    if (st) return st;
    c->have_partials = true;
    c->have_up = false;
    c->resident_partials = false;
    return PHB_OK;
}

static int prepare_root(phb_ctx* c, int node_a, int node_b, double length, const double* root_pmats) {
    int st = node_operand_ok(c, node_a);
    if (st) return st;
    st = node_operand_ok(c, node_b);
    if (st) return st;
    PHB_REQUIRE(c, node_a != node_b, PHB_ERR_INVALID, "root edge: both ends are the same node");
    PHB_REQUIRE(c, c->have_mixture, PHB_ERR_STATE, "root edge: frequencies / mixture not set");
    const size_t blk = (size_t)c->K * c->A * c->A;
    double* d_root_p = c->d_pmats + (size_t)(2 * c->max_rows()) * blk;
    if (root_pmats != nullptr) {
        PHB_CUDA(c, cudaMemcpyAsync(d_root_p, root_pmats, 2 * blk * 8, cudaMemcpyHostToDevice, c->stream));
        PHB_CUDA(c, cudaStreamSynchronize(c->stream));
    } else {
        PHB_REQUIRE(c, c->have_model, PHB_ERR_STATE, "root edge: no eigen-system set and no matrices given");
        PHB_REQUIRE(c, length >= 0 && std::isfinite(length), PHB_ERR_INVALID, "root edge: bad length");
        c->h_root_two[0] = 0.0;      // P(0) on a's side, P(length) on b's: tree_model.py:189-190
        c->h_root_two[1] = length;
        double* d_len = c->d_lengths + 2 * (size_t)c->max_rows();
        // pageable source: staged before the call returns - no synchronisation needed for a member array
        PHB_CUDA(c, cudaMemcpyAsync(d_len, c->h_root_two, sizeof c->h_root_two, cudaMemcpyHostToDevice, c->stream));
        st = launch_build_pmatrices(c, d_len, 2, d_root_p, 0, 0);
        if (st) return st;
    }
    st = launch_tip_tables(c, 2 * c->max_rows(), 2);
    if (st) return st;
    c->root_a = node_a;
    c->root_b = node_b;
    c->root_len = length;
    return PHB_OK;
}

// Every matrix an lnL-only evaluation needs - the rows' 2 n_rows and the root edge's two - in ONE pmatrix launch and ONE
// tip-table launch: the root lengths sit right behind the row lengths.  (At the reference's own test size, 10 taxa x
// 1000 patterns, an evaluation is launch-latency: two launches, a copy and a synchronisation fewer.)
static int build_eval_pmats(phb_ctx* c, int node_a, int node_b, double length) {
    const size_t n = c->lengths.size();
    if (n != 2 * (size_t)c->max_rows()) {   // two-tip tree, or a partial schedule: the two-step way
        int st = launch_build_pmatrices(c, c->d_lengths, (int)n, c->d_pmats, 0, 0);
        if (st) return st;
        st = launch_tip_tables(c, 0, (int)n);
        if (st) return st;
        c->have_pmats = true;
        return prepare_root(c, node_a, node_b, length, nullptr);
    }
    int st = node_operand_ok(c, node_a);
    if (st) return st;
    st = node_operand_ok(c, node_b);
    if (st) return st;
    PHB_REQUIRE(c, node_a != node_b, PHB_ERR_INVALID, "root edge: both ends are the same node");
    PHB_REQUIRE(c, c->have_mixture, PHB_ERR_STATE, "root edge: frequencies / mixture not set");
    PHB_REQUIRE(c, length >= 0 && std::isfinite(length), PHB_ERR_INVALID, "root edge: bad length");
    c->h_root_two[0] = 0.0;      // P(0) on a's side, P(length) on b's: tree_model.py:189-190
    c->h_root_two[1] = length;
    // pageable source: the copy is staged before the call returns, the member array only has to outlive the call
    PHB_CUDA(c, cudaMemcpyAsync(c->d_lengths + n, c->h_root_two, sizeof c->h_root_two, cudaMemcpyHostToDevice, c->stream));
    st = launch_build_pmatrices(c, c->d_lengths, (int)n + 2, c->d_pmats, 0, 0);
    if (st) return st;
    st = launch_tip_tables(c, 0, (int)n + 2);
    if (st) return st;
    c->have_pmats = true;
    c->root_a = node_a;
    c->root_b = node_b;
    c->root_len = length;
    return PHB_OK;
}

// enqueue only: root combine + mixture + log + weighted sum on the context's stream, total -> d_result[0]
static int root_lnl_enqueue(phb_ctx* c, int node_a, int node_b, double length, const double* root_pmats, bool want_cat) {
    int st = activate(c);
    if (st) return st;
    PHB_REQUIRE(c, !(c->flags & PHB_FLAG_NO_PARTIALS), PHB_ERR_STATE,
                "phb_root_lnl: context has no partial storage, use phb_lnl_resident");
    PHB_REQUIRE(c, c->have_tips, PHB_ERR_STATE, "phb_root_lnl: no tip data");
    PHB_REQUIRE(c, !c->codes_packed, PHB_ERR_STATE,
                "phb_root_lnl: the device holds packed codes (phb_lnl_from_host_packed); call phb_set_tips");
    PHB_REQUIRE(c, c->have_partials || c->n_rows() == 0, PHB_ERR_STATE,
                "phb_root_lnl: partials are stale, call phb_compute_partials first");
    st = prepare_root(c, node_a, node_b, length, root_pmats);
    if (st) return st;
    st = dna_supported(c) ? dna_root(c, node_a, node_b, want_cat, true)
                          : generic_root(c, node_a, node_b, want_cat, true);
    if (st) return st;
    c->have_root = true;
    return PHB_OK;
}

int phb_root_lnl(phb_ctx* c, int node_a, int node_b, double length, const double* root_pmats, double* total,
                 double* pattern_lnl, double* cat_lnl) {
    if (!c) return PHB_ERR_INVALID;
    PHB_REQUIRE(c, total != nullptr, PHB_ERR_INVALID, "phb_root_lnl: total is NULL");
    int st = root_lnl_enqueue(c, node_a, node_b, length, root_pmats, cat_lnl != nullptr);
    if (st) return st;
    PHB_CUDA(c, cudaMemcpyAsync(total, c->d_result, 8, cudaMemcpyDeviceToHost, c->stream));
    if (pattern_lnl)
        PHB_CUDA(c, cudaMemcpyAsync(pattern_lnl, c->d_pattern_lnl, (size_t)c->S * 8, cudaMemcpyDeviceToHost, c->stream));
    if (cat_lnl)
        PHB_CUDA(c, cudaMemcpyAsync(cat_lnl, c->d_cat_lnl, (size_t)c->S * c->K * 8, cudaMemcpyDeviceToHost, c->stream));
    PHB_CUDA(c, cudaStreamSynchronize(c->stream));
    return PHB_OK;
}

static int lnl_resident_enqueue(phb_ctx* c, int node_a, int node_b, double length) {
    int st = activate(c);
    if (st) return st;
    PHB_REQUIRE(c, c->have_tips && c->have_schedule && c->have_model && c->have_lengths, PHB_ERR_STATE,
                "phb_lnl_resident: tips, schedule, model and edge lengths must be set");
    PHB_REQUIRE(c, dna_supported(c), PHB_ERR_UNSUPPORTED, "phb_lnl_resident: only 4-state models with K in {1,2,4,8}");
    st = build_eval_pmats(c, node_a, node_b, length);
    if (st) return st;
    return dna_pair_lnl(c, node_a, node_b);
}

int phb_lnl_resident(phb_ctx* c, int node_a, int node_b, double length, double* total, double* pattern_lnl) {
    if (!c) return PHB_ERR_INVALID;
    PHB_REQUIRE(c, total != nullptr, PHB_ERR_INVALID, "phb_lnl_resident: total is NULL");
    int st = lnl_resident_enqueue(c, node_a, node_b, length);
    if (st) return st;
    PHB_CUDA(c, cudaMemcpyAsync(total, c->d_result, 8, cudaMemcpyDeviceToHost, c->stream));
    if (pattern_lnl)
        PHB_CUDA(c, cudaMemcpyAsync(pattern_lnl, c->d_pattern_lnl, (size_t)c->S * 8, cudaMemcpyDeviceToHost, c->stream));
    PHB_CUDA(c, cudaStreamSynchronize(c->stream));
    return PHB_OK;
}

// mode: 0 one byte per code, 1 two 4-bit codes per byte, 2 split 3-bit planes (codes = low plane, codes_hi = high plane)
static int lnl_from_host(phb_ctx* c, const uint8_t* codes, const uint8_t* codes_hi, int mode, int n_chunks, int node_a,
                         int node_b, double length, double* total, double* pattern_lnl) {
    const bool packed = mode == 1;
    if (!c) return PHB_ERR_INVALID;
    int st = activate(c);
    if (st) return st;
    PHB_REQUIRE(c, codes != nullptr, PHB_ERR_INVALID, "phb_lnl_from_host: NULL argument");
    PHB_REQUIRE(c, c->have_tips && c->have_schedule && c->have_model && c->have_lengths, PHB_ERR_STATE,
                "phb_lnl_from_host: tip layout (phb_set_tips), schedule, model and edge lengths must be set");
    PHB_REQUIRE(c, dna_supported(c), PHB_ERR_UNSUPPORTED, "phb_lnl_from_host: only 4-state models with K in {1,2,4,8}");
    PHB_REQUIRE(c, !packed || c->n_codes <= 16, PHB_ERR_UNSUPPORTED, "phb_lnl_from_host_packed: more than 16 codes");
    PHB_REQUIRE(c, mode != 2 || (codes_hi != nullptr && c->n_codes <= 8), PHB_ERR_UNSUPPORTED,
                "phb_lnl_from_host_split: needs both planes and a look-up table of at most 8 rows");
    st = build_eval_pmats(c, node_a, node_b, length);
    if (st) return st;
    c->have_partials = false;
    c->have_up = false;
    if (n_chunks <= 0) {
        // about 4 MB per chunk, between 4 and 64 chunks: at 1000 x 1M (0.5 GB of nibbles) 64 chunks beat 16 (69.9 vs 68.6
        // evaluations/s); at an eighth of that per GPU with eight GPUs on one host, 64 chunks are 192 small copies per
        // step and process, and their issue cost - not the bytes - bounds the evaluation
        const size_t per_code_x8 = mode == 0 ? 8 : (mode == 1 ? 4 : 3);
        const size_t bytes = (size_t)c->n_tips * (size_t)c->S * per_code_x8 / 8;
        n_chunks = (int)std::min<size_t>(64, std::max<size_t>(4, bytes >> 22));
    }
    st = dna_pair_from_host(c, codes, codes_hi, mode, n_chunks, node_a, node_b);
    if (st) return st;
    if (total == nullptr) return PHB_OK;   // stream-ordered form: phb_result_fetch / phb_sync complete the evaluation
    PHB_CUDA(c, cudaMemcpyAsync(total, c->d_result, 8, cudaMemcpyDeviceToHost, c->stream));
    if (pattern_lnl)
        PHB_CUDA(c, cudaMemcpyAsync(pattern_lnl, c->d_pattern_lnl, (size_t)c->S * 8, cudaMemcpyDeviceToHost, c->stream));
    return finish_stream(c);
}

int phb_lnl_from_host(phb_ctx* c, const uint8_t* codes, int n_chunks, int node_a, int node_b, double length,
                      double* total, double* pattern_lnl) {
    return lnl_from_host(c, codes, nullptr, 0, n_chunks, node_a, node_b, length, total, pattern_lnl);
}

int phb_lnl_from_host_packed(phb_ctx* c, const uint8_t* packed_codes, int n_chunks, int node_a, int node_b,
                             double length, double* total, double* pattern_lnl) {
    return lnl_from_host(c, packed_codes, nullptr, 1, n_chunks, node_a, node_b, length, total, pattern_lnl);
}

int phb_lnl_from_host_split(phb_ctx* c, const uint8_t* low_plane, const uint8_t* high_plane, int n_chunks, int node_a,
                            int node_b, double length, double* total, double* pattern_lnl) {
    if (c != nullptr && total == nullptr) return c->fail(PHB_ERR_INVALID, "phb_lnl_from_host_split: total is NULL");
    return lnl_from_host(c, low_plane, high_plane, 2, n_chunks, node_a, node_b, length, total, pattern_lnl);
}

int phb_lnl_from_host_split_async(phb_ctx* c, const uint8_t* low_plane, const uint8_t* high_plane, int n_chunks, int node_a,
                                  int node_b, double length) {
    if (!c) return PHB_ERR_INVALID;
    c->peer.use_now = c->peer.armed;   // (the guard type is defined further down; lnl_from_host ends with the reduction)
    c->peer.armed = false;
    const int st_ = lnl_from_host(c, low_plane, high_plane, 2, n_chunks, node_a, node_b, length, nullptr, nullptr);
    c->peer.use_now = false;
    return st_;
}
int phb_split_codes(const uint8_t* codes, int n_tips, int64_t n_patterns, uint8_t* low_plane, uint8_t* high_plane) {
    if (codes == nullptr || low_plane == nullptr || high_plane == nullptr || n_tips < 0 || n_patterns < 0) {
        set_thread_error("phb_split_codes: bad argument");
        return PHB_ERR_INVALID;
    }
    const int64_t row_lo = (n_patterns + 3) / 4, row_hi = (n_patterns + 7) / 8;
    for (int t = 0; t < n_tips; ++t) {
        const uint8_t* src = codes + (size_t)t * n_patterns;
        uint8_t* lo = low_plane + (size_t)t * row_lo;
        uint8_t* hi = high_plane + (size_t)t * row_hi;
        std::memset(lo, 0, (size_t)row_lo);
        std::memset(hi, 0, (size_t)row_hi);
        for (int64_t s = 0; s < n_patterns; ++s) {
            const unsigned v = src[s];
            if (v > 7u) {
                set_thread_error("phb_split_codes: a code does not fit in 3 bits");
                return PHB_ERR_INVALID;
            }
            lo[s >> 2] |= (uint8_t)((v & 3u) << (2 * (s & 3)));
            hi[s >> 3] |= (uint8_t)((v >> 2) << (s & 7));
        }
    }
    return PHB_OK;
}

int phb_pack_codes(const uint8_t* codes, int n_tips, int64_t n_patterns, uint8_t* out) {
    if (codes == nullptr || out == nullptr || n_tips < 0 || n_patterns < 0) {
        set_thread_error("phb_pack_codes: bad argument");
        return PHB_ERR_INVALID;
    }
    const int64_t row = (n_patterns + 1) / 2;
    for (int t = 0; t < n_tips; ++t) {
        const uint8_t* src = codes + (size_t)t * n_patterns;
        uint8_t* dst = out + (size_t)t * row;
        for (int64_t j = 0; j < row; ++j) {
            const unsigned lo = src[2 * j], hi = 2 * j + 1 < n_patterns ? src[2 * j + 1] : 0u;
            if (lo > 15u || hi > 15u) {
                set_thread_error("phb_pack_codes: a code does not fit in 4 bits");
                return PHB_ERR_INVALID;
            }
            dst[j] = (uint8_t)(lo | (hi << 4));
        }
    }
    return PHB_OK;
}

int phb_get_partials(phb_ctx* c, int node, double* out) {
    if (!c) return PHB_ERR_INVALID;
    int st = activate(c);
    if (st) return st;
    PHB_REQUIRE(c, out != nullptr, PHB_ERR_INVALID, "phb_get_partials: out is NULL");
    PHB_REQUIRE(c, node >= 0 && node < c->n_nodes, PHB_ERR_INVALID, "phb_get_partials: node id out of range");
    const size_t S = (size_t)c->S, K = c->K, A = c->A;
    if (c->node_tip[node] >= 0) {
        PHB_REQUIRE(c, c->have_tips && !c->codes_packed, PHB_ERR_STATE, "phb_get_partials: no (unpacked) tip data");
        std::vector<uint8_t> codes(S);
        std::vector<double> lut(256 * A);
        PHB_CUDA(c, cudaMemcpyAsync(codes.data(), c->d_codes + (size_t)c->node_tip[node] * c->code_pitch, S, cudaMemcpyDeviceToHost,
                                    c->stream));
        PHB_CUDA(c, cudaMemcpyAsync(lut.data(), c->d_lut, lut.size() * 8, cudaMemcpyDeviceToHost, c->stream));
        PHB_CUDA(c, cudaStreamSynchronize(c->stream));
        for (size_t s = 0; s < S; ++s)
            for (size_t k = 0; k < K; ++k) std::memcpy(out + (s * K + k) * A, lut.data() + (size_t)codes[s] * A, A * 8);
        return PHB_OK;
    }
    PHB_REQUIRE(c, c->have_partials && c->node_slot[node] >= 0, PHB_ERR_STATE,
                "phb_get_partials: partials of this node have not been computed");
    PHB_CUDA(c, cudaMemcpyAsync(out, c->d_clv + (size_t)c->node_slot[node] * c->clv_stride(), c->clv_stride() * 8,
                                cudaMemcpyDeviceToHost, c->stream));
    PHB_CUDA(c, cudaStreamSynchronize(c->stream));
    return PHB_OK;
}

static int scalers_to_host(phb_ctx* c, const int32_t* d_exp, double* out) {
    const size_t S = (size_t)c->S, K = c->K;
    std::vector<int32_t> e(S);
    PHB_CUDA(c, cudaMemcpyAsync(e.data(), d_exp, S * 4, cudaMemcpyDeviceToHost, c->stream));
    PHB_CUDA(c, cudaStreamSynchronize(c->stream));
    for (size_t s = 0; s < S; ++s)
        for (size_t k = 0; k < K; ++k) out[s * K + k] = (double)e[s] * kLn2;
    return PHB_OK;
}

int phb_get_scalers(phb_ctx* c, int node, double* out) {
    if (!c) return PHB_ERR_INVALID;
    int st = activate(c);
    if (st) return st;
    PHB_REQUIRE(c, out != nullptr, PHB_ERR_INVALID, "phb_get_scalers: out is NULL");
    PHB_REQUIRE(c, node >= 0 && node < c->n_nodes, PHB_ERR_INVALID, "phb_get_scalers: node id out of range");
    if (c->node_tip[node] >= 0) {
        std::memset(out, 0, (size_t)c->S * c->K * 8);
        return PHB_OK;
    }
    PHB_REQUIRE(c, c->have_partials && c->node_slot[node] >= 0, PHB_ERR_STATE,
                "phb_get_scalers: partials of this node have not been computed");
    return scalers_to_host(c, c->d_scale + (size_t)c->node_slot[node] * c->S, out);
}

int phb_get_root_partials(phb_ctx* c, double* out_partials, double* out_scalers) {
    if (!c) return PHB_ERR_INVALID;
    int st = activate(c);
    if (st) return st;
    PHB_REQUIRE(c, c->have_root, PHB_ERR_STATE, "phb_get_root_partials: phb_root_lnl has not run");
    if (out_partials) {
        PHB_CUDA(c, cudaMemcpyAsync(out_partials, c->d_root_clv, c->clv_stride() * 8, cudaMemcpyDeviceToHost, c->stream));
        PHB_CUDA(c, cudaStreamSynchronize(c->stream));
    }
    if (out_scalers) return scalers_to_host(c, c->d_root_scale, out_scalers);
    return PHB_OK;
}

int phb_compute_up_partials(phb_ctx* c, int node_a, int node_b, double length) {
    if (!c) return PHB_ERR_INVALID;
    int st = activate(c);
    if (st) return st;
    PHB_REQUIRE(c, c->d_up != nullptr, PHB_ERR_STATE, "phb_compute_up_partials: context lacks PHB_FLAG_UP_PARTIALS");
    PHB_REQUIRE(c, c->have_partials, PHB_ERR_STATE, "phb_compute_up_partials: run phb_compute_partials first");
    PHB_REQUIRE(c, c->have_model, PHB_ERR_STATE, "phb_compute_up_partials: needs the eigen-system (reversible model)");
    PHB_REQUIRE(c, c->have_pmats, PHB_ERR_STATE, "phb_compute_up_partials: transition matrices not built");
    st = prepare_root(c, node_a, node_b, length, nullptr);   // builds P(length) for the root edge
    if (st) return st;
    st = launch_up_partials(c, node_a, node_b);
    if (st) return st;
    c->have_up = true;
    return PHB_OK;
}

int phb_edge_derivatives(phb_ctx* c, int n_edges, const int32_t* nodes, const double* lengths, int chain_rule,
                         double* out) {
    if (!c) return PHB_ERR_INVALID;
    int st = activate(c);
    if (st) return st;
    PHB_REQUIRE(c, n_edges >= 0 && (n_edges == 0 || (nodes && lengths && out)), PHB_ERR_INVALID,
                "phb_edge_derivatives: NULL argument");
    PHB_REQUIRE(c, c->have_up, PHB_ERR_STATE, "phb_edge_derivatives: run phb_compute_up_partials first");
    return launch_edge_derivatives(c, n_edges, nodes, lengths, chain_rule, out);
}

// ---- re-rooting in place (Traversal.optimising_traversal, utils.py:137-188) --------------------------------------
int phb_update_node(phb_ctx* c, int node, int child_a, double len_a, int child_b, double len_b) {
    if (!c) return PHB_ERR_INVALID;
    int st = activate(c);
    if (st) return st;
    PHB_REQUIRE(c, !(c->flags & PHB_FLAG_NO_PARTIALS), PHB_ERR_STATE, "phb_update_node: context has no partial storage");
    PHB_REQUIRE(c, c->have_tips && !c->codes_packed && c->have_schedule && c->have_model, PHB_ERR_STATE,
                "phb_update_node: tips, schedule and eigen-system must be set");
    PHB_REQUIRE(c, c->have_partials, PHB_ERR_STATE, "phb_update_node: run phb_compute_partials first");
    PHB_REQUIRE(c, node >= 0 && node < c->n_nodes && c->node_tip[node] < 0 && c->node_slot[node] >= 0, PHB_ERR_INVALID,
                "phb_update_node: node must be an internal node of the schedule");
    const int kids[2] = {child_a, child_b};
    const double lens[2] = {len_a, len_b};
    OpRow row{};
    row.dst = c->node_slot[node];
    for (int i = 0; i < 2; ++i) {
        st = node_operand_ok(c, kids[i]);
        if (st) return st;
        PHB_REQUIRE(c, kids[i] != node && lens[i] >= 0 && std::isfinite(lens[i]), PHB_ERR_INVALID,
                    "phb_update_node: bad child or branch length");
        row.kind[i] = c->node_tip[kids[i]] >= 0 ? SRC_TIP : SRC_GLOBAL;
        row.src[i] = c->node_tip[kids[i]] >= 0 ? c->node_tip[kids[i]] : c->node_slot[kids[i]];
        row.pidx[i] = 2 * c->max_rows() + i;
    }
    PHB_REQUIRE(c, child_a != child_b, PHB_ERR_INVALID, "phb_update_node: both children are the same node");
    if (row.kind[0] != SRC_TIP && row.kind[1] == SRC_TIP) {   // canonical order: tips first
        std::swap(row.kind[0], row.kind[1]);
        std::swap(row.src[0], row.src[1]);
        std::swap(row.pidx[0], row.pidx[1]);
    }
    // the two matrices go where the root edge's would (the two spare blocks behind the rows' matrices)
    c->h_root_two[0] = len_a;
    c->h_root_two[1] = len_b;
    const size_t blk = (size_t)c->K * c->A * c->A;
    double* d_len = c->d_lengths + 2 * (size_t)c->max_rows();
    PHB_CUDA(c, cudaMemcpyAsync(d_len, c->h_root_two, sizeof c->h_root_two, cudaMemcpyHostToDevice, c->stream));
    st = launch_build_pmatrices(c, d_len, 2, c->d_pmats + (size_t)(2 * c->max_rows()) * blk, 0, 0);
    if (st) return st;
    st = launch_tip_tables(c, 2 * c->max_rows(), 2);
    if (st) return st;
    c->h_spare_row = row;   // member: the source of an asynchronous copy must outlive the call
    OpRow* d_row = c->d_rows + c->max_rows();
    PHB_CUDA(c, cudaMemcpyAsync(d_row, &c->h_spare_row, sizeof(OpRow), cudaMemcpyHostToDevice, c->stream));
    static const std::vector<int32_t> one_level = {0, 1};
    const RowSet rs{d_row, 1, &one_level};
    st = run_rows(c, rs, PHB_MODE_LEVEL);
    if (st) return st;
    c->have_up = false;        // whatever the pre-order pass left refers to the old rooting
    c->have_root = false;
    c->resident_partials = false;
    return PHB_OK;
}

int phb_branch_derivatives(phb_ctx* c, int node_a, int node_b, int n_lengths, const double* lengths, int chain_rule,
                           double* out) {
    if (!c) return PHB_ERR_INVALID;
    int st = activate(c);
    if (st) return st;
    PHB_REQUIRE(c, n_lengths >= 0 && (n_lengths == 0 || (lengths && out)), PHB_ERR_INVALID,
                "phb_branch_derivatives: NULL argument");
    PHB_REQUIRE(c, !(c->flags & PHB_FLAG_NO_PARTIALS) && c->have_tips && !c->codes_packed && c->have_model && c->have_mixture,
                PHB_ERR_STATE, "phb_branch_derivatives: needs stored partials, tips and a reversible model");
    PHB_REQUIRE(c, c->have_partials || c->n_rows() == 0, PHB_ERR_STATE, "phb_branch_derivatives: run phb_compute_partials first");
    st = node_operand_ok(c, node_a);
    if (st) return st;
    st = node_operand_ok(c, node_b);
    if (st) return st;
    PHB_REQUIRE(c, node_a != node_b, PHB_ERR_INVALID, "phb_branch_derivatives: both ends are the same node");
    std::vector<int32_t> near((size_t)n_lengths, node_a), far((size_t)n_lengths, node_b);
    return launch_edge_derivatives(c, n_lengths, near.data(), lengths, chain_rule, out, far.data());
}

// ---- stream-ordered forms ------------------------------------------------------------------------------------
// phb_peer_sum_next arms exactly the next stream-ordered scalar-lnL entry point, whether it succeeds or not
namespace {
struct PeerUse {
    phb_ctx* c;
    explicit PeerUse(phb_ctx* ctx) : c(ctx) {
        c->peer.use_now = c->peer.armed;
        c->peer.armed = false;
    }
    ~PeerUse() { c->peer.use_now = false; }
};
}  // namespace

int phb_lnl_resident_async(phb_ctx* c, int node_a, int node_b, double length) {
    if (!c) return PHB_ERR_INVALID;
    PeerUse peer(c);
    return lnl_resident_enqueue(c, node_a, node_b, length);
}

int phb_root_lnl_async(phb_ctx* c, int node_a, int node_b, double length) {
    if (!c) return PHB_ERR_INVALID;
    PeerUse peer(c);
    return root_lnl_enqueue(c, node_a, node_b, length, nullptr, false);
}

int phb_lnl_from_host_packed_async(phb_ctx* c, const uint8_t* packed_codes, int n_chunks, int node_a, int node_b,
                                   double length) {
    if (!c) return PHB_ERR_INVALID;
    PeerUse peer(c);
    return lnl_from_host(c, packed_codes, nullptr, 1, n_chunks, node_a, node_b, length, nullptr, nullptr);
}

// ---- sum over the ranks of one box inside the reduction kernel (no collective library call) --------------------------
int phb_peer_buffer(phb_ctx* c, void* handle_out) {
    if (!c) return PHB_ERR_INVALID;
    int st = activate(c);
    if (st) return st;
    PHB_REQUIRE(c, handle_out != nullptr, PHB_ERR_INVALID, "phb_peer_buffer: NULL handle");
    static_assert(sizeof(cudaIpcMemHandle_t) == PHB_PEER_HANDLE_BYTES, "handle size is part of the ABI");
    if (c->peer.own == nullptr) {
        const size_t bytes = 2 * kMaxPeers * 16;
        PHB_CUDA(c, cudaMalloc(&c->peer.own, bytes));
        PHB_CUDA(c, cudaMemset(c->peer.own, 0, bytes));
        PHB_CUDA(c, cudaDeviceSynchronize());
    }
    cudaIpcMemHandle_t h;
    PHB_CUDA(c, cudaIpcGetMemHandle(&h, c->peer.own));
    std::memcpy(handle_out, &h, sizeof h);
    return PHB_OK;
}

int phb_peer_connect(phb_ctx* c, int rank, int world, const void* handles) {
    if (!c) return PHB_ERR_INVALID;
    int st = activate(c);
    if (st) return st;
    PHB_REQUIRE(c, handles != nullptr && world >= 1 && world <= kMaxPeers && rank >= 0 && rank < world, PHB_ERR_INVALID,
                "phb_peer_connect: bad rank / world (at most 16 ranks) or NULL handles");
    PHB_REQUIRE(c, c->peer.own != nullptr, PHB_ERR_STATE, "phb_peer_connect: call phb_peer_buffer first");
    PHB_CUDA(c, cudaStreamSynchronize(c->stream));
    peer_close(c);
    for (int r = 0; r < world; ++r) {
        if (r == rank) {
            c->peer.cells[r] = c->peer.own;
            continue;
        }
        cudaIpcMemHandle_t h;
        std::memcpy(&h, static_cast<const unsigned char*>(handles) + (size_t)r * sizeof h, sizeof h);
        void* ptr = nullptr;
        const cudaError_t e = cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            cudaGetLastError();
            peer_close(c);
            return c->fail(PHB_ERR_CUDA, std::string("phb_peer_connect: cannot map the buffer of rank ") + std::to_string(r) + " (" +
                                             cudaGetErrorString(e) + ")");
        }
        c->peer.cells[r] = ptr;
    }
    c->peer.rank = rank;
    c->peer.world = world;
    c->peer.connected = true;
    return PHB_OK;
}

int phb_peer_sum_next(phb_ctx* c) {
    if (!c) return PHB_ERR_INVALID;
    PHB_REQUIRE(c, c->peer.connected, PHB_ERR_STATE, "phb_peer_sum_next: no peers connected (phb_peer_connect)");
    c->peer.armed = true;
    return PHB_OK;
}

// ---- pipelined host-fed evaluations: two in flight, the copy of one under the walk of the other ------------------
int phb_lnl_from_host_submit(phb_ctx* c, const uint8_t* codes, const uint8_t* high_plane, int n_chunks, int node_a,
                             int node_b, double length, int* slot_out) {
    if (!c) return PHB_ERR_INVALID;
    PeerUse peer(c);
    int st = activate(c);
    if (st) return st;
    PHB_REQUIRE(c, codes != nullptr && slot_out != nullptr, PHB_ERR_INVALID, "phb_lnl_from_host_submit: NULL argument");
    PHB_REQUIRE(c, c->have_tips && c->have_schedule && c->have_model && c->have_lengths, PHB_ERR_STATE,
                "phb_lnl_from_host_submit: tip layout (phb_set_tips), schedule, model and edge lengths must be set");
    PHB_REQUIRE(c, dna_supported(c), PHB_ERR_UNSUPPORTED, "phb_lnl_from_host_submit: only 4-state models with K in {1,2,4,8}");
    const int mode = high_plane != nullptr ? 2 : 1;
    PHB_REQUIRE(c, c->n_codes <= (mode == 2 ? 8 : 16), PHB_ERR_UNSUPPORTED,
                "phb_lnl_from_host_submit: the look-up table has too many rows for this code format");
    st = build_eval_pmats(c, node_a, node_b, length);
    if (st) return st;
    c->have_partials = false;
    c->have_up = false;
    if (n_chunks <= 0) {
        const size_t bytes = (size_t)c->n_tips * (size_t)c->S * (mode == 1 ? 4 : 3) / 8;
        n_chunks = (int)std::min<size_t>(64, std::max<size_t>(4, bytes >> 22));
    }
    const int slot = c->next_slot;
    st = dna_pair_from_host(c, codes, high_plane, mode, n_chunks, node_a, node_b, slot);
    if (st) return st;
    c->next_slot = 1 - slot;
    *slot_out = slot;
    return PHB_OK;
}

int phb_result_post(phb_ctx* c, int slot) {
    if (!c) return PHB_ERR_INVALID;
    int st = activate(c);
    if (st) return st;
    PHB_REQUIRE(c, (slot == 0 || slot == 1) && c->h_results != nullptr, PHB_ERR_INVALID, "phb_result_post: no such evaluation in flight");
    PHB_CUDA(c, cudaMemcpyAsync(c->h_results + slot, c->d_result + slot, 8, cudaMemcpyDeviceToHost, c->stream));
    PHB_CUDA(c, cudaMemcpyAsync(reinterpret_cast<int*>(c->h_results + 2) + slot, c->d_flags + slot * (kMaxFlagChunks + 1) + kMaxFlagChunks,
                                sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    PHB_CUDA(c, cudaEventRecord(c->result_event[slot], c->stream));
    return PHB_OK;
}

int phb_result_wait(phb_ctx* c, int slot, double* out) {
    if (!c) return PHB_ERR_INVALID;
    int st = activate(c);
    if (st) return st;
    PHB_REQUIRE(c, (slot == 0 || slot == 1) && out != nullptr && c->h_results != nullptr, PHB_ERR_INVALID,
                "phb_result_wait: no such evaluation in flight");
    PHB_CUDA(c, cudaEventSynchronize(c->result_event[slot]));
    int* const late = reinterpret_cast<int*>(c->h_results + 2) + slot;
    if (*late) {
        *late = 0;
        cudaMemsetAsync(c->d_flags + slot * (kMaxFlagChunks + 1) + kMaxFlagChunks, 0, sizeof(int), c->stream);
        return c->fail(PHB_ERR_CUDA, "phb_result_wait: a chunk of tip codes never arrived on the device");
    }
    *out = c->h_results[slot];
    return PHB_OK;
}

int phb_edge_derivatives_async(phb_ctx* c, int n_edges, const int32_t* nodes, const double* lengths, int chain_rule) {
    if (!c) return PHB_ERR_INVALID;
    int st = activate(c);
    if (st) return st;
    PHB_REQUIRE(c, n_edges >= 0 && (n_edges == 0 || (nodes && lengths)), PHB_ERR_INVALID,
                "phb_edge_derivatives_async: NULL argument");
    PHB_REQUIRE(c, 3 * (size_t)n_edges <= c->result_doubles, PHB_ERR_INVALID,
                "phb_edge_derivatives_async: more edges than the device result buffer holds (3 doubles per edge)");
    PHB_REQUIRE(c, c->have_up, PHB_ERR_STATE, "phb_edge_derivatives_async: run phb_compute_up_partials first");
    return launch_edge_derivatives(c, n_edges, nodes, lengths, chain_rule, nullptr);
}

int phb_device_result(phb_ctx* c, void** device_ptr, int64_t* capacity_doubles) {
    if (!c) return PHB_ERR_INVALID;
    PHB_REQUIRE(c, device_ptr != nullptr, PHB_ERR_INVALID, "phb_device_result: NULL argument");
    *device_ptr = c->d_result;
    if (capacity_doubles) *capacity_doubles = (int64_t)c->result_doubles;
    return PHB_OK;
}

int phb_result_fetch(phb_ctx* c, int n, double* out) {
    if (!c) return PHB_ERR_INVALID;
    int st = activate(c);
    if (st) return st;
    PHB_REQUIRE(c, n >= 0 && (size_t)n <= c->result_doubles && (n == 0 || out), PHB_ERR_INVALID,
                "phb_result_fetch: bad count or NULL output");
    if (n) PHB_CUDA(c, cudaMemcpyAsync(out, c->d_result, (size_t)n * 8, cudaMemcpyDeviceToHost, c->stream));
    return finish_stream(c);
}

}  // extern "C"
