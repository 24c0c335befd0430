// Post-order walks of 4-state models that keep their operands on chip: the store walk (every node block written to the
// partials array), the parking plan both walks share, and the host side of the lnL-only walk, whose device code and
// launch templates live in pair_walk.cuh (instantiated in clv_dna_pair_k*.cu).
#include <algorithm>
#include <cstdlib>

#include "pair_walk.cuh"

namespace phb {

namespace {

// ---- the same walk with every node block written to the caller-visible partials array --------------------------
// (PHB_MODE_RESIDENT: what TreeModel.partials, the pre-order pass and the derivative kernels read.)  Differences to
// the lnL-only kernel: a parked operand is the producer row's block of the partials array [row][S][K][4] (+ its
// exponents [row][S]), fetched one row ahead into a pattern-major operand tile; every row's result goes registers ->
// padded staging tile -> coalesced streaming 128-bit stores; a lane owns patterns l and l + 32 of the tile so that
// both tiles are conflict-free; there is no root step and no scratch.
struct PairStoreArgs {
    const PairRow* rows;
    int n_steps;
    const unsigned char* opbase;
    const uint8_t* codes;
    size_t pitch;
    double* clv;       // [row][S][K][4]
    int32_t* scale;    // [row][S]
    int64_t S, n_tiles;
};

template <int K, int NC, int PPT>
__global__ void __launch_bounds__(32, 11) dna_pair_store_kernel(const PairStoreArgs p) {
    using L = PairLayout<K, NC, PPT>;
    constexpr int ROWB = K * 32 + 16, TILE_BYTES = L::TILE * ROWB, OPIN_BYTES = TILE_BYTES;   // exponents sit in the row padding
    constexpr int PIECES = K * 2;                        // 16-byte pieces per pattern
    constexpr int ROUNDS = L::TILE * PIECES / 32;        // warp-wide copy rounds per block
    extern __shared__ __align__(128) unsigned char smem[];
    const int lane = threadIdx.x;
    PairRow* const s_desc = reinterpret_cast<PairRow*>(smem);
    unsigned char* const s_stage = smem + L::DESC_BYTES;
    unsigned char* const s_opin = s_stage + 2 * L::STAGE_BYTES;
    unsigned char* const s_out = s_opin + OPIN_BYTES;
    const int wstride = gridDim.x, n_steps = p.n_steps;
    const int n_tiles = (int)p.n_tiles;
    const size_t S = (size_t)p.S;

    auto fetch_block = [&](int prod_row, int t) {
        const size_t site0 = (size_t)t * L::TILE;
        const int valid = (int)min((int64_t)L::TILE, p.S - (int64_t)site0);
        const unsigned char* src = reinterpret_cast<const unsigned char*>(p.clv + ((size_t)prod_row * S + site0) * (K * 4));
#pragma unroll
        for (int j = 0; j < ROUNDS; ++j) {
            const int c = lane + 32 * j;
            if (c < valid * PIECES) cp_async16(s_opin + (c / PIECES) * ROWB + (c % PIECES) * 16, src + (size_t)c * 16);
        }
        const int32_t* ex = p.scale + (size_t)prod_row * S + site0;
#pragma unroll
        for (int q = 0; q < PPT; ++q)
            if (lane + 32 * q < valid)
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(
                                                                                s_opin + (lane + 32 * q) * ROWB + K * 32)),
                             "l"(ex + lane + 32 * q)
                             : "memory");
    };
    auto stage_row = [&](const PairRow d, int t, int q) {
        unsigned char* st = s_stage + q * L::STAGE_BYTES;
        const int kind_a = (d.packed >> 24) & 3, kind_b = (d.packed >> 26) & 3;
        const unsigned char* ga = p.opbase + (size_t)d.off_a * 16 + lane * 16;
        const unsigned char* gb = p.opbase + (size_t)d.off_b * 16 + lane * 16;
#pragma unroll
        for (int j = 0; j < L::ROUNDS; ++j)
            if (j < L::P_ROUNDS || kind_a == KIND_TIP) cp_async16(st + j * 512 + lane * 16, ga + j * 512);
#pragma unroll
        for (int j = 0; j < L::ROUNDS; ++j)
            if (j < L::P_ROUNDS || kind_b == KIND_TIP) cp_async16(st + L::OPER_BYTES + j * 512 + lane * 16, gb + j * 512);
        constexpr int CL = L::TILE / 16;   // code rows are pitched and zero-padded: a whole tile can always be read
        const int which = lane >> 3, piece = lane & 7;
        const bool tip = which == 0 ? kind_a == KIND_TIP : kind_b == KIND_TIP;
        if (which < 2 && piece < CL && tip) {
            const int tip_row = which == 0 ? d.src_a : (int)(d.packed & 0xffffff);
            cp_async16(st + L::CODES_OFF + which * L::TILE + piece * 16,
                       p.codes + (size_t)tip_row * p.pitch + (size_t)t * L::TILE + piece * 16);
        }
    };

    int tile = blockIdx.x;
    if (tile >= n_tiles) return;
    if (lane < 2) cp_async16(&s_desc[lane], &p.rows[lane < n_steps ? lane : 0]);
    cp_async_commit();
    cp_async_wait_all();
    __syncwarp();
    stage_row(s_desc[0], tile, 0);
    cp_async_commit();

    double prev[PPT][K][4];
    int pe[PPT];
#pragma unroll
    for (int q = 0; q < PPT; ++q) {
        pe[q] = 0;
#pragma unroll
        for (int k = 0; k < K; ++k)
#pragma unroll
            for (int i = 0; i < 4; ++i) prev[q][k][i] = 0.0;
    }
    int row = 0, q = 0;
    int row2 = n_steps > 2 ? 2 : 0;
    while (true) {
        int row_n = row + 1, tile_n = tile;
        if (row_n == n_steps) {
            row_n = 0;
            tile_n += wstride;
        }
        const bool has_next = tile_n < n_tiles;
        cp_async_wait_all();
        __syncwarp();
        const uint32_t pk = s_desc[q & 3].packed;
        const int kinds = (pk >> 24) & 15;
        const bool opin_busy = (kinds >> 2) == KIND_SLOT;
        bool fetch_late = false;
        int slot_n = 0;
        if (lane == 0) cp_async16(&s_desc[(q + 2) & 3], &p.rows[row2]);
        if (has_next) {
            const PairRow dn = s_desc[(q + 1) & 3];
            stage_row(dn, tile_n, (q + 1) & 1);
            if (((dn.packed >> 26) & 3) == KIND_SLOT) {
                slot_n = dn.packed & 0xffffff;       // producer row of the parked operand
                PHB_DCHECK(slot_n < row_n && tile_n < n_tiles);   // written by this warp, for this tile, in an earlier row
                if (opin_busy) fetch_late = true;
                else fetch_block(slot_n, tile_n);
            }
        }
        cp_async_commit();

        const unsigned char* st = s_stage + (q & 1) * L::STAGE_BYTES;
        const int kind_a = kinds & 3, kind_b = kinds >> 2;
        if (kind_b == KIND_SLOT) {
            if (kind_a == KIND_PREV) pair_update<K, NC, PPT, CODES_BYTE, KIND_PREV, KIND_SLOT, LAYOUT_ARRAY>(st, s_opin, lane, prev, pe);
            else pair_update<K, NC, PPT, CODES_BYTE, KIND_TIP, KIND_SLOT, LAYOUT_ARRAY>(st, s_opin, lane, prev, pe);
        } else if (kind_b == KIND_PREV) {
            pair_update<K, NC, PPT, CODES_BYTE, KIND_TIP, KIND_PREV, LAYOUT_ARRAY>(st, s_opin, lane, prev, pe);
        } else {
            pair_update<K, NC, PPT, CODES_BYTE, KIND_TIP, KIND_TIP, LAYOUT_ARRAY>(st, s_opin, lane, prev, pe);
        }
        if (fetch_late) {   // each lane only ever touches its own rows of the operand tile
            __syncwarp();
            fetch_block(slot_n, tile_n);
            cp_async_commit();
        }
        // store: registers -> padded staging tile -> coalesced streaming stores into the block of this row
        {
            const size_t site0 = (size_t)tile * L::TILE;
            const int valid = (int)min((int64_t)L::TILE, p.S - (int64_t)site0);
            unsigned char* dst = reinterpret_cast<unsigned char*>(p.clv + ((size_t)row * S + site0) * (K * 4));
#pragma unroll
            for (int h = 0; h < PPT; ++h) {   // 32 patterns at a time through a half-size staging tile
                if (h) __syncwarp();
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    unsigned char* d = s_out + lane * ROWB + k * 32;
                    *reinterpret_cast<double2*>(d) = make_double2(prev[h][k][0], prev[h][k][1]);
                    *reinterpret_cast<double2*>(d + 16) = make_double2(prev[h][k][2], prev[h][k][3]);
                }
                __syncwarp();
#pragma unroll
                for (int j = 0; j < ROUNDS / PPT; ++j) {
                    const int c = lane + 32 * j;
                    if (c + h * 32 * PIECES < valid * PIECES) {
                        const int4 v = *reinterpret_cast<const int4*>(s_out + (c / PIECES) * ROWB + (c % PIECES) * 16);
                        __stcs(reinterpret_cast<int4*>(dst + (size_t)(c + h * 32 * PIECES) * 16), v);
                    }
                }
            }
#pragma unroll
            for (int h = 0; h < PPT; ++h)
                if (lane + 32 * h < valid) p.scale[(size_t)row * S + site0 + lane + 32 * h] = pe[h];
        }
        if (!has_next) break;
        row = row_n;
        tile = tile_n;
        if (++row2 == n_steps) row2 = 0;
        ++q;
    }
    cp_async_wait_all();
}

template <int K, int NC, int PPT>
int launch_pair_store(Ctx* c, int n_steps) {
    using L = PairLayout<K, NC, PPT>;
    constexpr int ROWB = K * 32 + 16;
    PairStoreArgs a;
    a.rows = static_cast<const PairRow*>(c->d_res_rows);
    a.n_steps = n_steps;
    a.opbase = reinterpret_cast<const unsigned char*>(c->d_pmats);
    a.codes = c->d_codes;
    a.pitch = c->code_pitch;
    a.clv = c->d_clv;
    a.scale = c->d_scale;
    a.S = c->S;
    a.n_tiles = (c->S + L::TILE - 1) / L::TILE;
    auto kern = dna_pair_store_kernel<K, NC, PPT>;
    const size_t smem = L::DESC_BYTES + 2 * L::STAGE_BYTES + (size_t)L::TILE * ROWB + 32 * ROWB;   // operand tile + half-size staging tile
    if (smem > c->smem_optin) return c->fail(PHB_ERR_UNSUPPORTED, "pair store kernel: does not fit in shared memory");
    PHB_CUDA(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    PHB_CUDA(c, cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    int per_sm = 0;
    PHB_CUDA(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 32, smem));
    if (per_sm < 1) per_sm = 1;
    const int64_t grid = std::max<int64_t>(1, std::min<int64_t>(a.n_tiles, (int64_t)c->sm_count * per_sm));
    kern<<<(int)grid, 32, smem, c->stream>>>(a);
    c->launches++;
    PHB_CUDA(c, cudaGetLastError());
    c->resident_warps = per_sm;
    return PHB_OK;
}

// patterns per lane: 2 everywhere; 4 is built for K = 4 (PHB_PAIR_PPT forces it)
int pair_ppt(const Ctx* c, int mode) {
    (void)mode;
    if (c->K != 4) return 2;
    if (tuning().pair_ppt == 2 || tuning().pair_ppt == 4) return tuning().pair_ppt;
    // A tile is one long sequential job (a walk over the whole tree), so what counts is the number of WAVES: when the
    // 64-pattern tiles need a second, nearly empty wave of warps but the 128-pattern tiles fit in one, the bigger
    // tiles win (125k patterns - the 8-GPU shard of the 1M-pattern alignment: 2.19 vs 2.36 ms); otherwise the
    // 64-pattern tiles do (250k: 4.00 vs 4.47 ms; 1M: 15.3 vs 15.8 ms on the same box).
    // (Tried and dropped, profiles/r02g_tile_mix_ab.jsonl: 96-pattern tiles, nine warps per SM - the walk is bound per
    // sub-partition and nine warps put three on one of them, 2.09 vs 1.96 ms; and 64- and 96-pattern walks mixed in
    // one CTA so that every sub-partition holds 224 patterns - two loop bodies competing for the instruction cache,
    // 2.64 ms.)
    const int64_t t64 = (c->S + 63) / 64, t128 = (c->S + 127) / 128;
    const int64_t wave2 = (int64_t)c->sm_count * PairLayout<4, 8, 2>::MIN_CTAS, wave4 = (int64_t)c->sm_count * PairLayout<4, 8, 4>::MIN_CTAS;
    return (t64 > wave2 && t128 <= wave4) ? 4 : 2;
}

int launch_pair_k(Ctx* c, int mode, int n_steps, int n_slots, int64_t b, int64_t e, double* ps, int max_grid,
                  int* grid_out, int chunk_shift = -1) {
    static_assert(kTipTabCodes == 16, "tip tables are staged with 8 or 16 rows per category");
    const int nc = tip_table_rows(c), ppt = pair_ppt(c, mode);
    switch (c->K) {
        case 4: return nc == 8 ? launch_pair_k4n8(c, ppt, mode, n_steps, n_slots, b, e, ps, max_grid, grid_out, chunk_shift)
                               : launch_pair_k4n16(c, ppt, mode, n_steps, n_slots, b, e, ps, max_grid, grid_out, chunk_shift);
        case 1:
        case 2: return launch_pair_k12(c, nc, mode, n_steps, n_slots, b, e, ps, max_grid, grid_out, chunk_shift);
        case 8: return launch_pair_k8(c, nc, mode, n_steps, n_slots, b, e, ps, max_grid, grid_out, chunk_shift);
    }
    return c->fail(PHB_ERR_UNSUPPORTED, "pair kernel needs K in {1,2,4,8}");
}

// resident_plan rows -> this kernel's descriptors (operand offsets resolved on the host), uploaded
int upload_pair_rows(Ctx* c, const ResPlan& plan, bool sym) {
    if (c->n_codes > kTipTabCodes)
        return c->fail(PHB_ERR_UNSUPPORTED, "pair kernel: look-up tables of more than 16 rows are not covered");
    const size_t tab_bytes = (size_t)c->K * tip_table_rows(c) * 32, p_bytes = (size_t)c->K * (sym ? 80 : 128);
    // symmetric form: an internal operand's block is the packed upper triangle of diag(pi) P in d_rmats
    const size_t p_base = sym ? reinterpret_cast<const unsigned char*>(c->d_rmats) - reinterpret_cast<const unsigned char*>(c->d_pmats) : 0;
    const size_t tab_base = reinterpret_cast<const unsigned char*>(c->d_tiptab) - reinterpret_cast<const unsigned char*>(c->d_pmats);
    std::vector<PairRow> rows(plan.rows.size());
    for (size_t r = 0; r < plan.rows.size(); ++r) {
        const ResRow& s = plan.rows[r];
        const int pidx_a = s.pidx_a, pidx_b = (int)(s.packed & 0xffffff);
        const int kind_a = (s.packed >> 24) & 3, kind_b = (s.packed >> 26) & 3;
        if (kind_a == KIND_SLOT) return c->fail(PHB_ERR_STATE, "pair kernel: plan is not in canonical operand order");
        if (s.src_b < 0 || s.src_b >= (1 << 24)) return c->fail(PHB_ERR_UNSUPPORTED, "pair kernel: too many tips");
        const size_t oa = kind_a == KIND_TIP ? tab_base + pidx_a * tab_bytes : p_base + pidx_a * p_bytes;
        const size_t ob = kind_b == KIND_TIP ? tab_base + pidx_b * tab_bytes : p_base + pidx_b * p_bytes;
        PairRow d;
        d.off_a = (uint32_t)(oa / 16);
        d.off_b = (uint32_t)(ob / 16);
        d.src_a = s.src_a;
        d.packed = (uint32_t)s.src_b | (s.packed & 0xff000000u);
        rows[r] = d;
    }
    PHB_CUDA(c, cudaMemcpyAsync(c->d_res_rows, rows.data(), rows.size() * sizeof(PairRow), cudaMemcpyHostToDevice, c->stream));
    PHB_CUDA(c, cudaStreamSynchronize(c->stream));   // `rows` is a stack object
    return PHB_OK;
}

// plan + upload unless d_res_rows already holds exactly this plan (kind 1: lnL-only walk with the root step on edge
// (root_a, root_b); kind 2: every block stored, no root step)
int cached_pair_plan(Ctx* c, int kind, int root_a, int root_b, int* n_steps, int* n_slots) {
    auto& rc = c->res_cache;
    const bool sym = kind == 1 && pair_sym(c);
    if (rc.kind == kind && rc.root_a == root_a && rc.root_b == root_b && rc.gen == c->sched_gen && rc.sym == sym) {
        *n_steps = rc.n_steps;
        *n_slots = rc.n_slots;
        return PHB_OK;
    }
    rc.kind = 0;
    ResPlan plan;
    int st = kind == 2 ? plan_rows(c, -1, -1, false, true, &plan) : plan_rows(c, root_a, root_b, true, false, &plan);
    if (st) return st;
    *n_steps = (int)plan.rows.size();
    *n_slots = plan.n_slots;
    if (plan.rows.empty()) return PHB_OK;
    st = upload_pair_rows(c, plan, sym);
    if (st) return st;
    rc.kind = kind;
    rc.sym = sym;
    rc.root_a = root_a;
    rc.root_b = root_b;
    rc.gen = c->sched_gen;
    rc.n_steps = *n_steps;
    rc.n_slots = *n_slots;
    return PHB_OK;
}

}  // namespace

// Parking plan of the operand-resident walks: the schedule is walked like a register allocator - a finished block that
// the next row does not consume is parked in the lowest free scratch slot (lnL-only walk) or simply lives in the partials
// array (store walk); a slot is recycled once its block has been read.  Operands come out in the canonical order
// TIP <= PREV <= SLOT; a row with two parked operands means the schedule is not a post-order walk: rejected.
int plan_rows(Ctx* c, int root_a, int root_b, bool with_root, bool store, ResPlan* out) {
    const int n_rows = c->n_rows();
    std::vector<int> park_of_node(c->n_nodes, -1);
    std::vector<int> consumer_row(c->n_nodes, -1);
    for (int r = 0; r < n_rows; ++r)
        for (int i = 1; i <= 2; ++i) consumer_row[c->rows_raw[3 * r + i]] = r;
    if (with_root) {
        if (c->node_tip[root_a] < 0) consumer_row[root_a] = n_rows;
        if (c->node_tip[root_b] < 0) consumer_row[root_b] = n_rows;
    }
    std::vector<char> busy;
    auto grab = [&]() {
        for (size_t i = 0; i < busy.size(); ++i)
            if (!busy[i]) {
                busy[i] = 1;
                return (int)i;
            }
        busy.push_back(1);
        return (int)busy.size() - 1;
    };
    auto make = [&](int r, int node_a, int node_b, int pidx_a, int pidx_b, int dst_node) -> int {
        int nodes[2] = {node_a, node_b}, pidx[2] = {pidx_a, pidx_b}, kind[2], src[2];
        for (int i = 0; i < 2; ++i) {
            const int nd = nodes[i];
            if (c->node_tip[nd] >= 0) {
                kind[i] = KIND_TIP;
                src[i] = c->node_tip[nd];
            } else if (c->node_row[nd] == r - 1) {
                kind[i] = KIND_PREV;
                src[i] = 0;
            } else {
                kind[i] = KIND_SLOT;
                src[i] = store ? c->node_row[nd] : park_of_node[nd];
                if (src[i] < 0) return c->fail(PHB_ERR_STATE, "resident plan: operand was never parked");
                if (!store) busy[src[i]] = 0;   // recycled after this row has read it
            }
        }
        if (kind[0] == KIND_SLOT && kind[1] == KIND_SLOT)
            return c->fail(PHB_ERR_UNSUPPORTED,
                           "resident kernel needs a post-order schedule (second child = previous row); "
                           "use Traversal.locality_order() or another mode");
        if (kind[0] > kind[1]) {   // canonical operand order TIP <= PREV <= SLOT (children commute)
            std::swap(kind[0], kind[1]);
            std::swap(src[0], src[1]);
            std::swap(pidx[0], pidx[1]);
        }
        int dst = 15;
        if (!store && dst_node >= 0 && consumer_row[dst_node] != r + 1 && consumer_row[dst_node] >= 0) {
            dst = grab();
            if (dst >= kScratchSlots) return c->fail(PHB_ERR_UNSUPPORTED, "resident plan: tree needs more than 15 parked blocks");
            park_of_node[dst_node] = dst;
        }
        if (pidx[1] >= (1 << 24)) return c->fail(PHB_ERR_UNSUPPORTED, "resident plan: too many rows");
        ResRow row;
        row.src_a = src[0];
        row.src_b = src[1];
        row.pidx_a = pidx[0];
        row.packed = (uint32_t)pidx[1] | ((uint32_t)kind[0] << 24) | ((uint32_t)kind[1] << 26) | ((uint32_t)dst << 28);
        out->rows.push_back(row);
        return PHB_OK;
    };
    out->rows.clear();
    for (int r = 0; r < n_rows; ++r) {
        int st = make(r, c->rows_raw[3 * r + 1], c->rows_raw[3 * r + 2], 2 * r, 2 * r + 1, c->rows_raw[3 * r]);
        if (st) return st;
    }
    if (with_root) {
        const int rp = 2 * c->max_rows();
        int st = make(n_rows, root_a, root_b, rp, rp + 1, -1);
        if (st) return st;
    }
    out->n_slots = std::max<int>((int)busy.size(), 1);
    return PHB_OK;
}

// Post-order pass with every node block stored (PHB_MODE_RESIDENT / PHB_MODE_AUTO of phb_compute_partials)
int dna_pair_store(Ctx* c) {
    int n_steps = 0, n_slots = 0;
    int st = cached_pair_plan(c, 2, -1, -1, &n_steps, &n_slots);
    if (st) return st;
    if (n_steps == 0) return PHB_OK;
    switch (c->K * 100 + tip_table_rows(c)) {
        case 408: return launch_pair_store<4, 8, 2>(c, n_steps);
        case 108: return launch_pair_store<1, 8, 2>(c, n_steps);
        case 116: return launch_pair_store<1, 16, 2>(c, n_steps);
        case 208: return launch_pair_store<2, 8, 2>(c, n_steps);
        case 216: return launch_pair_store<2, 16, 2>(c, n_steps);
        case 416: return launch_pair_store<4, 16, 2>(c, n_steps);
        // eight categories: one pattern per lane (two would need 128 registers for the block alone)
        case 808: return launch_pair_store<8, 8, 1>(c, n_steps);
        case 816: return launch_pair_store<8, 16, 1>(c, n_steps);
    }
    return PHB_ERR_UNSUPPORTED;
}

// One evaluation from the tip codes resident on the device: per-pattern lnL + their weighted sum in d_result[0]
int dna_pair_lnl(Ctx* c, int root_a, int root_b) {
    int n_steps = 0, n_slots = 0;
    int st = cached_pair_plan(c, 1, root_a, root_b, &n_steps, &n_slots);
    if (st) return st;
    int grid = 0;
    const int tile = 32 * pair_ppt(c, c->codes_mode);
    const int64_t n_tiles = (c->S + tile - 1) / tile;
    st = launch_pair_k(c, c->codes_mode, n_steps, n_slots, 0, n_tiles, c->d_partial_sums, kPartialCap, &grid);
    if (st) return st;
    c->resident_slots = n_slots;
    return launch_final_reduce(c, c->d_partial_sums, grid, 1, c->d_result);
}

// Whole evaluation starting from HOST tip codes, in ONE launch.  The pattern axis is cut into chunks; the copy engine
// moves chunk after chunk on a second stream and drops a 4-byte flag behind each one; the kernel - already running,
// every SM busy - starts a tile as soon as the flag of its chunk shows this evaluation's epoch (patterns are
// independent).  No per-chunk launches, no tail per chunk; the copy of chunk i+1 overlaps the pruning of chunk i.
// packed: two 4-bit codes per byte (even pattern in the low nibble), rows of (S + 1) / 2 bytes - half the bytes over
// PCIe.  Flags and data are written by memcpy from pinned memory only (copy engine): nothing here needs an SM while
// the kernel occupies all of them.  One synchronisation at the very end (caller).
// mode CODES_SPLIT3: codes_host is the plane of 2-bit values, rows of (S + 3) / 4 bytes, codes_hi_host the plane of high
// bits, rows of (S + 7) / 8 bytes - 3 / 8 of a byte per code over PCIe.
int dna_pair_from_host(Ctx* c, const uint8_t* codes_host, const uint8_t* codes_hi_host, int mode, int n_chunks, int root_a,
                       int root_b, int slot) {
    const bool packed = mode == CODES_NIBBLE;
    const bool pipelined = slot >= 0;
    const int s = pipelined ? slot : 0;
    if (pipelined && mode == CODES_BYTE) return c->fail(PHB_ERR_UNSUPPORTED, "pipelined host-fed evaluations take packed codes (two slots must fit the code buffer)");
    int n_steps = 0, n_slots = 0;
    int st = cached_pair_plan(c, 1, root_a, root_b, &n_steps, &n_slots);
    if (st) return st;
    if (c->copy_stream == nullptr) {
        PHB_CUDA(c, cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
        for (int i = 0; i < kMaxChunks; ++i) PHB_CUDA(c, cudaEventCreateWithFlags(&c->chunk_events[i], cudaEventDisableTiming));
        PHB_CUDA(c, cudaEventCreateWithFlags(&c->start_event, cudaEventDisableTiming));
        for (int i = 0; i < 2; ++i) {
            PHB_CUDA(c, cudaEventCreateWithFlags(&c->slot_done[i], cudaEventDisableTiming));
            PHB_CUDA(c, cudaEventCreateWithFlags(&c->result_event[i], cudaEventDisableTiming));
            PHB_CUDA(c, cudaEventCreateWithFlags(&c->copies_done[i], cudaEventDisableTiming));
        }
    }
    if (c->h_epoch == nullptr) {
        PHB_CUDA(c, cudaHostAlloc(reinterpret_cast<void**>(&c->h_epoch), 128, cudaHostAllocDefault));   // one word per slot, 64 bytes apart
        PHB_CUDA(c, cudaHostAlloc(reinterpret_cast<void**>(&c->h_results), 64, cudaHostAllocDefault));
        PHB_CUDA(c, cudaMemsetAsync(c->d_flags, 0, 2 * (kMaxFlagChunks + 1) * sizeof(int), c->stream));
        PHB_CUDA(c, cudaStreamSynchronize(c->stream));   // once per context: the copy stream must not race the reset
    }
    const int tile = 32 * pair_ppt(c, mode);
    const int64_t n_tiles = (c->S + tile - 1) / tile;
    n_chunks = (int)std::max<int64_t>(1, std::min<int64_t>(std::min(n_chunks, kMaxFlagChunks), n_tiles));
    int chunk_shift = 0;   // chunks are a power of two of tiles: the kernel finds a tile's flag with a shift
    while (((int64_t)1 << chunk_shift) * n_chunks < n_tiles) ++chunk_shift;
    const int64_t tpc = (int64_t)1 << chunk_shift;
    n_chunks = (int)((n_tiles + tpc - 1) / tpc);
    c->flag_epoch = c->flag_epoch >= (1 << 30) ? 1 : c->flag_epoch + 1;
    int* const h_epoch = c->h_epoch + 16 * s;      // the copy engine reads it when the flag copy executes: one word per slot
    // ... so the flag copies of the slot's previous evaluation must have executed before the word changes (they have,
    // unless the caller submits a third evaluation while the copies of the first are still queued)
    if (pipelined) PHB_CUDA(c, cudaEventSynchronize(c->copies_done[s]));
    *h_epoch = c->flag_epoch;
    int* const d_flags = c->d_flags + s * (kMaxFlagChunks + 1);
    c->d_flags_cur = d_flags;
    if (pipelined) {
        // the copies may run ahead of the compute stream - that is the point - but not into a slot a queued walk still reads
        PHB_CUDA(c, cudaStreamWaitEvent(c->copy_stream, c->slot_done[s], 0));
    } else {
        // the copy stream must not overtake work already queued on the compute stream (previous evaluation, flag reset)
        PHB_CUDA(c, cudaEventRecord(c->start_event, c->stream));
        PHB_CUDA(c, cudaStreamWaitEvent(c->copy_stream, c->start_event, 0));
    }
    const bool split = mode == CODES_SPLIT3;
    const size_t host_row = split ? ((size_t)c->S + 3) / 4 : (packed ? ((size_t)c->S + 1) / 2 : (size_t)c->S);
    const size_t dev_pitch = split ? c->code_pitch / 4 : (packed ? c->code_pitch / 2 : c->code_pitch);
    const int per_tile = split ? tile / 4 : (packed ? tile / 2 : tile);   // bytes of one tile in a code row
    const size_t host_row_hi = ((size_t)c->S + 7) / 8, dev_pitch_hi = c->code_pitch / 8;
    uint8_t* const d_lo = c->d_codes_ws + (size_t)s * ((size_t)c->n_tips * c->code_pitch / 2);   // slot 1: the upper half
    uint8_t* const d_hi = d_lo + (size_t)c->n_tips * (c->code_pitch / 4);
    c->codes_packed = mode != CODES_BYTE;
    c->codes_mode = mode;
    c->d_codes = d_lo;
    for (int i = 0; i < n_chunks; ++i) {
        const int64_t b = tpc * i, e = std::min<int64_t>(tpc * (i + 1), n_tiles);
        const size_t c0 = (size_t)b * per_tile, c1 = std::min<size_t>((size_t)e * per_tile, host_row);
        PHB_CUDA(c, cudaMemcpy2DAsync(d_lo + c0, dev_pitch, codes_host + c0, host_row, c1 - c0, (size_t)c->n_tips,
                                      cudaMemcpyHostToDevice, c->copy_stream));
        if (split) {
            const size_t h0 = (size_t)b * (tile / 8), h1 = std::min<size_t>((size_t)e * (tile / 8), host_row_hi);
            PHB_CUDA(c, cudaMemcpy2DAsync(d_hi + h0, dev_pitch_hi, codes_hi_host + h0, host_row_hi, h1 - h0, (size_t)c->n_tips,
                                          cudaMemcpyHostToDevice, c->copy_stream));
        }
        PHB_CUDA(c, cudaMemcpyAsync(d_flags + i, h_epoch, sizeof(int), cudaMemcpyHostToDevice, c->copy_stream));
    }
    if (pipelined) PHB_CUDA(c, cudaEventRecord(c->copies_done[s], c->copy_stream));
    int grid = 0;
    st = launch_pair_k(c, mode, n_steps, n_slots, 0, n_tiles, c->d_partial_sums, kPartialCap, &grid, chunk_shift);
    if (st) return st;
    c->resident_slots = n_slots;
    if (!pipelined) {
        c->pipelined_pending = true;
        return launch_final_reduce(c, c->d_partial_sums, grid, 1, c->d_result);
    }
    st = launch_final_reduce(c, c->d_partial_sums, grid, 1, c->d_result + s);
    if (st) return st;
    PHB_CUDA(c, cudaEventRecord(c->slot_done[s], c->stream));
    return PHB_OK;
}

}  // namespace phb
