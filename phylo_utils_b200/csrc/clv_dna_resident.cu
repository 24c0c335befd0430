// Operand-stack-resident pruning for 4-state models: the whole post-order walk of a pattern tile in ONE
// launch, with no dependency between warps and no operand ever waited for.
//
// Same arithmetic as clv_dna.cu (reference `clv`, numba_likelihood_engine.py:10-46; root = tree_model.py:178-217),
// different data movement.  Patterns are independent, and a post-order walk is a stack machine: with the
// larger child subtree finished first, 2/3 of the internal operands are the block produced by the row just
// before, and a 1000-taxon tree never has more than ~6 finished-but-unconsumed blocks alive.  So:
//
//   * every WARP owns a tile of 32 patterns (lane = pattern, the K categories are looped inside the thread)
//     and walks ALL rows for it, independently of every other warp (no block barrier in the row loop);
//   * the result of a row stays in REGISTERS when the next row consumes it;
//   * a result needed later is "parked": written through a padded shared-memory staging tile with coalesced
//     128-bit stores either to the caller-visible partials array (STORE = true, every row) or to a small
//     per-warp scratch area that lives in L2 (STORE = false, ~1/3 of the rows), and fetched back with
//     cp.async ONE ROW BEFORE it is consumed;
//   * everything else a row needs - its 16-byte descriptor, the two P blocks (2*K*128 B), the tile's tip
//     codes (32 B per tip operand) - is read-only and is also brought in by cp.async one row ahead
//     (descriptors two rows ahead), into double buffers private to the warp.  Each lane reads back exactly
//     the 16-byte chunks it wrote itself, so no cross-thread memory ordering is involved;
//   * ROOT = true appends the virtual-root step (root combine, pi-dot, Gamma mixture, log, weighted sum) as
//     a final pseudo-row, so one launch yields the per-pattern lnL and the block sums.
//
// One thread per pattern means the per-pattern maximum / exponent is thread-local (no shuffles), P rows are
// warp-wide broadcast reads, and all bookkeeping is paid once per 32 pattern-node updates.
// Shared memory per warp is ~16 KB (K = 4), so 12+ warps per SM stay resident.
#include <algorithm>
#include <cstdlib>

#include "common.cuh"
#include "resident_plan.cuh"

namespace phb {

namespace {

constexpr int kMaxWarps = 1;      // one warp per CTA: warps share nothing but the (tiny) look-up table
constexpr int kMinCtas = 16;      // register budget: 65536 / (16 * 32) = 128 per thread

struct ResArgs {
    const ResRow* rows;     // [n_steps]
    int n_steps;            // rows walked per tile (n_rows, +1 with ROOT)
    const double* pmats;
    const double* tiptab;   // [pidx][K][NC][4]: P . lut[code], what a tip operand contributes
    const uint8_t* codes;
    size_t pitch;
    double* clv;            // STORE: partials [row][S][K][4]
    int32_t* scale;         // STORE: exponents [row][S]
    unsigned char* scratch; // !STORE: [warp][slot][block | 32 exponents]
    int n_slots;
    const double* freqs;    // ROOT
    const double* catw;
    const double* weights;
    double* pattern_lnl;
    double* partial_sums;
    int64_t S;
    int64_t wt_begin, wt_end;   // warp tiles (32 patterns each) covered by this launch
    int warp_bytes;
};

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gmem_src) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(s), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// geometry of one warp's shared memory.  A warp tile is 32 patterns.
template <int K, int NC>
struct WarpLayout {
    static constexpr int SPW = 32;
    static constexpr int TAB_BYTES = K * NC * 32;            // tip table of one operand: [k][code][4 doubles]
    static constexpr int OPER_BYTES = TAB_BYTES > K * 128 ? TAB_BYTES : K * 128;   // ... or its P block [k][16 doubles]
    static constexpr int P_BYTES = 2 * OPER_BYTES;
    static constexpr int STAGE_BYTES = P_BYTES + 2 * 32;     // + 32 codes per operand
    static constexpr int DESC_BYTES = 4 * 16;                // descriptor ring (two rows ahead, double use)
    static constexpr int ROWB = K * 32 + 16;                 // one pattern's block row, padded: conflict-free LDS.128
    static constexpr int TILE_BYTES = 32 * ROWB;             // a parked block in shared memory
    static constexpr int OPIN_BYTES = TILE_BYTES + 128;      // + 32 exponents
    static constexpr int BLOCK_BYTES = 32 * K * 32;          // a parked block in global memory (dense)
    static constexpr int SCRATCH_SLOT = BLOCK_BYTES + 128;
    static constexpr int CHUNKS = BLOCK_BYTES / 16;          // 16-byte chunks per block
    // two operand tiles; the tile of the row being computed doubles as its output staging tile once the
    // operand has been consumed
    static constexpr int WARP_BYTES = DESC_BYTES + 2 * STAGE_BYTES + 2 * OPIN_BYTES;
};

struct Cursor {
    int row;
    int64_t wt;
};

// prev[k] <- (P1[k] . a[k]) * (P2[k] . b[k]) for this lane's pattern; returns the cumulative exponent
template <int K, int NC, int KA, int KB>
__device__ __forceinline__ int row_update(const unsigned char* st, const unsigned char* oa, const unsigned char* ob,
                                          int lane, double (&prev)[K][4], int prev_e) {
    using L = WarpLayout<K, NC>;
    // a tip operand contributes the row `code` of its staged table T[k] = P[k] . lut - no arithmetic
    int e = 0;
    const unsigned char* ta = st + (KA == KIND_TIP ? (int)st[L::P_BYTES + lane] * 32 : 0);
    const unsigned char* tb = st + L::OPER_BYTES + (KB == KIND_TIP ? (int)st[L::P_BYTES + 32 + lane] * 32 : 0);
    if (KA == KIND_PREV || KB == KIND_PREV) e += prev_e;
    if (KA == KIND_SLOT) e += *reinterpret_cast<const int*>(oa + L::TILE_BYTES + lane * 4);
    if (KB == KIND_SLOT) e += *reinterpret_cast<const int*>(ob + L::TILE_BYTES + lane * 4);
    int mh = 0;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        double a[4], b[4], x[4], y[4];
        if (KA == KIND_TIP) {
            const double2 lo = *reinterpret_cast<const double2*>(ta + k * NC * 32);
            const double2 hi = *reinterpret_cast<const double2*>(ta + k * NC * 32 + 16);
            x[0] = lo.x; x[1] = lo.y; x[2] = hi.x; x[3] = hi.y;
        } else if (KA == KIND_PREV) {
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = prev[k][i];
        } else {
            const double2 lo = *reinterpret_cast<const double2*>(oa + lane * L::ROWB + k * 32);
            const double2 hi = *reinterpret_cast<const double2*>(oa + lane * L::ROWB + k * 32 + 16);
            a[0] = lo.x; a[1] = lo.y; a[2] = hi.x; a[3] = hi.y;
        }
        if (KB == KIND_TIP) {
            const double2 lo = *reinterpret_cast<const double2*>(tb + k * NC * 32);
            const double2 hi = *reinterpret_cast<const double2*>(tb + k * NC * 32 + 16);
            y[0] = lo.x; y[1] = lo.y; y[2] = hi.x; y[3] = hi.y;
        } else if (KB == KIND_PREV) {
#pragma unroll
            for (int i = 0; i < 4; ++i) b[i] = prev[k][i];
        } else {
            const double2 lo = *reinterpret_cast<const double2*>(ob + lane * L::ROWB + k * 32);
            const double2 hi = *reinterpret_cast<const double2*>(ob + lane * L::ROWB + k * 32 + 16);
            b[0] = lo.x; b[1] = lo.y; b[2] = hi.x; b[3] = hi.y;
        }
        // P rows are read as warp-wide broadcasts (every lane, same address)
        if (KA != KIND_TIP) {
            const double2* q1 = reinterpret_cast<const double2*>(st + k * 128);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const double2 r0 = q1[2 * i], r1 = q1[2 * i + 1];
                x[i] = fma(r1.y, a[3], fma(r1.x, a[2], fma(r0.y, a[1], r0.x * a[0])));
            }
        }
        if (KB != KIND_TIP) {
            const double2* q2 = reinterpret_cast<const double2*>(st + L::OPER_BYTES + k * 128);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const double2 t0 = q2[2 * i], t1 = q2[2 * i + 1];
                y[i] = fma(t1.y, b[3], fma(t1.x, b[2], fma(t0.y, b[1], t0.x * b[0])));
            }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const double o = x[i] * y[i];
            prev[k][i] = o;
            mh = max(mh, __double2hiint(o));   // partials are >= 0: the high word orders them
        }
    }
    const bool small = mh < kScaleThresholdHi && mh >= 0x00100000;   // 0 < max < 2^-128
    if (__any_sync(0xffffffffu, small)) {
        if (small) {
            const int shift = 1023 - (mh >> 20);
            const double f = pow2i(shift);
#pragma unroll
            for (int k = 0; k < K; ++k)
#pragma unroll
                for (int i = 0; i < 4; ++i) prev[k][i] *= f;
            e -= shift;
        }
    }
    return e;
}

template <int K, int NC, bool STORE, bool ROOT>
__global__ void __launch_bounds__(kMaxWarps * 32, kMinCtas) dna_resident_kernel(const ResArgs p) {
    using L = WarpLayout<K, NC>;
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ double s_red[kMaxWarps];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, n_warps = blockDim.x >> 5;
    unsigned char* wbase = smem + (size_t)warp * p.warp_bytes;
    ResRow* s_desc = reinterpret_cast<ResRow*>(wbase);
    unsigned char* s_stage = wbase + L::DESC_BYTES;
    unsigned char* s_opin = s_stage + 2 * L::STAGE_BYTES;

    const int64_t n_wt = p.wt_end;
    const int64_t gwarp = (int64_t)blockIdx.x * n_warps + warp;
    const int64_t wstride = (int64_t)gridDim.x * n_warps;
    const int n_steps = p.n_steps;
    const size_t S = (size_t)p.S;
    unsigned char* my_scratch = STORE ? nullptr : p.scratch + (size_t)gwarp * p.n_slots * L::SCRATCH_SLOT;

    Cursor cd{0, p.wt_begin + gwarp};   // descriptor prefetch cursor (two rows ahead)
    Cursor cp = cd;        // data prefetch cursor (one row ahead)
    Cursor cc = cd;        // compute cursor
    int qd = 0, qp = 0, qc = 0;

    auto advance = [&](Cursor& c) {
        if (++c.row == n_steps) {
            c.row = 0;
            c.wt += wstride;
        }
    };
    auto prefetch_desc = [&]() {
        if (cd.wt < n_wt && lane == 0) cp_async16(&s_desc[qd & 3], &p.rows[cd.row]);
        ++qd;
        advance(cd);
    };
    // where a parked block lives in global memory
    auto block_base = [&](int id, int64_t site0, const unsigned char*& exps) -> const unsigned char* {
        if (STORE) {
            exps = reinterpret_cast<const unsigned char*>(p.scale + (size_t)id * S + site0);
            return reinterpret_cast<const unsigned char*>(p.clv + ((size_t)id * S + site0) * (K * 4));
        }
        const unsigned char* b = my_scratch + (size_t)id * L::SCRATCH_SLOT;
        exps = b + L::BLOCK_BYTES;
        return b;
    };
    auto fetch_block = [&](int id, int64_t site0, unsigned char* tile) {
        const unsigned char* exps;
        const unsigned char* src = block_base(id, site0, exps);
        const int64_t valid = STORE ? min((int64_t)32, p.S - site0) * (K * 2) : (int64_t)L::CHUNKS;
#pragma unroll
        for (int j = 0; j < L::CHUNKS / 32; ++j) {
            const int c = lane + 32 * j;                       // chunk -> (pattern, piece)
            if (c < valid) cp_async16(tile + (c / (K * 2)) * L::ROWB + (c % (K * 2)) * 16, src + (size_t)c * 16);
        }
        if (!STORE || site0 + lane < p.S) cp_async4(tile + L::TILE_BYTES + lane * 4, exps + lane * 4);
    };
    auto prefetch_data = [&]() {
        if (cp.wt < n_wt) {
            const ResRow d = s_desc[qp & 3];
            unsigned char* st = s_stage + (size_t)(qp & 1) * L::STAGE_BYTES;
            const int pidx_b = d.packed & 0xffffff, kind_a = (d.packed >> 24) & 3, kind_b = (d.packed >> 26) & 3;
            // per operand: its tip table (tip) or its P block (anything else)
            const char* pa = kind_a == KIND_TIP ? reinterpret_cast<const char*>(p.tiptab + (size_t)d.pidx_a * K * NC * 4)
                                                : reinterpret_cast<const char*>(p.pmats + (size_t)d.pidx_a * K * 16);
            const char* pb = kind_b == KIND_TIP ? reinterpret_cast<const char*>(p.tiptab + (size_t)pidx_b * K * NC * 4)
                                                : reinterpret_cast<const char*>(p.pmats + (size_t)pidx_b * K * 16);
            const int cha = kind_a == KIND_TIP ? L::TAB_BYTES / 16 : K * 8, chb = kind_b == KIND_TIP ? L::TAB_BYTES / 16 : K * 8;
            for (int c = lane; c < cha; c += 32) cp_async16(st + c * 16, pa + c * 16);
            for (int c = lane; c < chb; c += 32) cp_async16(st + L::OPER_BYTES + c * 16, pb + c * 16);
            const int64_t site0 = cp.wt * 32;
            if (kind_a == KIND_TIP && lane < 2)
                cp_async16(st + L::P_BYTES + lane * 16, p.codes + (size_t)d.src_a * p.pitch + site0 + lane * 16);
            if (kind_b == KIND_TIP && lane >= 2 && lane < 4)
                cp_async16(st + L::P_BYTES + 32 + (lane - 2) * 16,
                           p.codes + (size_t)d.src_b * p.pitch + site0 + (lane - 2) * 16);
            unsigned char* opin = s_opin + (size_t)(qp & 1) * L::OPIN_BYTES;
            if (kind_b == KIND_SLOT && kind_a != KIND_SLOT) fetch_block(d.src_b, site0, opin);
        }
        ++qp;
        advance(cp);
    };

    // prologue: descriptors of rows 0 and 1, then the data of row 0
    prefetch_desc();
    prefetch_desc();
    cp_async_commit();
    cp_async_wait_all();
    __syncwarp();
    prefetch_data();
    cp_async_commit();

    double prev[K][4];
    int prev_e = 0;
#pragma unroll
    for (int k = 0; k < K; ++k)
#pragma unroll
        for (int i = 0; i < 4; ++i) prev[k][i] = 0.0;
    double acc = 0.0;

    while (cc.wt < n_wt) {
        cp_async_wait_all();       // everything issued one row ago has had a whole row to land
        __syncwarp();
        prefetch_desc();           // row + 2
        prefetch_data();           // row + 1
        cp_async_commit();

        const ResRow d = s_desc[qc & 3];
        const unsigned char* st = s_stage + (size_t)(qc & 1) * L::STAGE_BYTES;
        unsigned char* opin = s_opin + (size_t)(qc & 1) * L::OPIN_BYTES;
        const int kinds = (d.packed >> 24) & 15, dst_slot = d.packed >> 28;   // kind_a | kind_b << 2
        const int64_t site0 = cc.wt * 32;
        int e;
        switch (kinds) {
            case KIND_TIP | (KIND_TIP << 2):
                e = row_update<K, NC, KIND_TIP, KIND_TIP>(st, opin, opin, lane, prev, prev_e);
                break;
            case KIND_TIP | (KIND_PREV << 2):
                e = row_update<K, NC, KIND_TIP, KIND_PREV>(st, opin, opin, lane, prev, prev_e);
                break;
            case KIND_PREV | (KIND_SLOT << 2):
                e = row_update<K, NC, KIND_PREV, KIND_SLOT>(st, opin, opin, lane, prev, prev_e);
                break;
            case KIND_TIP | (KIND_SLOT << 2):
                e = row_update<K, NC, KIND_TIP, KIND_SLOT>(st, opin, opin, lane, prev, prev_e);
                break;
            default:   // not a row shape the host plan may emit (it rejects SLOT/SLOT rows): do not touch memory
                e = 0;
                break;
        }
        prev_e = e;
        const int64_t s = site0 + lane;
        const bool is_root = ROOT && cc.row == n_steps - 1;
        if (!is_root) {
            if (STORE || dst_slot != 15) {
                // park: registers -> padded staging tile -> coalesced 128-bit stores.  The operand tile of this
                // row is dead by now and serves as the staging tile.
                unsigned char* s_out = opin;
                __syncwarp();
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    *reinterpret_cast<double2*>(s_out + lane * L::ROWB + k * 32) = make_double2(prev[k][0], prev[k][1]);
                    *reinterpret_cast<double2*>(s_out + lane * L::ROWB + k * 32 + 16) = make_double2(prev[k][2], prev[k][3]);
                }
                __syncwarp();
                const unsigned char* exps;
                unsigned char* dst = const_cast<unsigned char*>(block_base(STORE ? cc.row : dst_slot, site0, exps));
                const int64_t valid = STORE ? min((int64_t)32, p.S - site0) * (K * 2) : (int64_t)L::CHUNKS;
#pragma unroll
                for (int j = 0; j < L::CHUNKS / 32; ++j) {
                    const int c = lane + 32 * j;
                    if (c < valid) {
                        const int4 v = *reinterpret_cast<const int4*>(s_out + (c / (K * 2)) * L::ROWB + (c % (K * 2)) * 16);
                        if (STORE) __stcs(reinterpret_cast<int4*>(dst + (size_t)c * 16), v);
                        else *reinterpret_cast<int4*>(dst + (size_t)c * 16) = v;
                    }
                }
                if (!STORE || s < p.S) *reinterpret_cast<int*>(const_cast<unsigned char*>(exps) + lane * 4) = e;
            }
        } else {
            double mix = 0.0;
#pragma unroll
            for (int k = 0; k < K; ++k) {
                double f = p.freqs[0] * prev[k][0];
                f = fma(p.freqs[1], prev[k][1], f);
                f = fma(p.freqs[2], prev[k][2], f);
                f = fma(p.freqs[3], prev[k][3], f);
                if (f > 0) mix = fma(p.catw[k], f, mix);
            }
            if (s < p.S) {
                const double lnl = mix > 0 ? log(mix) + (double)e * kLn2 : -INFINITY;
                p.pattern_lnl[s] = lnl;
                acc += (p.weights ? p.weights[s] : 1.0) * lnl;
            }
        }
        ++qc;
        advance(cc);
    }
    cp_async_wait_all();
    if (ROOT) {
        acc = warp_sum(acc);
        if (lane == 0) s_red[warp] = acc;
        __syncthreads();
        if (threadIdx.x == 0) {
            double t = 0;
            for (int w = 0; w < n_warps; ++w) t += s_red[w];
            p.partial_sums[blockIdx.x] = t;
        }
    }
}

}  // namespace

// ---- host side: parking plan + launch -------------------------------------------------------------------------
// Walk the schedule like a register allocator: a result consumed by the very next row stays in
// registers; anything else is parked.  In STORE mode the parking place is the node's own block of the
// partials array (id = producer row); otherwise the lowest free scratch slot, recycled once consumed.
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

namespace {

template <int K, int NC, bool STORE, bool ROOT>
int launch_resident(Ctx* c, const ResPlan& plan, int64_t wt_begin, int64_t wt_end, double* partial_sums, int max_grid,
                    int* grid_out) {
    using L = WarpLayout<K, NC>;
    ResArgs a;
    a.rows = static_cast<const ResRow*>(c->d_res_rows);
    a.n_steps = (int)plan.rows.size();
    a.pmats = c->d_pmats;
    a.tiptab = c->d_tiptab;
    a.codes = c->d_codes;
    a.pitch = c->code_pitch;
    a.clv = c->d_clv;
    a.scale = c->d_scale;
    a.scratch = c->d_scratch;
    a.n_slots = plan.n_slots;
    a.freqs = c->model_freqs();
    a.catw = c->model_catw();
    a.weights = c->d_weights;
    a.pattern_lnl = c->d_pattern_lnl;
    a.partial_sums = partial_sums;
    a.S = c->S;
    a.wt_begin = wt_begin;
    a.wt_end = wt_end;
    a.warp_bytes = (L::WARP_BYTES + 127) / 128 * 128;
    auto kern = dna_resident_kernel<K, NC, STORE, ROOT>;
    cudaFuncAttributes fa;
    PHB_CUDA(c, cudaFuncGetAttributes(&fa, kern));
    // pick the CTA width that keeps the most warps resident per SM
    const size_t lut_bytes = 0, budget = c->smem_optin, sm_total = c->smem_per_sm;
    int best_w = 1, best_total = 0, best_ctas = 1;
    const int force_w = tuning().resident_warps;
    for (int w = 1; w <= kMaxWarps; ++w) {
        if (force_w && w != force_w) continue;
        const size_t cta = lut_bytes + (size_t)w * a.warp_bytes + 1024;   // + per-CTA reservation
        if (cta > budget) break;
        int ctas = (int)(sm_total / cta);
        const int by_regs = 65536 / (std::max(fa.numRegs, 32) * 32 * w);
        ctas = std::min(ctas, std::min(by_regs, 32));
        if (ctas * w > best_total) {
            best_total = ctas * w;
            best_w = w;
            best_ctas = ctas;
        }
    }
    if (best_total == 0) return c->fail(PHB_ERR_UNSUPPORTED, "resident kernel: does not fit in shared memory");
    const size_t smem = lut_bytes + (size_t)best_w * a.warp_bytes;
    PHB_CUDA(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    PHB_CUDA(c, cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    int per_sm = 0;
    PHB_CUDA(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, best_w * 32, smem));
    if (per_sm < 1) per_sm = 1;
    per_sm = std::min(per_sm, best_ctas);
    const int64_t n_wt = wt_end - wt_begin;
    const int64_t blocks_needed = (n_wt + best_w - 1) / best_w;
    int64_t grid = std::min<int64_t>(blocks_needed, (int64_t)c->sm_count * per_sm);
    grid = std::min<int64_t>(grid, max_grid);
    if (!STORE) {   // every resident warp needs its own scratch stripe
        const int64_t cap = (int64_t)(c->scratch_bytes / ((size_t)plan.n_slots * L::SCRATCH_SLOT)) / best_w;
        if (cap < 1) return c->fail(PHB_ERR_NOMEM, "resident kernel: scratch area too small");
        grid = std::min(grid, cap);
    }
    if (grid < 1) grid = 1;
    kern<<<(int)grid, best_w * 32, smem, c->stream>>>(a);
    c->launches++;
    PHB_CUDA(c, cudaGetLastError());
    c->resident_warps = best_w * per_sm;
    *grid_out = (int)grid;
    return PHB_OK;
}

template <bool STORE, bool ROOT>
int launch_resident_k(Ctx* c, const ResPlan& plan, int64_t b, int64_t e, double* ps, int max_grid, int* grid_out) {
    if (c->n_codes > kTipTabCodes)
        return c->fail(PHB_ERR_UNSUPPORTED, "resident kernel: look-up tables of more than 16 rows are not covered");
    static_assert(kTipTabCodes == 16, "tip tables are staged with 8 or 16 rows per category");
    const int key = c->K * 100 + tip_table_rows(c);
    switch (key) {
        case 108: return launch_resident<1, 8, STORE, ROOT>(c, plan, b, e, ps, max_grid, grid_out);
        case 116: return launch_resident<1, 16, STORE, ROOT>(c, plan, b, e, ps, max_grid, grid_out);
        case 208: return launch_resident<2, 8, STORE, ROOT>(c, plan, b, e, ps, max_grid, grid_out);
        case 216: return launch_resident<2, 16, STORE, ROOT>(c, plan, b, e, ps, max_grid, grid_out);
        case 408: return launch_resident<4, 8, STORE, ROOT>(c, plan, b, e, ps, max_grid, grid_out);
        case 416: return launch_resident<4, 16, STORE, ROOT>(c, plan, b, e, ps, max_grid, grid_out);
        case 808: return launch_resident<8, 8, STORE, ROOT>(c, plan, b, e, ps, max_grid, grid_out);
        case 816: return launch_resident<8, 16, STORE, ROOT>(c, plan, b, e, ps, max_grid, grid_out);
    }
    return c->fail(PHB_ERR_UNSUPPORTED, "resident kernel needs K in {1,2,4,8}");
}

}  // namespace

int upload_plan(Ctx* c, const ResPlan& plan) {
    c->res_cache.kind = 0;   // the descriptor buffer no longer holds a pair-kernel plan
    PHB_CUDA(c, cudaMemcpyAsync(c->d_res_rows, plan.rows.data(), plan.rows.size() * sizeof(ResRow),
                                cudaMemcpyHostToDevice, c->stream));
    PHB_CUDA(c, cudaStreamSynchronize(c->stream));   // plan.rows is a stack object
    return PHB_OK;
}

// store = write every node block to the partials array; with_root = append the root step and reduce the lnL
int dna_resident(Ctx* c, int root_a, int root_b, bool store, bool with_root) {
    ResPlan plan;
    int st = plan_rows(c, root_a, root_b, with_root, store, &plan);
    if (st) return st;
    if (plan.rows.empty()) return PHB_OK;
    st = upload_plan(c, plan);
    if (st) return st;
    int grid = 0;
    const int64_t n_wt = (c->S + 31) / 32;
    if (store && with_root) st = launch_resident_k<true, true>(c, plan, 0, n_wt, c->d_partial_sums, kPartialCap, &grid);
    else if (store) st = launch_resident_k<true, false>(c, plan, 0, n_wt, c->d_partial_sums, kPartialCap, &grid);
    else if (with_root) st = launch_resident_k<false, true>(c, plan, 0, n_wt, c->d_partial_sums, kPartialCap, &grid);
    else return c->fail(PHB_ERR_INVALID, "resident kernel: nothing to produce");
    if (st) return st;
    c->resident_slots = plan.n_slots;
    if (with_root) return launch_final_reduce(c, c->d_partial_sums, grid, 1, c->d_result);
    return PHB_OK;
}

// Whole evaluation starting from HOST tip codes: the pattern axis is cut into chunks; chunk i+1 is copied
// host->device on a second stream while the resident kernel walks chunk i (patterns are independent, so a
// chunk can be evaluated as soon as its codes have landed).  One synchronisation at the very end.
int dna_resident_from_host(Ctx* c, const uint8_t* codes_host, int n_chunks, int root_a, int root_b) {
    ResPlan plan;
    int st = plan_rows(c, root_a, root_b, true, false, &plan);
    if (st) return st;
    st = upload_plan(c, plan);
    if (st) return st;
    if (c->copy_stream == nullptr) {
        PHB_CUDA(c, cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
        for (int i = 0; i < kMaxChunks; ++i) PHB_CUDA(c, cudaEventCreateWithFlags(&c->chunk_events[i], cudaEventDisableTiming));
        PHB_CUDA(c, cudaEventCreateWithFlags(&c->start_event, cudaEventDisableTiming));
    }
    const int64_t n_wt = (c->S + 31) / 32;
    n_chunks = (int)std::max<int64_t>(1, std::min<int64_t>(std::min(n_chunks, kMaxChunks), n_wt));
    // the copy stream must not overtake work already queued on the compute stream (previous evaluation)
    PHB_CUDA(c, cudaEventRecord(c->start_event, c->stream));
    PHB_CUDA(c, cudaStreamWaitEvent(c->copy_stream, c->start_event, 0));
    int* d_flag = reinterpret_cast<int*>(c->d_result + 4 * kMaxEdgeBatch - 1);
    PHB_CUDA(c, cudaMemsetAsync(d_flag, 0, sizeof(int), c->stream));
    int parts = 0;
    for (int i = 0; i < n_chunks; ++i) {
        const int64_t b = n_wt * i / n_chunks, e = n_wt * (i + 1) / n_chunks;
        const int64_t s0 = b * 32, s1 = std::min<int64_t>(e * 32, c->S);
        PHB_CUDA(c, cudaMemcpy2DAsync(c->d_codes_ws + s0, c->code_pitch, codes_host + s0, (size_t)c->S, (size_t)(s1 - s0),
                                      (size_t)c->n_tips, cudaMemcpyHostToDevice, c->copy_stream));
        PHB_CUDA(c, cudaEventRecord(c->chunk_events[i], c->copy_stream));
        PHB_CUDA(c, cudaStreamWaitEvent(c->stream, c->chunk_events[i], 0));
        int grid = 0;
        st = launch_resident_k<false, true>(c, plan, b, e, c->d_partial_sums + parts, kPartialCap / n_chunks, &grid);
        if (st) return st;
        parts += grid;
    }
    c->d_codes = c->d_codes_ws;
    c->resident_slots = plan.n_slots;
    return launch_final_reduce(c, c->d_partial_sums, parts, 1, c->d_result);
}

}  // namespace phb
