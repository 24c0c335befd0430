// Operand-stack-resident pruning for 4-state models: the whole post-order walk of a pattern tile
// without touching HBM for operands.
//
// Same arithmetic as clv_dna.cu (reference `clv`, numba_likelihood_engine.py:10-46; root = tree_model.py:178-217),
// different data movement.  Patterns are independent, and a post-order walk is a stack machine: with the
// larger child subtree finished first, a 1000-taxon tree never has more than ~6 finished-but-unconsumed
// partial blocks alive (9 for a perfectly balanced 1024-taxon tree).  So:
//
//   * every WARP owns a tile of 32 patterns (lane = pattern, the K categories are looped inside the thread)
//     and walks ALL rows for it, independently of every other warp (no block barrier in the row loop);
//   * the result of a row stays in REGISTERS when the next row consumes it (2/3 of the internal operands),
//     otherwise it is parked in one of a few per-warp SHARED-MEMORY slots allotted by the host like a
//     register allocator;
//   * what does come from global memory is read-only and tiny - the row descriptor (16 B), the two P
//     blocks (2*K*128 B) and the tip codes of the tile (a few bytes) - and is brought in by cp.async into a
//     per-warp ring several rows ahead, so no load latency sits on the critical path;
//   * STORE = true additionally streams every finished block to HBM (partials[row][pattern][k][:], one
//     256-bit store per thread) for callers that need the per-node partials (TreeModel.partials,
//     derivatives); STORE = false is the pure lnL evaluation: tip codes in, per-pattern lnL out;
//   * ROOT = true appends the virtual-root step (root combine, pi-dot, Gamma mixture, log, weighted
//     sum) as a final pseudo-row, so one launch yields the per-pattern lnL and the block sums.
//
// One thread per pattern means the per-pattern maximum / exponent is thread-local (no shuffles), P rows are
// warp-wide broadcast reads, and all bookkeeping (descriptor, prefetch, exponents) is paid once per 32
// pattern-node updates.
#include <algorithm>
#include <cstdlib>

#include "common.cuh"

namespace phb {

namespace {

constexpr int kWarps = 4;          // warps per CTA (they only share the look-up table)
constexpr int kNS = 4;             // data ring depth (rows in flight)
constexpr int kND = 8;             // descriptor ring depth, >= 2*kNS - 1
constexpr int KIND_TIP = 0, KIND_PREV = 1, KIND_SLOT = 2;

// 16-byte row descriptor
struct __align__(16) ResRow {
    int32_t src_a;   // tip row or slot id
    int32_t src_b;
    int32_t pidx_a;  // P block of operand a
    uint32_t packed; // pidx_b [0:24) | kind_a [24:26) | kind_b [26:28) | dst slot [28:32) (15 = none)
};

struct ResArgs {
    const ResRow* rows;     // [n_rows (+1 root)]
    int n_steps;            // rows walked per tile (n_rows, +1 with ROOT)
    int n_rows;
    const double* pmats;
    const uint8_t* codes;
    size_t pitch;
    const double* lut;
    double* clv;            // STORE
    int32_t* scale;         // STORE
    const double* freqs;    // ROOT
    const double* catw;
    const double* weights;
    double* pattern_lnl;
    double* partial_sums;
    int64_t S;
    int n_slots;
    int warp_bytes;
};

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gmem_src) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(s), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void st256_stream(double* p, const double (&v)[4]) {
    asm volatile("st.global.cs.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(v[0]), "d"(v[1]), "d"(v[2]), "d"(v[3])
                 : "memory");
}

__device__ __forceinline__ void matvec4r(const double (&P)[16], const double (&v)[4], double (&out)[4]) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        double acc = P[4 * i] * v[0];
        acc = fma(P[4 * i + 1], v[1], acc);
        acc = fma(P[4 * i + 2], v[2], acc);
        acc = fma(P[4 * i + 3], v[3], acc);
        out[i] = acc;
    }
}

template <int K>
__device__ __forceinline__ int combine_scale(const double (&x)[4], const double (&y)[4], double (&o)[4]) {
#pragma unroll
    for (int i = 0; i < 4; ++i) o[i] = x[i] * y[i];
    const double m = fmax(fmax(o[0], o[1]), fmax(o[2], o[3]));
    int hi = __double2hiint(m);
#pragma unroll
    for (int d = K / 2; d > 0; d >>= 1) hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, d));
    int shift = 0;
    if (hi < kScaleThresholdHi && hi >= 0x00100000) {
        shift = 1023 - (hi >> 20);
        const double f = pow2i(shift);
#pragma unroll
        for (int i = 0; i < 4; ++i) o[i] *= f;
    }
    return -shift;
}

// geometry of one warp's shared memory.  A warp tile is 32 patterns: lane = pattern, the K categories
// are looped inside the thread.
template <int K>
struct WarpLayout {
    static constexpr int SPW = 32;                      // patterns per warp tile
    static constexpr int P_BYTES = 2 * K * 128;         // both P blocks, [operand][k][16 doubles]
    static constexpr int CODE_BYTES = 32;               // per operand
    static constexpr int STAGE_BYTES = P_BYTES + 2 * CODE_BYTES;
    static constexpr int DESC_BYTES = kND * 16;
    // a parked block: [k][half][lane] 16-byte pieces (every LDS.128/STS.128 touches 512 contiguous bytes)
    // followed by one exponent per lane
    static constexpr int SLOT_BYTES = K * 2 * 32 * 16 + 32 * 4;
    static constexpr int FIXED_BYTES = DESC_BYTES + kNS * STAGE_BYTES;
};

struct Cursor {
    int row;
    int64_t wt;
};

// o[k] = (P1[k] . a[k]) * (P2[k] . b[k]) for one pattern; operands per kind; result left in `prev`
template <int K, int KA, int KB>
__device__ __forceinline__ int row_update(const unsigned char* st, const unsigned char* s_slots, int src_a, int src_b,
                                          const double (*s_lut)[4], int lane, double (&prev)[K][4], int prev_e) {
    using L = WarpLayout<K>;
    double ta[4], tb[4];
    int e = 0;
    if (KA == KIND_TIP) {
        const int code = st[L::P_BYTES + lane];
        const double2 lo = *reinterpret_cast<const double2*>(&s_lut[code][0]);
        const double2 hi = *reinterpret_cast<const double2*>(&s_lut[code][2]);
        ta[0] = lo.x; ta[1] = lo.y; ta[2] = hi.x; ta[3] = hi.y;
    }
    if (KB == KIND_TIP) {
        const int code = st[L::P_BYTES + L::CODE_BYTES + lane];
        const double2 lo = *reinterpret_cast<const double2*>(&s_lut[code][0]);
        const double2 hi = *reinterpret_cast<const double2*>(&s_lut[code][2]);
        tb[0] = lo.x; tb[1] = lo.y; tb[2] = hi.x; tb[3] = hi.y;
    }
    if (KA == KIND_PREV || KB == KIND_PREV) e += prev_e;
    const unsigned char* sa = s_slots + (size_t)src_a * L::SLOT_BYTES;
    const unsigned char* sb = s_slots + (size_t)src_b * L::SLOT_BYTES;
    if (KA == KIND_SLOT) e += *reinterpret_cast<const int*>(sa + K * 1024 + lane * 4);
    if (KB == KIND_SLOT) e += *reinterpret_cast<const int*>(sb + K * 1024 + lane * 4);
    double m = 0.0;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        double a[4], b[4], x[4], y[4];
        if (KA == KIND_TIP) {
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = ta[i];
        } else if (KA == KIND_PREV) {
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = prev[k][i];
        } else {
            const double2 lo = *reinterpret_cast<const double2*>(sa + (k * 2 + 0) * 512 + lane * 16);
            const double2 hi = *reinterpret_cast<const double2*>(sa + (k * 2 + 1) * 512 + lane * 16);
            a[0] = lo.x; a[1] = lo.y; a[2] = hi.x; a[3] = hi.y;
        }
        if (KB == KIND_TIP) {
#pragma unroll
            for (int i = 0; i < 4; ++i) b[i] = tb[i];
        } else if (KB == KIND_PREV) {
#pragma unroll
            for (int i = 0; i < 4; ++i) b[i] = prev[k][i];
        } else {
            const double2 lo = *reinterpret_cast<const double2*>(sb + (k * 2 + 0) * 512 + lane * 16);
            const double2 hi = *reinterpret_cast<const double2*>(sb + (k * 2 + 1) * 512 + lane * 16);
            b[0] = lo.x; b[1] = lo.y; b[2] = hi.x; b[3] = hi.y;
        }
        // P rows are read as warp-wide broadcasts (every lane, same address)
        const double2* q1 = reinterpret_cast<const double2*>(st + k * 128);
        const double2* q2 = reinterpret_cast<const double2*>(st + K * 128 + k * 128);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const double2 r0 = q1[2 * i], r1 = q1[2 * i + 1];
            x[i] = fma(r1.y, a[3], fma(r1.x, a[2], fma(r0.y, a[1], r0.x * a[0])));
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const double2 r0 = q2[2 * i], r1 = q2[2 * i + 1];
            y[i] = fma(r1.y, b[3], fma(r1.x, b[2], fma(r0.y, b[1], r0.x * b[0])));
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            prev[k][i] = x[i] * y[i];
            m = fmax(m, prev[k][i]);
        }
    }
    const int hi = __double2hiint(m);
    if (hi < kScaleThresholdHi && hi >= 0x00100000) {   // 0 < max < 2^-128: rescale the whole pattern
        const int shift = 1023 - (hi >> 20);
        const double f = pow2i(shift);
#pragma unroll
        for (int k = 0; k < K; ++k)
#pragma unroll
            for (int i = 0; i < 4; ++i) prev[k][i] *= f;
        e -= shift;
    }
    return e;
}

template <int K, bool STORE, bool ROOT>
__global__ void __launch_bounds__(kWarps * 32) dna_resident_kernel(const ResArgs p) {
    using L = WarpLayout<K>;
    extern __shared__ __align__(128) unsigned char smem[];
    double(*s_lut)[4] = reinterpret_cast<double(*)[4]>(smem);
    __shared__ double s_red[kWarps];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 256 * 4; i += kWarps * 32) (&s_lut[0][0])[i] = p.lut[i];
    __syncthreads();

    unsigned char* wbase = smem + 256 * 32 + (size_t)warp * p.warp_bytes;
    ResRow* s_desc = reinterpret_cast<ResRow*>(wbase);
    unsigned char* s_stage = wbase + L::DESC_BYTES;
    unsigned char* s_slots = wbase + L::FIXED_BYTES;

    const int64_t n_wt = (p.S + L::SPW - 1) / L::SPW;
    const int64_t wstride = (int64_t)gridDim.x * kWarps;
    const int n_steps = p.n_steps;
    const size_t S = (size_t)p.S;

    Cursor cd{0, (int64_t)blockIdx.x * kWarps + warp};   // descriptor prefetch cursor
    Cursor cp = cd;                                      // data prefetch cursor
    Cursor cc = cd;                                      // compute cursor
    int qd = 0, qp = 0, qc = 0;

    auto advance = [&](Cursor& c) {
        if (++c.row == n_steps) {
            c.row = 0;
            c.wt += wstride;
        }
    };
    auto prefetch_desc = [&]() {
        if (cd.wt < n_wt && lane == 0) cp_async16(&s_desc[qd], &p.rows[cd.row]);
        qd = (qd + 1) % kND;
        advance(cd);
    };
    auto prefetch_data = [&]() {
        if (cp.wt < n_wt) {
            const ResRow d = s_desc[qp % kND];
            unsigned char* st = s_stage + (size_t)(qp % kNS) * L::STAGE_BYTES;
            const int pidx_b = d.packed & 0xffffff, kind_a = (d.packed >> 24) & 3, kind_b = (d.packed >> 26) & 3;
            const char* pa = reinterpret_cast<const char*>(p.pmats + (size_t)d.pidx_a * K * 16);
            const char* pb = reinterpret_cast<const char*>(p.pmats + (size_t)pidx_b * K * 16);
            constexpr int CH = K * 128 / 16;   // 16-byte chunks per P block
            for (int c = lane; c < 2 * CH; c += 32) {
                const char* src = c < CH ? pa + c * 16 : pb + (c - CH) * 16;
                cp_async16(st + c * 16, src);
            }
            const int64_t site0 = cp.wt * L::SPW;
            if (kind_a == KIND_TIP && lane < 2)
                cp_async16(st + L::P_BYTES + lane * 16, p.codes + (size_t)d.src_a * p.pitch + site0 + lane * 16);
            if (kind_b == KIND_TIP && lane >= 2 && lane < 4)
                cp_async16(st + L::P_BYTES + L::CODE_BYTES + (lane - 2) * 16,
                           p.codes + (size_t)d.src_b * p.pitch + site0 + (lane - 2) * 16);
        }
        ++qp;
        advance(cp);
    };

    for (int i = 0; i < kND - 1; ++i) prefetch_desc();
    cp_async_commit();
    cp_async_wait<0>();
    __syncwarp();
    for (int i = 0; i < kNS - 1; ++i) {
        prefetch_data();
        cp_async_commit();
    }

    double prev[K][4];
    int prev_e = 0;
#pragma unroll
    for (int k = 0; k < K; ++k)
#pragma unroll
        for (int i = 0; i < 4; ++i) prev[k][i] = 0.0;
    double acc = 0.0;

    while (cc.wt < n_wt) {
        cp_async_wait<kNS - 2>();
        __syncwarp();
        prefetch_desc();
        prefetch_data();
        cp_async_commit();

        const ResRow d = s_desc[qc % kND];
        const unsigned char* st = s_stage + (size_t)(qc % kNS) * L::STAGE_BYTES;
        const int kinds = (d.packed >> 24) & 15, dst_slot = d.packed >> 28;   // kind_a | kind_b << 2
        int e;
        switch (kinds) {
            case KIND_TIP | (KIND_TIP << 2):
                e = row_update<K, KIND_TIP, KIND_TIP>(st, s_slots, d.src_a, d.src_b, s_lut, lane, prev, prev_e);
                break;
            case KIND_TIP | (KIND_PREV << 2):
                e = row_update<K, KIND_TIP, KIND_PREV>(st, s_slots, d.src_a, d.src_b, s_lut, lane, prev, prev_e);
                break;
            case KIND_TIP | (KIND_SLOT << 2):
                e = row_update<K, KIND_TIP, KIND_SLOT>(st, s_slots, d.src_a, d.src_b, s_lut, lane, prev, prev_e);
                break;
            case KIND_PREV | (KIND_SLOT << 2):
                e = row_update<K, KIND_PREV, KIND_SLOT>(st, s_slots, d.src_a, d.src_b, s_lut, lane, prev, prev_e);
                break;
            case KIND_SLOT | (KIND_SLOT << 2):
                e = row_update<K, KIND_SLOT, KIND_SLOT>(st, s_slots, d.src_a, d.src_b, s_lut, lane, prev, prev_e);
                break;
            default:   // not a canonical row shape: the host plan is broken, do not touch memory
                e = 0;
                break;
        }
        prev_e = e;
        const int64_t s = cc.wt * L::SPW + lane;
        const bool is_root = ROOT && cc.row == n_steps - 1;
        if (!is_root) {
            if (dst_slot != 15) {
                unsigned char* sl = s_slots + (size_t)dst_slot * L::SLOT_BYTES;
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    *reinterpret_cast<double2*>(sl + (k * 2 + 0) * 512 + lane * 16) = make_double2(prev[k][0], prev[k][1]);
                    *reinterpret_cast<double2*>(sl + (k * 2 + 1) * 512 + lane * 16) = make_double2(prev[k][2], prev[k][3]);
                }
                *reinterpret_cast<int*>(sl + K * 1024 + lane * 4) = e;
            }
            if (STORE && s < p.S) {
                double* out = p.clv + ((size_t)cc.row * S + (size_t)s) * (K * 4);
#pragma unroll
                for (int k = 0; k < K; ++k) st256_stream(out + k * 4, prev[k]);
                p.scale[(size_t)cc.row * S + s] = e;
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
    cp_async_wait<0>();
    if (ROOT) {
        acc = warp_sum(acc);
        if (lane == 0) s_red[warp] = acc;
        __syncthreads();
        if (threadIdx.x == 0) {
            double t = 0;
#pragma unroll
            for (int w = 0; w < kWarps; ++w) t += s_red[w];
            p.partial_sums[blockIdx.x] = t;
        }
    }
}

// ---- host side: slot allocation + launch --------------------------------------------------------------------
struct ResPlan {
    std::vector<ResRow> rows;
    int n_slots = 0;
};

// Walk the schedule like a register allocator: a result consumed by the very next row stays in
// registers; anything else gets the lowest free slot until its consumer has read it.
int plan_rows(Ctx* c, int root_a, int root_b, bool with_root, ResPlan* out) {
    const int n_rows = c->n_rows();
    std::vector<int> slot_of_node(c->n_nodes, -1);
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
                src[i] = slot_of_node[nd];
                if (src[i] < 0) return c->fail(PHB_ERR_STATE, "resident plan: operand was never parked");
                busy[src[i]] = 0;   // free after this row has read it
            }
        }
        if (kind[0] > kind[1]) {   // canonical operand order TIP <= PREV <= SLOT (children commute): 5 row shapes
            std::swap(kind[0], kind[1]);
            std::swap(src[0], src[1]);
            std::swap(pidx[0], pidx[1]);
        }
        int dst = 15;
        if (dst_node >= 0 && consumer_row[dst_node] != r + 1 && consumer_row[dst_node] >= 0) {
            dst = grab();
            if (dst >= 15) return c->fail(PHB_ERR_UNSUPPORTED, "resident plan: tree needs more than 15 live blocks");
            slot_of_node[dst_node] = dst;
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
    out->n_slots = (int)busy.size();
    return PHB_OK;
}

template <int K, bool STORE, bool ROOT>
int launch_resident(Ctx* c, const ResPlan& plan, int* grid_out) {
    using L = WarpLayout<K>;
    ResArgs a;
    a.rows = static_cast<const ResRow*>(c->d_res_rows);
    a.n_rows = c->n_rows();
    a.n_steps = (int)plan.rows.size();
    a.pmats = c->d_pmats;
    a.codes = c->d_codes;
    a.pitch = c->code_pitch;
    a.lut = c->d_lut;
    a.clv = c->d_clv;
    a.scale = c->d_scale;
    a.freqs = c->model_freqs();
    a.catw = c->model_catw();
    a.weights = c->d_weights;
    a.pattern_lnl = c->d_pattern_lnl;
    a.partial_sums = c->d_partial_sums;
    a.S = c->S;
    a.n_slots = plan.n_slots;
    a.warp_bytes = (L::FIXED_BYTES + std::max(plan.n_slots, 1) * L::SLOT_BYTES + 127) / 128 * 128;
    const size_t smem = 256 * 32 + (size_t)kWarps * a.warp_bytes;
    if (smem > c->smem_optin)
        return c->fail(PHB_ERR_UNSUPPORTED, "resident kernel: operand stack does not fit in shared memory");
    auto kern = dna_resident_kernel<K, STORE, ROOT>;
    PHB_CUDA(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    PHB_CUDA(c, cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    int per_sm = 0;
    PHB_CUDA(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kWarps * 32, smem));
    if (per_sm < 1) per_sm = 1;
    const int64_t n_wt = (c->S + L::SPW - 1) / L::SPW;
    const int64_t blocks_needed = (n_wt + kWarps - 1) / kWarps;
    int64_t grid = std::min<int64_t>(blocks_needed, (int64_t)c->sm_count * per_sm);
    grid = std::min<int64_t>(grid, kMaxReduceBlocks);
    if (grid < 1) grid = 1;
    kern<<<(int)grid, kWarps * 32, smem, c->stream>>>(a);
    c->launches++;
    PHB_CUDA(c, cudaGetLastError());
    *grid_out = (int)grid;
    return PHB_OK;
}

template <bool STORE, bool ROOT>
int launch_resident_k(Ctx* c, const ResPlan& plan, int* grid_out) {
    switch (c->K) {
        case 1: return launch_resident<1, STORE, ROOT>(c, plan, grid_out);
        case 2: return launch_resident<2, STORE, ROOT>(c, plan, grid_out);
        case 4: return launch_resident<4, STORE, ROOT>(c, plan, grid_out);
        case 8: return launch_resident<8, STORE, ROOT>(c, plan, grid_out);
    }
    return c->fail(PHB_ERR_UNSUPPORTED, "resident kernel needs K in {1,2,4,8}");
}

}  // namespace

// mode: store = also write every node block; with_root = append the root step and reduce the lnL
int dna_resident(Ctx* c, int root_a, int root_b, bool store, bool with_root) {
    ResPlan plan;
    int st = plan_rows(c, root_a, root_b, with_root, &plan);
    if (st) return st;
    if (plan.rows.empty()) return PHB_OK;
    PHB_CUDA(c, cudaMemcpyAsync(c->d_res_rows, plan.rows.data(), plan.rows.size() * sizeof(ResRow),
                                cudaMemcpyHostToDevice, c->stream));
    PHB_CUDA(c, cudaStreamSynchronize(c->stream));   // plan.rows is a stack object
    int grid = 0;
    if (store && with_root) st = launch_resident_k<true, true>(c, plan, &grid);
    else if (store) st = launch_resident_k<true, false>(c, plan, &grid);
    else if (with_root) st = launch_resident_k<false, true>(c, plan, &grid);
    else return c->fail(PHB_ERR_INVALID, "resident kernel: nothing to produce");
    if (st) return st;
    c->resident_slots = plan.n_slots;
    if (with_root) return launch_final_reduce(c, c->d_partial_sums, grid, 1, c->d_result);
    return PHB_OK;
}

}  // namespace phb
