// Operand-stack-resident pruning for 4-state models: the whole post-order walk of a pattern tile
// without touching HBM for operands.
//
// Same arithmetic as clv_dna.cu (reference `clv`, numba_likelihood_engine.py:10-46; root = tree_model.py:178-217),
// different data movement.  Patterns are independent, and a post-order walk is a stack machine: with the
// larger child subtree finished first, a 1000-taxon tree never has more than ~6 finished-but-unconsumed
// partial blocks alive (9 for a perfectly balanced 1024-taxon tree).  So:
//
//   * every WARP owns a tile of (32/K)*U patterns and walks ALL rows for it, independently of every other
//     warp (no block barrier in the row loop);
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
// Thread mapping inside a warp is the one of clv_dna.cu: lane = (pattern g, category k), 4 doubles each.
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

// geometry of one warp's shared memory
template <int K, int U>
struct WarpLayout {
    static constexpr int SPI = 32 / K;                 // patterns per warp iteration
    static constexpr int SPW = SPI * U;                // patterns per warp tile
    static constexpr int P_BYTES = 2 * K * 128;        // both P blocks
    static constexpr int CODE_BYTES = SPW < 16 ? 16 : (SPW + 15) / 16 * 16;   // per operand, padded
    static constexpr int STAGE_BYTES = P_BYTES + 2 * CODE_BYTES;
    static constexpr int DESC_BYTES = kND * 16;
    static constexpr int SLOT_BYTES = U * 32 * 32 + U * 32 * 4;   // vectors + exponents
    static constexpr int FIXED_BYTES = DESC_BYTES + kNS * STAGE_BYTES;
};

struct Cursor {
    int row;
    int64_t wt;
};

template <int K, int U, bool STORE, bool ROOT>
__global__ void __launch_bounds__(kWarps * 32) dna_resident_kernel(const ResArgs p) {
    using L = WarpLayout<K, U>;
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

    const int g = lane / K, k = lane % K;
    const int64_t n_wt = (p.S + L::SPW - 1) / L::SPW;
    const int64_t wstride = (int64_t)gridDim.x * kWarps;
    const int n_steps = p.n_steps;
    const size_t S = (size_t)p.S;

    Cursor cd{0, (int64_t)blockIdx.x * kWarps + warp};   // descriptor prefetch cursor
    Cursor cp = cd;                                      // data prefetch cursor
    Cursor cc = cd;                                      // compute cursor
    int qd = 0, qp = 0, qc = 0;                          // ring positions (step counters mod ring size)

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
            for (int c = lane; c < CH; c += 32) {
                cp_async16(st + c * 16, pa + c * 16);
                cp_async16(st + K * 128 + c * 16, pb + c * 16);
            }
            const int64_t site0 = cp.wt * L::SPW;
            constexpr int CC = L::SPW < 16 ? 1 : L::CODE_BYTES / 16;
            if (kind_a == KIND_TIP && lane < CC) {
                const uint8_t* src = p.codes + (size_t)d.src_a * p.pitch + site0 + lane * 16;
                if (L::SPW >= 16) cp_async16(st + L::P_BYTES + lane * 16, src);
                else cp_async8(st + L::P_BYTES, src);
            }
            if (kind_b == KIND_TIP && lane >= 16 && lane < 16 + CC) {
                const uint8_t* src = p.codes + (size_t)d.src_b * p.pitch + site0 + (lane - 16) * 16;
                if (L::SPW >= 16) cp_async16(st + L::P_BYTES + L::CODE_BYTES + (lane - 16) * 16, src);
                else cp_async8(st + L::P_BYTES + L::CODE_BYTES, src);
            }
        }
        ++qp;
        advance(cp);
    };

    // ---- prologue: fill the descriptor ring, then the data ring ------------------------------------------
    for (int i = 0; i < kND - 1; ++i) prefetch_desc();
    cp_async_commit();
    cp_async_wait<0>();
    __syncwarp();
    for (int i = 0; i < kNS - 1; ++i) {
        prefetch_data();
        cp_async_commit();
    }

    double prev[U][4];
    int prev_e[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
        prev_e[u] = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) prev[u][i] = 0.0;
    }
    double acc = 0.0;

    while (cc.wt < n_wt) {
        cp_async_wait<kNS - 2>();
        __syncwarp();
        prefetch_desc();
        prefetch_data();
        cp_async_commit();

        const ResRow d = s_desc[qc % kND];
        const unsigned char* st = s_stage + (size_t)(qc % kNS) * L::STAGE_BYTES;
        const int kind_a = (d.packed >> 24) & 3, kind_b = (d.packed >> 26) & 3, dst_slot = d.packed >> 28;
        double P1[16], P2[16];
        {
            const double* q1 = reinterpret_cast<const double*>(st) + k * 16;
            const double* q2 = reinterpret_cast<const double*>(st + K * 128) + k * 16;
#pragma unroll
            for (int i = 0; i < 16; i += 2) {
                const double2 v1 = *reinterpret_cast<const double2*>(q1 + i);
                const double2 v2 = *reinterpret_cast<const double2*>(q2 + i);
                P1[i] = v1.x; P1[i + 1] = v1.y;
                P2[i] = v2.x; P2[i + 1] = v2.y;
            }
        }
        const uint8_t* codes_a = st + L::P_BYTES;
        const uint8_t* codes_b = st + L::P_BYTES + L::CODE_BYTES;
        const int64_t site0 = cc.wt * L::SPW;
        const bool is_root = ROOT && cc.row == n_steps - 1;

        auto operand = [&](int kind, int src, const uint8_t* codes, int u, double (&v)[4], int& e) {
            if (kind == KIND_TIP) {
                const int code = codes[u * L::SPI + g];
                const double2 lo = *reinterpret_cast<const double2*>(&s_lut[code][0]);
                const double2 hi = *reinterpret_cast<const double2*>(&s_lut[code][2]);
                v[0] = lo.x; v[1] = lo.y; v[2] = hi.x; v[3] = hi.y;
                e = 0;
            } else if (kind == KIND_PREV) {
#pragma unroll
                for (int i = 0; i < 4; ++i) v[i] = prev[u][i];
                e = prev_e[u];
            } else {
                const unsigned char* sl = s_slots + (size_t)src * L::SLOT_BYTES;
                const double2 lo = *reinterpret_cast<const double2*>(sl + (u * 32 + lane) * 32);
                const double2 hi = *reinterpret_cast<const double2*>(sl + (u * 32 + lane) * 32 + 16);
                v[0] = lo.x; v[1] = lo.y; v[2] = hi.x; v[3] = hi.y;
                e = *reinterpret_cast<const int*>(sl + U * 32 * 32 + (u * 32 + lane) * 4);
            }
        };

        if (!is_root) {
#pragma unroll
            for (int u = 0; u < U; ++u) {
                double a[4], b[4], x[4], y[4], o[4];
                int ea, eb;
                operand(kind_a, d.src_a, codes_a, u, a, ea);
                operand(kind_b, d.src_b, codes_b, u, b, eb);
                matvec4r(P1, a, x);
                matvec4r(P2, b, y);
                const int e = ea + eb + combine_scale<K>(x, y, o);
#pragma unroll
                for (int i = 0; i < 4; ++i) prev[u][i] = o[i];
                prev_e[u] = e;
                if (dst_slot != 15) {
                    unsigned char* sl = s_slots + (size_t)dst_slot * L::SLOT_BYTES;
                    *reinterpret_cast<double2*>(sl + (u * 32 + lane) * 32) = make_double2(o[0], o[1]);
                    *reinterpret_cast<double2*>(sl + (u * 32 + lane) * 32 + 16) = make_double2(o[2], o[3]);
                    *reinterpret_cast<int*>(sl + U * 32 * 32 + (u * 32 + lane) * 4) = e;
                }
                if (STORE) {
                    const int64_t s = site0 + u * L::SPI + g;
                    if (s < p.S) {
                        st256_stream(p.clv + (((size_t)cc.row * S + (size_t)s) * K + k) * 4, o);
                        if (k == 0) p.scale[(size_t)cc.row * S + s] = e;
                    }
                }
            }
        } else {
            double pi[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) pi[i] = p.freqs[i];
            const double wk = p.catw[k];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                double a[4], b[4], x[4], y[4], o[4];
                int ea, eb;
                operand(kind_a, d.src_a, codes_a, u, a, ea);
                operand(kind_b, d.src_b, codes_b, u, b, eb);
                matvec4r(P1, a, x);
                matvec4r(P2, b, y);
                const int e = ea + eb + combine_scale<K>(x, y, o);
                double f = pi[0] * o[0];
                f = fma(pi[1], o[1], f);
                f = fma(pi[2], o[2], f);
                f = fma(pi[3], o[3], f);
                double mix = f > 0 ? wk * f : 0.0;
#pragma unroll
                for (int dd = K / 2; dd > 0; dd >>= 1) mix += __shfl_xor_sync(0xffffffffu, mix, dd);
                const int64_t s = site0 + u * L::SPI + g;
                if (k == 0 && s < p.S) {
                    const double lnl = mix > 0 ? log(mix) + (double)e * kLn2 : -INFINITY;
                    p.pattern_lnl[s] = lnl;
                    acc += (p.weights ? p.weights[s] : 1.0) * lnl;
                }
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

template <int K, int U, bool STORE, bool ROOT>
int launch_resident(Ctx* c, const ResPlan& plan, int* grid_out) {
    using L = WarpLayout<K, U>;
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
    if (smem > c->smem_optin) return c->fail(PHB_ERR_UNSUPPORTED, "resident kernel: operand stack does not fit in shared memory");
    auto kern = dna_resident_kernel<K, U, STORE, ROOT>;
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

template <int K, bool STORE, bool ROOT>
int launch_resident_u(Ctx* c, const ResPlan& plan, int* grid_out) {
    // Larger U amortises the P reload of a row over more patterns, smaller U leaves room for more warps.
    // Take the largest U that still lets >= 8 warps per SM live with this tree's slot count and keeps the
    // chip busy.
    int force = c->resident_u;
    if (const char* env = getenv("PHB_RESIDENT_U")) force = atoi(env);
    auto fits = [&](int u, int sites_per_iter) {
        const int slot = u * 32 * 32 + u * 32 * 4;
        const int spw = sites_per_iter * u;
        const int code = spw < 16 ? 16 : (spw + 15) / 16 * 16;
        const size_t warp_bytes = kND * 16 + kNS * (2 * K * 128 + 2 * code) + (size_t)std::max(plan.n_slots, 1) * slot;
        const size_t cta = 256 * 32 + kWarps * warp_bytes;
        const bool enough_tiles = (c->S + spw - 1) / spw >= (int64_t)c->sm_count * 8;
        return cta * 2 <= c->smem_optin && enough_tiles;
    };
    const int spi = 32 / K;
    if (force == 4 || (force == 0 && fits(4, spi))) return launch_resident<K, 4, STORE, ROOT>(c, plan, grid_out);
    if (force == 2 || (force == 0 && (fits(2, spi) || spi * 1 < 8))) return launch_resident<K, 2, STORE, ROOT>(c, plan, grid_out);
    return launch_resident<K, 1, STORE, ROOT>(c, plan, grid_out);
}

template <bool STORE, bool ROOT>
int launch_resident_k(Ctx* c, const ResPlan& plan, int* grid_out) {
    switch (c->K) {
        case 1: return launch_resident_u<1, STORE, ROOT>(c, plan, grid_out);
        case 2: return launch_resident_u<2, STORE, ROOT>(c, plan, grid_out);
        case 4: return launch_resident_u<4, STORE, ROOT>(c, plan, grid_out);
        case 8: return launch_resident_u<8, STORE, ROOT>(c, plan, grid_out);
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
