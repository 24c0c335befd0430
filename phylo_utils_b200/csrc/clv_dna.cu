// Felsenstein pruning for 4-state (nucleotide) models - the headline path.
//
// What one schedule row computes is the reference's `clv` gufunc
// (/root/reference/phylo_utils/likelihood/numba_likelihood_engine.py:10-46):
//
//     parent[s][k][:] = (P1[k] . child1[s][k][:]) * (P2[k] . child2[s][k][:])
//
// with underflow protection.  The reference rescales per (pattern, category) by the category
// maximum and keeps natural-log scalers; here every pattern carries ONE cumulative binary
// exponent (int32): when the largest entry over all categories drops below 2^-128 the whole
// pattern is multiplied by the exact power of two that brings it back to [1,2).  Exact scaling
// adds no rounding error, so  stored * 2^exponent  equals the reference's  clv * exp(scaler)
// up to the reference's own division rounding.
//
// Data layout: partials[node][pattern][category][state] (fp64, state innermost, 32 B per
// (pattern, category)).  A thread owns one (pattern, category) pair: one 256-bit load per child,
// one 256-bit store per row, its category's two 4x4 P matrices in registers, the 4x4 . 4
// contraction in registers and the per-pattern maximum via width-K shuffles.  A warp therefore
// touches 1 KB of contiguous memory per load/store instruction.
//
// Two ways to walk the schedule:
//   * TILE  - persistent CTAs; a CTA owns a tile of patterns and walks ALL rows for it in one
//             launch.  Patterns are independent, so no grid-wide dependency exists.  The block a row
//             just wrote stays in shared memory and feeds the next row (SRC_PREV), and anything older
//             is re-read by the very thread that wrote it - usually from L2, because at any moment
//             the chip works on a window of patterns x one tree neighbourhood that fits in 126 MB.
//   * LEVEL - one launch per tree level, work item = (row of the level, pattern tile); this is for
//             alignments too short to fill 148 SMs with pattern tiles alone.
#include <cstdlib>

#include "common.cuh"

namespace phb {

namespace {

constexpr int kThreads = 128;

struct DnaArgs {
    const OpRow* rows;
    int row_begin, row_end;
    const double* pmats;    // [pidx][K][16]
    const uint8_t* codes;   // [tip][pitch]
    size_t pitch;
    const double* lut;      // [256][4]
    double* clv;            // [slot][S][K][4]
    int32_t* scale;         // [slot][S]
    int64_t S;
    int64_t n_tiles;
};

__device__ __forceinline__ void matvec4(const double (&P)[16], const double (&v)[4], double (&out)[4]) {
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
__device__ __forceinline__ int group_max_hi(int hi) {
#pragma unroll
    for (int o = K / 2; o > 0; o >>= 1) hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    return hi;
}

// product of the two child contributions, per-pattern rescale; returns the exponent to add
template <int K>
__device__ __forceinline__ int combine_and_scale(const double (&x)[4], const double (&y)[4], double (&o)[4]) {
#pragma unroll
    for (int i = 0; i < 4; ++i) o[i] = x[i] * y[i];
    const double m = fmax(fmax(o[0], o[1]), fmax(o[2], o[3]));
    const int hi = group_max_hi<K>(__double2hiint(m));
    int shift = 0;
    if (hi < kScaleThresholdHi && hi >= 0x00100000) {  // 0 < max < 2^-128 (normal numbers only)
        shift = 1023 - (hi >> 20);                      // brings the pattern maximum into [1, 2)
        const double f = pow2i(shift);
#pragma unroll
        for (int i = 0; i < 4; ++i) o[i] *= f;
    }
    return -shift;
}

template <int K, int U, int KA, int KB, bool LEVEL>
__device__ __forceinline__ void process_row(const DnaArgs& p, const OpRow& row, int64_t site0, double (*s_prev)[4],
                                            int* s_prev_e, const double (*s_lut)[4], int tid) {
    constexpr int SPI = kThreads / K;
    constexpr int UC = U < 4 ? U : 4;  // iterations whose loads are issued together
    const int g = tid / K, k = tid % K;

    double P1[16], P2[16];
    {
        const double* q1 = p.pmats + ((size_t)row.pidx[0] * K + k) * 16;
        const double* q2 = p.pmats + ((size_t)row.pidx[1] * K + k) * 16;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            ld256_nc(q1 + 4 * i, *reinterpret_cast<double(*)[4]>(&P1[4 * i]));
            ld256_nc(q2 + 4 * i, *reinterpret_cast<double(*)[4]>(&P2[4 * i]));
        }
    }
    const size_t S = (size_t)p.S;
    const double* ga = (KA == SRC_GLOBAL) ? p.clv + (size_t)row.src[0] * S * (K * 4) : nullptr;
    const double* gb = (KB == SRC_GLOBAL) ? p.clv + (size_t)row.src[1] * S * (K * 4) : nullptr;
    const int32_t* ea_ptr = (KA == SRC_GLOBAL) ? p.scale + (size_t)row.src[0] * S : nullptr;
    const int32_t* eb_ptr = (KB == SRC_GLOBAL) ? p.scale + (size_t)row.src[1] * S : nullptr;
    const uint8_t* ta = (KA == SRC_TIP) ? p.codes + (size_t)row.src[0] * p.pitch : nullptr;
    const uint8_t* tb = (KB == SRC_TIP) ? p.codes + (size_t)row.src[1] * p.pitch : nullptr;
    double* out = p.clv + (size_t)row.dst * S * (K * 4);
    int32_t* out_e = p.scale + (size_t)row.dst * S;

#pragma unroll
    for (int u0 = 0; u0 < U; u0 += UC) {
        double a[UC][4], b[UC][4];
        int ea[UC], eb[UC];
        // ---- issue every load of this chunk before any arithmetic -------------------------------
#pragma unroll
        for (int j = 0; j < UC; ++j) {
            const int u = u0 + j;
            const int64_t s = site0 + (int64_t)u * SPI + g;
            const bool ok = s < p.S;
            const size_t ss = ok ? (size_t)s : 0;
            if (KA == SRC_GLOBAL) {
                ld256(ga + (ss * K + k) * 4, a[j]);
                ea[j] = ea_ptr[ss];
            } else if (KA == SRC_TIP) {
                ea[j] = ta[ss];  // code, turned into a vector below
            }
            if (KB == SRC_GLOBAL) {
                ld256(gb + (ss * K + k) * 4, b[j]);
                eb[j] = eb_ptr[ss];
            } else if (KB == SRC_TIP) {
                eb[j] = tb[ss];
            }
        }
#pragma unroll
        for (int j = 0; j < UC; ++j) {
            const int u = u0 + j;
            const int64_t s = site0 + (int64_t)u * SPI + g;
            const bool ok = s < p.S;
            if (KA == SRC_TIP) {
                const int code = ea[j];
#pragma unroll
                for (int i = 0; i < 4; ++i) a[j][i] = s_lut[code][i];
                ea[j] = 0;
            } else if (KA == SRC_PREV) {
#pragma unroll
                for (int i = 0; i < 4; ++i) a[j][i] = s_prev[u * kThreads + tid][i];
                ea[j] = s_prev_e[u * kThreads + tid];
            }
            if (KB == SRC_TIP) {
                const int code = eb[j];
#pragma unroll
                for (int i = 0; i < 4; ++i) b[j][i] = s_lut[code][i];
                eb[j] = 0;
            } else if (KB == SRC_PREV) {
#pragma unroll
                for (int i = 0; i < 4; ++i) b[j][i] = s_prev[u * kThreads + tid][i];
                eb[j] = s_prev_e[u * kThreads + tid];
            }
            double x[4], y[4], o[4];
            matvec4(P1, a[j], x);
            matvec4(P2, b[j], y);
            const int e = ea[j] + eb[j] + combine_and_scale<K>(x, y, o);
            if (ok) {
                st256(out + ((size_t)s * K + k) * 4, o);
                if (k == 0) out_e[s] = e;
            }
            if (!LEVEL) {
#pragma unroll
                for (int i = 0; i < 4; ++i) s_prev[u * kThreads + tid][i] = o[i];
                s_prev_e[u * kThreads + tid] = e;
            }
        }
    }
}

template <int K, int U, bool LEVEL>
__device__ __forceinline__ void dispatch_row(const DnaArgs& p, const OpRow& row, int64_t site0, double (*s_prev)[4],
                                             int* s_prev_e, const double (*s_lut)[4], int tid) {
    int ka = row.kind[0], kb = row.kind[1];
    if (LEVEL) {
        if (ka == SRC_PREV) ka = SRC_GLOBAL;
        if (kb == SRC_PREV) kb = SRC_GLOBAL;
    }
    // host guarantees TIP <= PREV <= GLOBAL ordering of (ka, kb) in tile mode and TIP <= GLOBAL in level mode
    if (ka == SRC_TIP && kb == SRC_TIP)
        process_row<K, U, SRC_TIP, SRC_TIP, LEVEL>(p, row, site0, s_prev, s_prev_e, s_lut, tid);
    else if (ka == SRC_TIP && kb == SRC_GLOBAL)
        process_row<K, U, SRC_TIP, SRC_GLOBAL, LEVEL>(p, row, site0, s_prev, s_prev_e, s_lut, tid);
    else if (ka == SRC_GLOBAL && kb == SRC_GLOBAL)
        process_row<K, U, SRC_GLOBAL, SRC_GLOBAL, LEVEL>(p, row, site0, s_prev, s_prev_e, s_lut, tid);
    else if (!LEVEL && ka == SRC_TIP && kb == SRC_PREV)
        process_row<K, U, SRC_TIP, SRC_PREV, LEVEL>(p, row, site0, s_prev, s_prev_e, s_lut, tid);
    else if (!LEVEL && ka == SRC_PREV && kb == SRC_GLOBAL)
        process_row<K, U, SRC_PREV, SRC_GLOBAL, LEVEL>(p, row, site0, s_prev, s_prev_e, s_lut, tid);
}

template <int K, int U, bool LEVEL>
__global__ void __launch_bounds__(kThreads) dna_prune_kernel(const DnaArgs p) {
    constexpr int TS = (kThreads / K) * U;
    extern __shared__ __align__(32) unsigned char smem_raw[];
    double(*s_lut)[4] = reinterpret_cast<double(*)[4]>(smem_raw);                        // [256][4]
    double(*s_prev)[4] = reinterpret_cast<double(*)[4]>(smem_raw + 256 * 32);             // [U*128][4]
    int* s_prev_e = reinterpret_cast<int*>(smem_raw + 256 * 32 + (LEVEL ? 0 : U * kThreads * 32));
    const int tid = threadIdx.x;
    for (int i = tid; i < 256 * 4; i += kThreads) (&s_lut[0][0])[i] = p.lut[i];
    __syncthreads();

    if (LEVEL) {
        const int64_t items = (int64_t)(p.row_end - p.row_begin) * p.n_tiles;
        for (int64_t it = blockIdx.x; it < items; it += gridDim.x) {
            const int r = p.row_begin + (int)(it / p.n_tiles);
            const int64_t tile = it % p.n_tiles;
            const OpRow row = p.rows[r];
            dispatch_row<K, U, true>(p, row, tile * TS, s_prev, s_prev_e, s_lut, tid);
        }
    } else {
        for (int64_t tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
            for (int r = p.row_begin; r < p.row_end; ++r) {
                const OpRow row = p.rows[r];
                dispatch_row<K, U, false>(p, row, tile * TS, s_prev, s_prev_e, s_lut, tid);
            }
        }
    }
}

template <int K, int U, bool LEVEL>
int launch_prune(Ctx* c, const OpRow* d_rows, int row_begin, int row_end) {
    constexpr int TS = (kThreads / K) * U;
    DnaArgs a;
    a.rows = d_rows;
    a.row_begin = row_begin;
    a.row_end = row_end;
    a.pmats = c->d_pmats;
    a.codes = c->d_codes;
    a.pitch = c->code_pitch;
    a.lut = c->d_lut;
    a.clv = c->d_clv;
    a.scale = c->d_scale;
    a.S = c->S;
    a.n_tiles = (c->S + TS - 1) / TS;
    const size_t smem = 256 * 32 + (LEVEL ? 0 : (size_t)U * kThreads * (32 + 4));
    auto kern = dna_prune_kernel<K, U, LEVEL>;
    static bool configured[64] = {};
    if (!configured[c->device & 63]) {
        PHB_CUDA(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured[c->device & 63] = true;
    }
    int per_sm = 0;
    PHB_CUDA(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kThreads, smem));
    if (per_sm < 1) per_sm = 1;
    const int64_t items = LEVEL ? (int64_t)(row_end - row_begin) * a.n_tiles : a.n_tiles;
    const int64_t cap = (int64_t)c->sm_count * per_sm;
    const int grid = (int)(items < cap ? items : cap);
    if (grid <= 0) return PHB_OK;
    kern<<<grid, kThreads, smem, c->stream>>>(a);
    c->launches++;
    PHB_CUDA(c, cudaGetLastError());
    return PHB_OK;
}

// pick the unroll (tile size) so that there are enough tiles to occupy the chip
template <int K, bool LEVEL>
int launch_prune_u(Ctx* c, const OpRow* d_rows, int row_begin, int row_end, int64_t parallel_rows) {
    // One tile per SM is enough to keep the biggest tile: every row costs a thread 32 loads for its two P blocks,
    // which a big tile amortises over 8 elements (cfg5 shard, 62.5k patterns: up pass 26.4 ms at four tiles per SM,
    // 17.4 ms at one).  PHB_TILE_WANT overrides (tuning knob).
    const int64_t want = (int64_t)c->sm_count * (tuning().tile_want > 0 ? tuning().tile_want : 1);
    const int spi = kThreads / K;
    auto tiles = [&](int u) { return ((c->S + (int64_t)spi * u - 1) / ((int64_t)spi * u)) * parallel_rows; };
    if (tiles(8) >= want) return launch_prune<K, 8, LEVEL>(c, d_rows, row_begin, row_end);
    if (tiles(4) >= want) return launch_prune<K, 4, LEVEL>(c, d_rows, row_begin, row_end);
    if (tiles(2) >= want) return launch_prune<K, 2, LEVEL>(c, d_rows, row_begin, row_end);
    return launch_prune<K, 1, LEVEL>(c, d_rows, row_begin, row_end);
}

template <int K>
int run_rows_k(Ctx* c, const RowSet& rs, int mode) {
    const int n = rs.n_rows;
    if (n == 0) return PHB_OK;
    if (mode == PHB_MODE_LEVEL) {
        const std::vector<int32_t>& lv = *rs.levels;
        const int n_levels = (int)lv.size() - 1;
        for (int l = 0; l < n_levels; ++l) {
            const int b = lv[l], e = lv[l + 1];
            if (e <= b) continue;
            int st = launch_prune_u<K, true>(c, rs.d_rows, b, e, e - b);
            if (st != PHB_OK) return st;
        }
        return PHB_OK;
    }
    return launch_prune_u<K, false>(c, rs.d_rows, 0, n, 1);
}

// ---------------------------------------------------------------------------------------------------
// Root: virtual root on an edge, per-category site likelihoods, mixture, log, weighted sum.
// (tree_model.py:178-217: clv(P(0), P(len), ...) -> lnl_node -> logsumexp(. + log w) ; then sum)
// ---------------------------------------------------------------------------------------------------
struct DnaRootArgs {
    const double* pmats;  // [2][K][16]: P for child a, P for child b
    const uint8_t* codes;
    size_t pitch;
    const double* lut;
    const double* clv;
    const int32_t* scale;
    const double* freqs;
    const double* catw;
    const double* weights;  // [S] or null
    int src[2], kind[2];
    int64_t S;
    double* pattern_lnl;    // [S]
    double* cat_lnl;        // [S][K] or null
    double* root_clv;       // [S][K][4] or null
    int32_t* root_scale;    // [S] or null
    double* partial_sums;   // [gridDim.x]
};

template <int K>
__global__ void __launch_bounds__(kThreads) dna_root_kernel(const DnaRootArgs p) {
    constexpr int SPI = kThreads / K;
    __shared__ double s_lut[256][4];
    __shared__ double s_red[kThreads / 32];
    const int tid = threadIdx.x, g = tid / K, k = tid % K;
    for (int i = tid; i < 256 * 4; i += kThreads) (&s_lut[0][0])[i] = p.lut[i];
    __syncthreads();
    double Pa[16], Pb[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        Pa[i] = p.pmats[(size_t)k * 16 + i];
        Pb[i] = p.pmats[(size_t)(K + k) * 16 + i];
    }
    double pi[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) pi[i] = p.freqs[i];
    const double wk = p.catw[k];
    const size_t S = (size_t)p.S;
    double acc = 0.0;
    const int64_t n_iter = (p.S + SPI - 1) / SPI;
    for (int64_t it = blockIdx.x; it < n_iter; it += gridDim.x) {
        const int64_t s = it * SPI + g;
        const bool ok = s < p.S;
        const size_t ss = ok ? (size_t)s : 0;
        double a[4], b[4];
        int ea = 0, eb = 0;
        if (p.kind[0] == SRC_TIP) {
            const int code = p.codes[(size_t)p.src[0] * p.pitch + ss];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = s_lut[code][i];
        } else {
            ld256(p.clv + ((size_t)p.src[0] * S + ss) * (K * 4) + k * 4, a);
            ea = p.scale[(size_t)p.src[0] * S + ss];
        }
        if (p.kind[1] == SRC_TIP) {
            const int code = p.codes[(size_t)p.src[1] * p.pitch + ss];
#pragma unroll
            for (int i = 0; i < 4; ++i) b[i] = s_lut[code][i];
        } else {
            ld256(p.clv + ((size_t)p.src[1] * S + ss) * (K * 4) + k * 4, b);
            eb = p.scale[(size_t)p.src[1] * S + ss];
        }
        double x[4], y[4], o[4];
        matvec4(Pa, a, x);
        matvec4(Pb, b, y);
        const int e = ea + eb + combine_and_scale<K>(x, y, o);
        if (p.root_clv != nullptr && ok) {
            st256(p.root_clv + (ss * K + k) * 4, o);
            if (k == 0) p.root_scale[ss] = e;
        }
        double f = pi[0] * o[0];
        f = fma(pi[1], o[1], f);
        f = fma(pi[2], o[2], f);
        f = fma(pi[3], o[3], f);
        const double shift = (double)e * kLn2;
        if (p.cat_lnl != nullptr && ok) p.cat_lnl[ss * K + k] = f > 0 ? log(f) + shift : -INFINITY;
        double mix = f > 0 ? wk * f : 0.0;
#pragma unroll
        for (int o2 = K / 2; o2 > 0; o2 >>= 1) mix += __shfl_xor_sync(0xffffffffu, mix, o2);
        if (ok && k == 0) {
            const double lnl = mix > 0 ? log(mix) + shift : -INFINITY;
            p.pattern_lnl[ss] = lnl;
            acc += (p.weights ? p.weights[ss] : 1.0) * lnl;
        }
    }
    acc = warp_sum(acc);
    if ((tid & 31) == 0) s_red[tid >> 5] = acc;
    __syncthreads();
    if (tid == 0) {
        double t = 0;
#pragma unroll
        for (int w = 0; w < kThreads / 32; ++w) t += s_red[w];
        p.partial_sums[blockIdx.x] = t;
    }
}

template <int K>
int root_k(Ctx* c, int a, int b, bool want_cat, bool store_root) {
    DnaRootArgs p;
    p.pmats = c->d_pmats + (size_t)(2 * c->max_rows()) * c->K * 16;
    p.codes = c->d_codes;
    p.pitch = c->code_pitch;
    p.lut = c->d_lut;
    p.clv = c->d_clv;
    p.scale = c->d_scale;
    p.freqs = c->model_freqs();
    p.catw = c->model_catw();
    p.weights = c->d_weights;
    const int nodes[2] = {a, b};
    for (int i = 0; i < 2; ++i) {
        if (c->node_tip[nodes[i]] >= 0) {
            p.kind[i] = SRC_TIP;
            p.src[i] = c->node_tip[nodes[i]];
        } else {
            p.kind[i] = SRC_GLOBAL;
            p.src[i] = c->node_slot[nodes[i]];
        }
    }
    p.S = c->S;
    p.pattern_lnl = c->d_pattern_lnl;
    p.cat_lnl = want_cat ? c->d_cat_lnl : nullptr;
    p.root_clv = store_root ? c->d_root_clv : nullptr;
    p.root_scale = store_root ? c->d_root_scale : nullptr;
    p.partial_sums = c->d_partial_sums;
    const int64_t n_iter = (c->S + kThreads / K - 1) / (kThreads / K);
    int64_t grid = (int64_t)c->sm_count * 8;
    if (grid > n_iter) grid = n_iter;
    if (grid > kMaxReduceBlocks) grid = kMaxReduceBlocks;
    if (grid < 1) grid = 1;
    dna_root_kernel<K><<<(int)grid, kThreads, 0, c->stream>>>(p);
    c->launches++;
    PHB_CUDA(c, cudaGetLastError());
    return launch_final_reduce(c, c->d_partial_sums, (int)grid, 1, c->d_result);
}

}  // namespace

bool dna_supported(const Ctx* c) { return c->A == 4 && (c->K == 1 || c->K == 2 || c->K == 4 || c->K == 8); }

int dna_run_rows(Ctx* c, const RowSet& rs, int mode) {
    switch (c->K) {
        case 1: return run_rows_k<1>(c, rs, mode);
        case 2: return run_rows_k<2>(c, rs, mode);
        case 4: return run_rows_k<4>(c, rs, mode);
        case 8: return run_rows_k<8>(c, rs, mode);
    }
    return c->fail(PHB_ERR_UNSUPPORTED, "dna kernels need K in {1,2,4,8}");
}

int dna_root(Ctx* c, int a, int b, bool want_cat, bool store_root) {
    switch (c->K) {
        case 1: return root_k<1>(c, a, b, want_cat, store_root);
        case 2: return root_k<2>(c, a, b, want_cat, store_root);
        case 4: return root_k<4>(c, a, b, want_cat, store_root);
        case 8: return root_k<8>(c, a, b, want_cat, store_root);
    }
    return c->fail(PHB_ERR_UNSUPPORTED, "dna kernels need K in {1,2,4,8}");
}

}  // namespace phb
