// Pre-order ("up") partials and per-edge likelihood derivatives.
//
// The reference ships only the per-site, single-category primitive
// (lnl_branch_derivs, /root/reference/phylo_utils/likelihood/numba_likelihood_engine.py:49-57) and a
// re-rooting sweep table nobody consumes (utils.py:137-188): row [PAR,SIB,GPA,NOD,PAR] = "rebuild PAR's
// partial from SIB and GPA so that it faces NOD, then optimise edge NOD-PAR".  Here the same quantity -
// the partial of everything OUTSIDE NOD's subtree, seen from PAR's end of the edge - is computed for
// all nodes at once in a pre-order pass and kept next to the post-order partials:
//
//     up[c] = (P(len(p,sib)) . down[sib]) * (P(len(p,gpa)) . X),    X = up[p], or for a root child p
//                                                                    the other root child's down partial
//
// which is again a `clv` row, so the pass re-uses the pruning kernels (clv_dna.cu / clv_generic.cu) on a
// second row table whose destinations are blocks n_internal + node of the shared block array.  Valid for
// reversible models (pulley principle), exactly like the reference's re-rooting.
//
// Edge derivatives: for the edge above node c at trial length t, with a = down[c], b = up[c]:
//     f_k  = sum_i pi_i b_ki (P_k(t) a_k)_i,   f'_k, f''_k with dP/dt, d2P/dt2
//     L    = sum_k w_k f_k        (the per-pattern exponents of a and b are common to all k and cancel)
//     lnL  = sum_s wt_s [ log L_s + (e_a + e_b) ln 2 ]
//     dlnL = sum_s wt_s L'_s / L_s ,   d2lnL = sum_s wt_s [ L''_s / L_s - (L'_s / L_s)^2 ]
// i.e. lnl_branch_derivs composed over the Gamma mixture.
//
// Sum tables.  With P(t) = V e^(Lambda t) V^-1:  f_k^(d)(t) = sum_m g_d(lambda_m r_k) e^(lambda_m r_k t) s_km  with
//     s_km = (V^-1 a_k)_m (V^T (pi * b_k))_m
// which does not depend on t.  It is formed once per pre-order pass and kept in the edge's up block (whose up partial
// nothing else reads): by the pre-order walk itself for 4 states (up_dna_pair.cu), by the first derivative pass for
// 20 / 61 states (mma_edge_deriv_kernel's write-out).  Every further Newton iteration is a streaming dot product over
// ONE block per edge (dna_edge_st_kernel / edge_st_kernel) at HBM speed.  Ctx::up_sumtable / Ctx::st_ready record which
// edges have their table; a new pre-order pass starts over.
#include <algorithm>
#include <cstdlib>

#include "common.cuh"

namespace phb {

namespace {

// the two operands of an edge: a = partial below it, b = partial above it (block index or tip row + SRC_* kind)
struct EdgeDesc {
    int32_t src_a, kind_a, src_b, kind_b;
};

struct DerivArgs {
    const double* mats;     // [3][batch_cap][K][A][A]
    const uint8_t* codes;
    size_t pitch;
    const double* lut;
    const double* clv;      // shared block array (down blocks, then up blocks)
    const int32_t* scale;
    const double* freqs;
    const double* catw;
    const double* weights;
    int64_t S;
    int A, K, batch_cap, n_parts;
    const EdgeDesc* edges;  // [n_edges] operands at the two ends of every edge of the batch
    double* partial_sums;   // [n_edges * 3][n_parts]
};

constexpr int kDerivThreads = 128;
constexpr int kSitesPerThread = 4;

// grid = (n_parts, n_edges); thread = up to kSitesPerThread patterns; one category's three matrices in smem at a time
__global__ void __launch_bounds__(kDerivThreads) edge_deriv_kernel(const DerivArgs p) {
    extern __shared__ double sm[];
    const int A = p.A, K = p.K, e = blockIdx.y;
    double* M = sm;                       // [3][A][A]
    __shared__ double s_red[3][kDerivThreads / 32];
    const EdgeDesc ed = p.edges[e];
    const size_t S = (size_t)p.S;
    const int64_t span = (int64_t)kDerivThreads * kSitesPerThread;
    double tot[3] = {0.0, 0.0, 0.0};
    for (int64_t base = (int64_t)blockIdx.x * span; base < p.S; base += (int64_t)gridDim.x * span) {
        double acc[kSitesPerThread][3];
        const double* va[kSitesPerThread];
        const double* vb[kSitesPerThread];
        int ex[kSitesPerThread];
#pragma unroll
        for (int q = 0; q < kSitesPerThread; ++q) {
            acc[q][0] = acc[q][1] = acc[q][2] = 0.0;
            const int64_t s = base + (int64_t)q * kDerivThreads + threadIdx.x;
            const size_t ss = s < p.S ? (size_t)s : 0;
            ex[q] = 0;
            if (ed.kind_a == SRC_TIP) {
                va[q] = p.lut + (size_t)p.codes[(size_t)ed.src_a * p.pitch + ss] * A;
            } else {
                va[q] = p.clv + ((size_t)ed.src_a * S + ss) * K * A;
                ex[q] += p.scale[(size_t)ed.src_a * S + ss];
            }
            if (ed.kind_b == SRC_TIP) {
                vb[q] = p.lut + (size_t)p.codes[(size_t)ed.src_b * p.pitch + ss] * A;
            } else {
                vb[q] = p.clv + ((size_t)ed.src_b * S + ss) * K * A;
                ex[q] += p.scale[(size_t)ed.src_b * S + ss];
            }
        }
        for (int k = 0; k < K; ++k) {
            __syncthreads();
            for (int d = 0; d < 3; ++d) {
                const double* src = p.mats + (((size_t)d * p.batch_cap + e) * K + k) * A * A;
                for (int idx = threadIdx.x; idx < A * A; idx += kDerivThreads) M[d * A * A + idx] = src[idx];
            }
            __syncthreads();
            const double wk = p.catw[k];
#pragma unroll
            for (int q = 0; q < kSitesPerThread; ++q) {
                const double* a = va[q] + (ed.kind_a == SRC_TIP ? 0 : (size_t)k * A);
                const double* b = vb[q] + (ed.kind_b == SRC_TIP ? 0 : (size_t)k * A);
                double f0 = 0.0, f1 = 0.0, f2 = 0.0;
                for (int i = 0; i < A; ++i) {
                    double x0 = 0.0, x1 = 0.0, x2 = 0.0;
                    for (int j = 0; j < A; ++j) {
                        const double aj = a[j];
                        x0 = fma(M[i * A + j], aj, x0);
                        x1 = fma(M[A * A + i * A + j], aj, x1);
                        x2 = fma(M[2 * A * A + i * A + j], aj, x2);
                    }
                    const double pb = p.freqs[i] * b[i];
                    f0 = fma(pb, x0, f0);
                    f1 = fma(pb, x1, f1);
                    f2 = fma(pb, x2, f2);
                }
                acc[q][0] = fma(wk, f0, acc[q][0]);
                acc[q][1] = fma(wk, f1, acc[q][1]);
                acc[q][2] = fma(wk, f2, acc[q][2]);
            }
        }
#pragma unroll
        for (int q = 0; q < kSitesPerThread; ++q) {
            const int64_t s = base + (int64_t)q * kDerivThreads + threadIdx.x;
            if (s < p.S) {
                const double w = p.weights ? p.weights[s] : 1.0;
                const double L = acc[q][0];
                const double g = acc[q][1] / L;
                tot[0] += w * (L > 0 ? log(L) + (double)ex[q] * kLn2 : -INFINITY);
                tot[1] += w * g;
                tot[2] += w * (acc[q][2] / L - g * g);
            }
        }
    }
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        const double v = warp_sum(tot[d]);
        if ((threadIdx.x & 31) == 0) s_red[d][threadIdx.x >> 5] = v;
    }
    __syncthreads();
    if (threadIdx.x < 3) {
        double t = 0;
        for (int w = 0; w < kDerivThreads / 32; ++w) t += s_red[threadIdx.x][w];
        p.partial_sums[((size_t)e * 3 + threadIdx.x) * p.n_parts + blockIdx.x] = t;
    }
}

// 4-state specialisation: lane = (pattern, category) as in clv_dna.cu, one 256-bit load per end of the
// edge, the edge's three 4x4 matrices for this lane's category in registers, mixture by width-K shuffles.
// HBM-bound: 2 x K x 32 B per pattern and edge.
template <int K>
__global__ void __launch_bounds__(128) dna_edge_deriv_kernel(const DerivArgs p) {
    constexpr int SPI = 128 / K;
    __shared__ double s_lut[256][4];
    __shared__ double s_red[3][4];
    const int tid = threadIdx.x, g = tid / K, k = tid % K, e = blockIdx.y;
    for (int i = tid; i < 256 * 4; i += 128) (&s_lut[0][0])[i] = p.lut[i];
    __syncthreads();
    double M[3][16];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        const double* src = p.mats + (((size_t)d * p.batch_cap + e) * K + k) * 16;
#pragma unroll
        for (int i = 0; i < 16; ++i) M[d][i] = src[i];
    }
    double pi[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) pi[i] = p.freqs[i];
    const double wk = p.catw[k];
    const size_t S = (size_t)p.S;
    const EdgeDesc ed = p.edges[e];
    const int ka = ed.kind_a, kb = ed.kind_b;
    const size_t sa = (size_t)ed.src_a, sb = (size_t)ed.src_b;
    double tot[3] = {0.0, 0.0, 0.0};
    const int64_t n_iter = (p.S + SPI - 1) / SPI;
    // K consecutive pattern groups per trip.  After the width-K shuffles all K lanes of a group hold the same three
    // sums; lane k keeps those of round k, so that the expensive tail (log, two divisions) then runs ONCE per trip
    // with a different pattern in every lane instead of once per round in a quarter of the lanes.
    for (int64_t it0 = (int64_t)blockIdx.x * K; it0 < n_iter; it0 += (int64_t)gridDim.x * K) {
        double keep[3] = {1.0, 0.0, 0.0};
        int keep_ex = 0;
        size_t keep_s = 0;
        bool keep_ok = false;
#pragma unroll
        for (int u = 0; u < K; ++u) {
            const int64_t it = it0 + u;
            const int64_t s = it * SPI + g;
            const bool ok = it < n_iter && s < p.S;
            const size_t ss = ok ? (size_t)s : 0;
            double a[4], b[4];
            int ex = 0;
            if (ka == SRC_TIP) {
                const int code = p.codes[sa * p.pitch + ss];
#pragma unroll
                for (int i = 0; i < 4; ++i) a[i] = s_lut[code][i];
            } else {
                ld256(p.clv + ((sa * S + ss) * K + k) * 4, a);
                ex += p.scale[sa * S + ss];
            }
            if (kb == SRC_TIP) {
                const int code = p.codes[sb * p.pitch + ss];
#pragma unroll
                for (int i = 0; i < 4; ++i) b[i] = s_lut[code][i];
            } else {
                ld256(p.clv + ((sb * S + ss) * K + k) * 4, b);
                ex += p.scale[sb * S + ss];
            }
            double f[3];
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                double acc = 0.0;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    double x = M[d][4 * i] * a[0];
                    x = fma(M[d][4 * i + 1], a[1], x);
                    x = fma(M[d][4 * i + 2], a[2], x);
                    x = fma(M[d][4 * i + 3], a[3], x);
                    acc = fma(pi[i] * b[i], x, acc);
                }
                f[d] = wk * acc;
#pragma unroll
                for (int o = K / 2; o > 0; o >>= 1) f[d] += __shfl_xor_sync(0xffffffffu, f[d], o);
            }
            if (u == k) {
                keep[0] = f[0];
                keep[1] = f[1];
                keep[2] = f[2];
                keep_ex = ex;
                keep_s = ss;
                keep_ok = ok;
            }
        }
        if (keep_ok) {
            const double w = p.weights ? p.weights[keep_s] : 1.0;
            const double gq = keep[1] / keep[0];
            tot[0] += w * (keep[0] > 0 ? log(keep[0]) + (double)keep_ex * kLn2 : -INFINITY);
            tot[1] += w * gq;
            tot[2] += w * (keep[2] / keep[0] - gq * gq);
        }
    }
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        const double v = warp_sum(tot[d]);
        if ((tid & 31) == 0) s_red[d][tid >> 5] = v;
    }
    __syncthreads();
    if (tid < 3) p.partial_sums[((size_t)e * 3 + tid) * p.n_parts + blockIdx.x] = s_red[tid][0] + s_red[tid][1] + s_red[tid][2] + s_red[tid][3];
}

// ---- 20 and 61 states: edge derivatives on the FP64 tensor cores ------------------------------------------------
// With P_k(t) = V diag(exp(lambda r_k t)) V^-1 the three quantities of an edge share everything but a diagonal:
//     f_k^(d) = sum_m  g_d(lambda_m r_k) exp(lambda_m r_k t) . x_km . y_km,
//     x_k = V^-1 a_k,   y_k = V^T (pi * b_k)                       ("sum table" of the edge)
// so an edge costs two dense (A x A) . (A x N) products per category - the same DMMA tile loop as the pruning
// kernel (clv_mma.cu), with the two constant matrices staged ONCE per CTA - plus a 3A-term epilogue per pattern in
// fragment layout.  Nothing is written but the block sums.  The generic kernel below needs 3 A^2 shared-memory
// reads per pattern and category instead.
struct MmaDerivArgs {
    const double* m1;       // V^-1 [A][A] row-major
    const double* m2;       // [A][A]: m2[m][i] = V[i][m] pi[i]
    const double* coef;     // [edge][3][K][MROWS]: w_k g_d(lambda_m r_k) exp(lambda_m r_k t), zero for m >= A
    const uint8_t* codes;
    size_t pitch;
    const double* lut;
    const double* clv;
    const int32_t* scale;
    const double* weights;
    int64_t S, n_tiles;
    int K, n_parts;
    const EdgeDesc* edges;
    double* partial_sums;   // [n_edges * 3][n_parts]
    // sum-table write-out: an edge whose upper operand is an up block (block index >= st_first_block) gets its table
    // s_km = x_m y_m written over that block (and the summed exponents over its scalers); < 0: off
    int st_first_block;
    double* clv_rw;
    int32_t* scale_rw;
    // ... and the root edge (both of whose operands are down partials) into the otherwise unused up block of the root
    // child that owns it: up to two positions of the launch, -1 = none
    int st_extra_edge[2], st_extra_block[2];
};

__device__ __forceinline__ void dmma884(double (&d)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d[0]), "+d"(d[1])
                 : "d"(a), "d"(b));
}
constexpr int deriv_pad_pitch(int cols) {   // smallest pitch >= cols with pitch % 16 == 4: conflict-free 8 x 4 fragment loads
    int p = cols;
    while (p % 16 != 4) ++p;
    return p;
}
__device__ __forceinline__ void deriv_cp_async8(void* smem_dst, const void* gmem_src) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gmem_src) : "memory");
}

// coef[e][d][k][m]; grid = n_edges, one thread per (d, k, m)
__global__ void deriv_coef_kernel(const double* __restrict__ evals, const double* __restrict__ rates,
                                  const double* __restrict__ catw, const double* __restrict__ lengths, int A, int K,
                                  int mrows, int chain_rule, double* __restrict__ coef) {
    const int e = blockIdx.x;
    for (int idx = threadIdx.x; idx < 3 * K * mrows; idx += blockDim.x) {
        const int m = idx % mrows, k = (idx / mrows) % K, d = idx / (mrows * K);
        double v = 0.0;
        if (m < A) {
            const double lam = evals[m], r = rates[k];
            const double base = chain_rule ? lam * r : lam;
            const double g = d == 0 ? 1.0 : (d == 1 ? base : base * base);
            v = catw[k] * g * exp(lam * (lengths[e] * r));
        }
        coef[(size_t)e * 3 * K * mrows + idx] = v;
    }
}

// m2[m][i] = evecs[i][m] * freqs[i]
__global__ void deriv_m2_kernel(const double* __restrict__ evecs, const double* __restrict__ freqs, int A,
                                double* __restrict__ m2) {
    for (int idx = threadIdx.x; idx < A * A; idx += blockDim.x) {
        const int m = idx / A, i = idx - m * A;
        m2[idx] = evecs[i * A + m] * freqs[i];
    }
}

// grid = (n_parts, n_edges); a CTA walks pattern tiles of one edge, every warp owns NT * 8 patterns of the tile.
// DB: the operand rows of step (tile, category) + 1 stream into a second pair of row buffers while the products of the
// current step run (20 states; at 61 states one pair already fills the SM's shared memory).
template <int AA, int MT, int KS, int NT, int WARPS, bool DB>
__global__ void __launch_bounds__(WARPS * 32) mma_edge_deriv_kernel(const MmaDerivArgs p) {
    constexpr int A = AA, MROWS = MT * 8, KCOLS = KS * 4;
    constexpr int LDP = deriv_pad_pitch(KCOLS), LDL = deriv_pad_pitch(KCOLS);
    constexpr int TS = WARPS * NT * 8, WR = NT * 8, NBUF = DB ? 2 : 1;
    extern __shared__ double sm[];
    double* M1 = sm;                          // [MROWS][LDP]
    double* M2 = M1 + MROWS * LDP;            // [MROWS][LDP]
    double* La = M2 + MROWS * LDP;            // [NBUF][TS][LDL]   rows of the operand below the edge
    double* Lb = La + NBUF * TS * LDL;        // [NBUF][TS][LDL]   rows of the operand above the edge
    double* coef = Lb + NBUF * TS * LDL;      // [3][K][MROWS]
    __shared__ double s_red[3][WARPS];
    const int K = p.K, e = blockIdx.y;
    if (p.edges[e].kind_a == SRC_SUMTABLE) return;   // edge_st_kernel's edge (the whole CTA leaves)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int fr = lane >> 2, fc = lane & 3;
    const size_t S = (size_t)p.S;
    for (int i = threadIdx.x; i < 2 * MROWS * LDP + 2 * NBUF * TS * LDL; i += WARPS * 32) sm[i] = 0.0;
    __syncthreads();
    for (int i = threadIdx.x; i < A * A; i += WARPS * 32) {
        const int r = i / A, c = i - r * A;
        M1[r * LDP + c] = p.m1[i];
        M2[r * LDP + c] = p.m2[i];
    }
    for (int i = threadIdx.x; i < 3 * K * MROWS; i += WARPS * 32) coef[i] = p.coef[(size_t)e * 3 * K * MROWS + i];
    __syncthreads();
    const EdgeDesc ed = p.edges[e];
    const int ka = ed.kind_a, kb = ed.kind_b;
    const size_t sa = (size_t)ed.src_a, sb = (size_t)ed.src_b;
    // the upper operand is this edge's up block: leave the sum table there for the next iterations (edge_st_kernel)
    bool write_st = p.st_first_block >= 0 && kb == SRC_GLOBAL && ed.src_b >= p.st_first_block;
    size_t st_blk = sb;
    if (p.st_first_block >= 0 && (e == p.st_extra_edge[0] || e == p.st_extra_edge[1])) {
        write_st = true;
        st_blk = (size_t)(e == p.st_extra_edge[0] ? p.st_extra_block[0] : p.st_extra_block[1]);
    }
    // operand rows of this warp's patterns for one (tile, category) -> row buffer `buf`.  Rows of A doubles are 16-byte
    // aligned when A is even (A = 20: ten 16-byte pieces per row), 8-byte aligned otherwise (A = 61).  Tip rows do not
    // depend on the category: written when a tile meets a buffer for the first time.
    constexpr int PB = (A % 2 == 0) ? 16 : 8, PIECES = A * 8 / PB;
    auto stage = [&](int64_t tile, int k, int buf) {
        const int64_t wsite0 = tile * TS + (int64_t)warp * WR;
        const int n_valid = (int)max((int64_t)0, min((int64_t)WR, p.S - wsite0));
        auto one = [&](int kind, size_t src, double* mine) {
            if (kind == SRC_TIP) {
                if (k < NBUF)
                    for (int idx = lane; idx < n_valid * A; idx += 32) {
                        const int n = idx / A, j = idx - n * A;
                        mine[n * LDL + j] = __ldg(p.lut + (size_t)p.codes[src * p.pitch + wsite0 + n] * A + j);
                    }
                return;
            }
            const unsigned char* g = reinterpret_cast<const unsigned char*>(p.clv + ((src * S + wsite0) * K + k) * A);
            const unsigned sdst = (unsigned)__cvta_generic_to_shared(mine);
            for (int c = lane; c < n_valid * PIECES; c += 32) {
                const int n = c / PIECES, piece = c - n * PIECES;
                const unsigned d = sdst + n * (LDL * 8) + piece * PB;
                const unsigned char* q = g + (size_t)n * (K * A * 8) + piece * PB;
                if (PB == 16) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(q) : "memory");
                else asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(q) : "memory");
            }
        };
        one(ka, sa, La + ((size_t)buf * TS + (size_t)warp * WR) * LDL);
        one(kb, sb, Lb + ((size_t)buf * TS + (size_t)warp * WR) * LDL);
    };
    // the pattern whose tail (log, two divisions) this lane runs: all 32 lanes busy instead of four
    const int my_nt = fr >> 1, my_q = fr & 1;
    const int my_pi = my_nt * 8 + 2 * fc + my_q;
    double tot[3] = {0.0, 0.0, 0.0};
    double t[3][NT][2];
    double* st_ptr[NT][2];   // where component fr of this lane's patterns goes in the edge's sum table
    int my_ex = 0;
    int64_t tile = blockIdx.x;
    int k = 0, buf = 0;
    if (DB && tile < p.n_tiles) {
        stage(tile, 0, 0);
        asm volatile("cp.async.commit_group;" ::: "memory");
    }
    while (tile < p.n_tiles) {
        const int64_t wsite0 = tile * TS + (int64_t)warp * WR;
        int64_t tile_n = tile;
        int k_n = k + 1;
        if (k_n == K) {
            k_n = 0;
            tile_n += gridDim.x;
        }
        __syncwarp();   // every lane is done with the rows of the previous step
        if (DB) {
            if (tile_n < p.n_tiles) stage(tile_n, k_n, buf ^ 1);
            asm volatile("cp.async.commit_group;" ::: "memory");
            asm volatile("cp.async.wait_group 1;" ::: "memory");   // everything but the copies just issued has landed
        } else {
            stage(tile, k, 0);
            asm volatile("cp.async.commit_group;" ::: "memory");
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        __syncwarp();
        if (k == 0) {
#pragma unroll
            for (int d = 0; d < 3; ++d)
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) t[d][nt][0] = t[d][nt][1] = 0.0;
#pragma unroll
            for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const int64_t st_s = wsite0 + nt * 8 + 2 * fc + q;
                    st_ptr[nt][q] = (write_st && st_s < p.S) ? p.clv_rw + (st_blk * S + (size_t)st_s) * K * A + fr : nullptr;
                }
            // exponents of this lane's tail pattern: asked for now, needed K products later
            my_ex = 0;
            if (my_nt < NT && wsite0 + my_pi < p.S) {
                if (ka != SRC_TIP) my_ex += p.scale[sa * S + wsite0 + my_pi];
                if (kb != SRC_TIP) my_ex += p.scale[sb * S + wsite0 + my_pi];
            }
        }
        const double* myA = La + ((size_t)(DB ? buf : 0) * TS + (size_t)warp * WR) * LDL;
        const double* myB = Lb + ((size_t)(DB ? buf : 0) * TS + (size_t)warp * WR) * LDL;
        double x[MT][NT][2], y[MT][NT][2];
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) x[mt][nt][0] = x[mt][nt][1] = y[mt][nt][0] = y[mt][nt][1] = 0.0;
#pragma unroll 2
        for (int ks = 0; ks < KS; ++ks) {
            double fa[NT], fb[NT];
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
                fa[nt] = myA[(nt * 8 + fr) * LDL + ks * 4 + fc];
                fb[nt] = myB[(nt * 8 + fr) * LDL + ks * 4 + fc];
            }
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) {
                const double a1 = M1[(mt * 8 + fr) * LDP + ks * 4 + fc];
                const double a2 = M2[(mt * 8 + fr) * LDP + ks * 4 + fc];
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) {
                    dmma884(x[mt][nt], a1, fa[nt]);
                    dmma884(y[mt][nt], a2, fb[nt]);
                }
            }
        }
        // fragment element (mt, nt, q) = eigen-component mt*8+fr of pattern nt*8 + 2fc + q
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
            const double c0 = coef[(0 * K + k) * MROWS + mt * 8 + fr];
            const double c1 = coef[(1 * K + k) * MROWS + mt * 8 + fr];
            const double c2 = coef[(2 * K + k) * MROWS + mt * 8 + fr];
#pragma unroll
            for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const double xy = x[mt][nt][q] * y[mt][nt][q];
                    t[0][nt][q] = fma(c0, xy, t[0][nt][q]);
                    t[1][nt][q] = fma(c1, xy, t[1][nt][q]);
                    t[2][nt][q] = fma(c2, xy, t[2][nt][q]);
                    // eight lanes (fr) write eight consecutive components of one pattern: 64-byte runs.  The block's
                    // rows of this category have all been staged (cp.async waited above).  One pointer per pattern,
                    // set up once per tile (null: nothing to write) - the address arithmetic of 24 scattered stores
                    // per category was a third of the kernel's instructions.
                    if (st_ptr[nt][q] != nullptr && (mt * 8 + 8 <= A || mt * 8 + fr < A)) st_ptr[nt][q][k * A + mt * 8] = xy;
                }
        }
        if (k == K - 1) {
            // sum over the eigen-components held by the lanes that share fc; afterwards all eight of them hold the sums
            // of the eight patterns (nt, q) of that fc, and lane fr takes pattern (fr >> 1, fr & 1)
            double L = 1.0, t1 = 0.0, t2 = 0.0;
#pragma unroll
            for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    double v[3];
#pragma unroll
                    for (int d = 0; d < 3; ++d) {
                        v[d] = t[d][nt][q];
                        v[d] += __shfl_xor_sync(0xffffffffu, v[d], 4);
                        v[d] += __shfl_xor_sync(0xffffffffu, v[d], 8);
                        v[d] += __shfl_xor_sync(0xffffffffu, v[d], 16);
                    }
                    if (fr == nt * 2 + q) {
                        L = v[0];
                        t1 = v[1];
                        t2 = v[2];
                    }
                }
            const int64_t s = wsite0 + my_pi;
            if (my_nt < NT && s < p.S) {
                if (write_st) p.scale_rw[st_blk * S + s] = my_ex;   // the table's exponent: both ends'
                const double w = p.weights ? p.weights[s] : 1.0;
                const double g = t1 / L;
                tot[0] += w * (L > 0 ? log(L) + (double)my_ex * kLn2 : -INFINITY);
                tot[1] += w * g;
                tot[2] += w * (t2 / L - g * g);
            }
        }
        tile = tile_n;
        k = k_n;
        buf ^= 1;
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        const double v = warp_sum(tot[d]);
        if (lane == 0) s_red[d][warp] = v;
    }
    __syncthreads();
    if (threadIdx.x < 3) {
        double v = 0;
        for (int w = 0; w < WARPS; ++w) v += s_red[threadIdx.x][w];
        p.partial_sums[((size_t)e * 3 + threadIdx.x) * p.n_parts + blockIdx.x] = v;
    }
}

template <int AA, int MT, int KS, int NT, int WARPS, bool DB>
int launch_mma_derivs(Ctx* c, MmaDerivArgs& a, int n_edges) {
    constexpr int MROWS = MT * 8, KCOLS = KS * 4;
    constexpr int LDP = deriv_pad_pitch(KCOLS), LDL = deriv_pad_pitch(KCOLS);
    constexpr int TS = WARPS * NT * 8;
    a.n_tiles = (c->S + TS - 1) / TS;
    const size_t smem = (2 * (size_t)MROWS * LDP + (DB ? 4 : 2) * (size_t)TS * LDL + 3 * (size_t)c->K * MROWS) * sizeof(double);
    auto kern = mma_edge_deriv_kernel<AA, MT, KS, NT, WARPS, DB>;
    if (smem > c->smem_optin) return c->fail(PHB_ERR_UNSUPPORTED, "DMMA derivative kernel: does not fit in shared memory");
    PHB_CUDA(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    PHB_CUDA(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, WARPS * 32, smem));
    if (per_sm < 1) per_sm = 1;
    // enough CTAs per edge to fill the chip across the whole batch, few enough for the block-sum buffer
    // at least two CTAs per edge: with one, ~1000 edges are 3.4 waves of the chip and the last one runs a third full
    int64_t parts = std::max<int64_t>(2, ((int64_t)c->sm_count * per_sm + n_edges - 1) / n_edges);
    parts = std::min<int64_t>(parts, std::min<int64_t>(a.n_tiles, kPartialCap / (3 * n_edges)));
    a.n_parts = (int)parts;
    dim3 grid((unsigned)parts, (unsigned)n_edges);
    kern<<<grid, WARPS * 32, smem, c->stream>>>(a);
    c->launches++;
    PHB_CUDA(c, cudaGetLastError());
    return PHB_OK;
}

// Any state count, edges that already have their sum table (written by mma_edge_deriv_kernel on the first derivative
// pass after a pre-order pass): f^(d) = sum over j = (k, m) of coef[d][j] s[j] - three dot products of length K A over
// ONE contiguous row per pattern, no matrix product.  LP lanes share a pattern (row pieces of 16 bytes, coalesced),
// coefficients live in registers; every lane keeps the sums of one of 32 consecutive patterns, so the log / division
// tail runs once per 32 patterns with a distinct pattern in every lane.
struct StArgs {
    const double* coef;     // [edge][3][K][mrows]
    const double* clv;
    const int32_t* scale;
    const double* weights;
    const int32_t* edges;   // EdgeDesc array as ints: [4 e] = block of the table, [4 e + 1] = kind
    int64_t S;
    int K, A, mrows, n_parts;
    double* partial_sums;   // [n_edges * 3][n_parts]
};

template <int LP, int CPL>
__global__ void __launch_bounds__(128) edge_st_kernel(const StArgs p) {
    constexpr int PPW = 32 / LP;   // patterns per warp and sub-iteration
    __shared__ double s_red[3][4];
    const int e = blockIdx.y;
    if (p.edges[4 * e + 1] != SRC_SUMTABLE) return;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane / LP, l = lane % LP;
    const int row_doubles = p.K * p.A, row_chunks = row_doubles / 2;   // K A is even for every shape that gets here
    double cf[3][CPL][2];
#pragma unroll
    for (int i = 0; i < CPL; ++i)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int j = 2 * (l + LP * i) + h;
            const bool ok = j < row_doubles;
            const int k = ok ? j / p.A : 0, m = ok ? j - k * p.A : 0;
#pragma unroll
            for (int d = 0; d < 3; ++d) cf[d][i][h] = ok ? p.coef[(((size_t)e * 3 + d) * p.K + k) * p.mrows + m] : 0.0;
        }
    const size_t S = (size_t)p.S, blk = (size_t)p.edges[4 * e];
    const double* const base = p.clv + blk * S * row_doubles;
    const int32_t* const sc = p.scale + blk * S;
    double tot[3] = {0.0, 0.0, 0.0};
    // a warp takes 32 consecutive patterns per trip
    const int64_t n_trips = (p.S + 31) / 32;
    for (int64_t trip = (int64_t)blockIdx.x * 4 + warp; trip < n_trips; trip += (int64_t)gridDim.x * 4) {
        double keep[3] = {1.0, 0.0, 0.0};
        constexpr int UB = 4;   // sub-iterations whose loads are issued together
#pragma unroll 1
        for (int u0 = 0; u0 < LP; u0 += UB) {
            double2 v[UB][CPL];
#pragma unroll
            for (int b = 0; b < UB; ++b) {
                const int64_t s = trip * 32 + (u0 + b) * PPW + g;
                const double2* row = reinterpret_cast<const double2*>(base + (size_t)(s < p.S ? s : p.S - 1) * row_doubles);
#pragma unroll
                for (int i = 0; i < CPL; ++i) {
                    const int c = l + LP * i;
                    v[b][i] = c < row_chunks ? __ldg(row + c) : make_double2(0.0, 0.0);
                }
            }
#pragma unroll
            for (int b = 0; b < UB; ++b) {
                double f[3] = {0.0, 0.0, 0.0};
#pragma unroll
                for (int i = 0; i < CPL; ++i)
#pragma unroll
                    for (int d = 0; d < 3; ++d) f[d] = fma(cf[d][i][1], v[b][i].y, fma(cf[d][i][0], v[b][i].x, f[d]));
#pragma unroll
                for (int d = 0; d < 3; ++d)
#pragma unroll
                    for (int o = LP / 2; o > 0; o >>= 1) f[d] += __shfl_xor_sync(0xffffffffu, f[d], o);
                // lane u of group g keeps pattern u * PPW + g of the trip
                if (l == u0 + b) {
                    keep[0] = f[0];
                    keep[1] = f[1];
                    keep[2] = f[2];
                }
            }
        }
        const int64_t s = trip * 32 + l * PPW + g;
        if (s < p.S) {
            const double w = p.weights ? p.weights[s] : 1.0;
            const double gq = keep[1] / keep[0];
            tot[0] += w * (keep[0] > 0 ? log(keep[0]) + (double)sc[s] * kLn2 : -INFINITY);
            tot[1] += w * gq;
            tot[2] += w * (keep[2] / keep[0] - gq * gq);
        }
    }
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        const double v = warp_sum(tot[d]);
        if (lane == 0) s_red[d][warp] = v;
    }
    __syncthreads();
    if (tid < 3) p.partial_sums[((size_t)e * 3 + tid) * p.n_parts + blockIdx.x] = s_red[tid][0] + s_red[tid][1] + s_red[tid][2] + s_red[tid][3];
}

// ---- 4 states: the same sum-table form, scalar ------------------------------------------------------------------
// lane = (pattern, category) as in clv_dna.cu.  V^-1 and V^T diag(pi) travel in the kernel parameters, so the
// compiler feeds them to the FMAs straight from the constant bank: no registers, no loads.  A lane keeps only the
// twelve coefficients of its own category (w_k g_d(lambda_m r_k) exp(lambda_m r_k t)).  48 FP64 operations per
// (pattern, category) instead of 72, and a third of the registers of the matrix form above.
struct DnaSumArgs {
    double m1[16];          // V^-1, row-major
    double m2[16];          // m2[m][i] = V[i][m] pi[i]
    const double* coef;     // [edge][3][K][4]
    const uint8_t* codes;
    size_t pitch;
    const double* lut;
    const double* clv;
    const int32_t* scale;
    const double* weights;
    const EdgeDesc* edges;
    int64_t S;
    int n_parts;
    double* partial_sums;   // [n_edges * 3][n_parts]
    int n_list;             // > 0: grid.y walks list[] (the few edges of a sum-table pass that have no table: the root edge)
    int list[4];
};

template <int K>
__global__ void __launch_bounds__(128, 4) dna_edge_sumtable_kernel(const __grid_constant__ DnaSumArgs p) {
    constexpr int SPI = 128 / K;
    __shared__ double s_lut[256][4];
    __shared__ double s_red[3][4];
    const int tid = threadIdx.x, g = tid / K, k = tid % K, e = p.n_list > 0 ? p.list[blockIdx.y] : blockIdx.y;
    if (p.edges[e].kind_a == SRC_SUMTABLE) return;   // dna_edge_st_kernel's edge (the whole CTA leaves)
    for (int i = tid; i < 256 * 4; i += 128) (&s_lut[0][0])[i] = p.lut[i];
    __syncthreads();
    double cf[3][4];
#pragma unroll
    for (int d = 0; d < 3; ++d)
#pragma unroll
        for (int m = 0; m < 4; ++m) cf[d][m] = p.coef[(((size_t)e * 3 + d) * K + k) * 4 + m];
    const size_t S = (size_t)p.S;
    const EdgeDesc ed = p.edges[e];
    const int ka = ed.kind_a, kb = ed.kind_b;
    const size_t sa = (size_t)ed.src_a, sb = (size_t)ed.src_b;
    double tot[3] = {0.0, 0.0, 0.0};
    const int64_t n_iter = (p.S + SPI - 1) / SPI;
    // K consecutive pattern groups per trip; lane k keeps the sums of round k, so that the tail (log, two divisions)
    // runs once per trip with a different pattern in every lane
    for (int64_t it0 = (int64_t)blockIdx.x * K; it0 < n_iter; it0 += (int64_t)gridDim.x * K) {
        double keep[3] = {1.0, 0.0, 0.0};
        int keep_ex = 0;
        size_t keep_s = 0;
        bool keep_ok = false;
#pragma unroll
        for (int u = 0; u < K; ++u) {
            const int64_t it = it0 + u;
            const int64_t s = it * SPI + g;
            const bool ok = it < n_iter && s < p.S;
            const size_t ss = ok ? (size_t)s : 0;
            double a[4], b[4];
            int ex = 0;
            if (ka == SRC_TIP) {
                const int code = p.codes[sa * p.pitch + ss];
#pragma unroll
                for (int i = 0; i < 4; ++i) a[i] = s_lut[code][i];
            } else {
                ld256(p.clv + ((sa * S + ss) * K + k) * 4, a);
                ex += p.scale[sa * S + ss];
            }
            if (kb == SRC_TIP) {
                const int code = p.codes[sb * p.pitch + ss];
#pragma unroll
                for (int i = 0; i < 4; ++i) b[i] = s_lut[code][i];
            } else {
                ld256(p.clv + ((sb * S + ss) * K + k) * 4, b);
                ex += p.scale[sb * S + ss];
            }
            double f[3] = {0.0, 0.0, 0.0};
#pragma unroll
            for (int m = 0; m < 4; ++m) {
                double x = p.m1[4 * m] * a[0];
                x = fma(p.m1[4 * m + 1], a[1], x);
                x = fma(p.m1[4 * m + 2], a[2], x);
                x = fma(p.m1[4 * m + 3], a[3], x);
                double y = p.m2[4 * m] * b[0];
                y = fma(p.m2[4 * m + 1], b[1], y);
                y = fma(p.m2[4 * m + 2], b[2], y);
                y = fma(p.m2[4 * m + 3], b[3], y);
                const double xy = x * y;
                f[0] = fma(cf[0][m], xy, f[0]);
                f[1] = fma(cf[1][m], xy, f[1]);
                f[2] = fma(cf[2][m], xy, f[2]);
            }
#pragma unroll
            for (int d = 0; d < 3; ++d)
#pragma unroll
                for (int o = K / 2; o > 0; o >>= 1) f[d] += __shfl_xor_sync(0xffffffffu, f[d], o);
            if (u == k) {
                keep[0] = f[0];
                keep[1] = f[1];
                keep[2] = f[2];
                keep_ex = ex;
                keep_s = ss;
                keep_ok = ok;
            }
        }
        if (keep_ok) {
            const double w = p.weights ? p.weights[keep_s] : 1.0;
            const double gq = keep[1] / keep[0];
            tot[0] += w * (keep[0] > 0 ? log(keep[0]) + (double)keep_ex * kLn2 : -INFINITY);
            tot[1] += w * gq;
            tot[2] += w * (keep[2] / keep[0] - gq * gq);
        }
    }
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        const double v = warp_sum(tot[d]);
        if ((tid & 31) == 0) s_red[d][tid >> 5] = v;
    }
    __syncthreads();
    const size_t orow = p.n_list > 0 ? blockIdx.y : e;   // listed edges have their own block sums, indexed by list position
    if (tid < 3) p.partial_sums[(orow * 3 + tid) * p.n_parts + blockIdx.x] = s_red[tid][0] + s_red[tid][1] + s_red[tid][2] + s_red[tid][3];
}

// Edges whose up block holds the sum table s_km = (V^-1 a_k)_m (V^T (pi * b_k))_m left by the pre-order walk
// (up_dna_pair.cu): f_k^(d) = sum_m coef[d][k][m] s_km - ONE block read per edge, twelve FMAs per (pattern, category).
// lane = (pattern, category); 2K pattern groups per trip with all loads issued up front; lane k keeps the mixture
// sums of rounds k and k + K, so that the tail (log, two divisions) runs with a distinct pattern in every lane.
struct DnaStArgs {
    const double* coef;     // [edge][3][K][4]
    const double* clv;
    const int32_t* scale;
    const double* weights;
    const int32_t* blocks;  // the launch's EdgeDesc array as ints: [4 e] = block of the edge's sum table, [4 e + 1] = kind
    int64_t S;
    int n_parts;
    double* partial_sums;   // [n_edges * 3][n_parts]
};

template <int K>
__global__ void __launch_bounds__(128, 4) dna_edge_st_kernel(const DnaStArgs p) {
    constexpr int SPI = 128 / K, R = 2 * K;
    __shared__ double s_red[3][4];
    const int tid = threadIdx.x, g = tid / K, k = tid % K, e = blockIdx.y;
    double cf[3][4];
#pragma unroll
    for (int d = 0; d < 3; ++d)
#pragma unroll
        for (int m = 0; m < 4; ++m) cf[d][m] = p.coef[(((size_t)e * 3 + d) * K + k) * 4 + m];
    if (p.blocks[4 * e + 1] != SRC_SUMTABLE) return;   // EdgeDesc {src_a, kind_a, ...}: the root edge goes the general way
    const size_t S = (size_t)p.S, blk = (size_t)p.blocks[4 * e];
    const double* const base = p.clv + blk * S * (K * 4);
    const int32_t* const sc = p.scale + blk * S;
    double tot[3] = {0.0, 0.0, 0.0};
    const int64_t n_iter = (p.S + SPI - 1) / SPI;
    for (int64_t it0 = (int64_t)blockIdx.x * R; it0 < n_iter; it0 += (int64_t)gridDim.x * R) {
        double xy[R][4];
#pragma unroll
        for (int u = 0; u < R; ++u) {
            const int64_t s = (it0 + u) * SPI + g;
            ld256_nc(base + ((size_t)(s < p.S ? s : p.S - 1) * K + k) * 4, xy[u]);
        }
        double keep[2][3];
#pragma unroll
        for (int u = 0; u < R; ++u) {
            double f[3] = {0.0, 0.0, 0.0};
#pragma unroll
            for (int m = 0; m < 4; ++m) {
                f[0] = fma(cf[0][m], xy[u][m], f[0]);
                f[1] = fma(cf[1][m], xy[u][m], f[1]);
                f[2] = fma(cf[2][m], xy[u][m], f[2]);
            }
#pragma unroll
            for (int d = 0; d < 3; ++d)
#pragma unroll
                for (int o = K / 2; o > 0; o >>= 1) f[d] += __shfl_xor_sync(0xffffffffu, f[d], o);
            if (u % K == k) {
                keep[u / K][0] = f[0];
                keep[u / K][1] = f[1];
                keep[u / K][2] = f[2];
            }
        }
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int64_t s = (it0 + j * K + k) * SPI + g;
            if (s < p.S) {
                const double w = p.weights ? p.weights[s] : 1.0;
                const double gq = keep[j][1] / keep[j][0];
                tot[0] += w * (keep[j][0] > 0 ? log(keep[j][0]) + (double)sc[s] * kLn2 : -INFINITY);
                tot[1] += w * gq;
                tot[2] += w * (keep[j][2] / keep[j][0] - gq * gq);
            }
        }
    }
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        const double v = warp_sum(tot[d]);
        if ((tid & 31) == 0) s_red[d][tid >> 5] = v;
    }
    __syncthreads();
    if (tid < 3) p.partial_sums[((size_t)e * 3 + tid) * p.n_parts + blockIdx.x] = s_red[tid][0] + s_red[tid][1] + s_red[tid][2] + s_red[tid][3];
}

void fill_operand(const Ctx* c, int node, int* src, int* kind) {
    if (c->node_tip[node] >= 0) {
        *kind = SRC_TIP;
        *src = c->node_tip[node];
    } else {
        *kind = SRC_GLOBAL;
        *src = c->node_slot[node];
    }
}

}  // namespace

int launch_up_partials(Ctx* c, int node_a, int node_b) {
    const int n_rows = c->n_rows();
    c->up_rows.clear();
    c->up_levels.clear();
    c->up_sumtable = false;
    c->st_ready.assign(c->n_nodes, 0);
    if (n_rows == 0) return PHB_OK;
    // 4-state models whose post-order pass was the operand-resident walk: one pre-order walk (up_dna_pair.cu)
    if (c->resident_partials && !tuning().up_two_rows) {
        const int st = dna_up_walk(c, node_a, node_b);
        if (st != PHB_ERR_UNSUPPORTED) return st;
    }
    const int root_p = 2 * c->max_rows() + 1;  // P(root length), built by prepare_root
    std::vector<int> depth(c->n_nodes, 0), level_of;
    auto rank = [](int kind) { return kind == SRC_TIP ? 0 : 2; };
    // 20 / 61 states: one three-operand row per parent on the FP64 tensor cores (clv_mma.cu) - P.X once for both
    // children, three products instead of four
    if (mma_supported(c) && !tuning().disable_mma && !tuning().up_two_rows) {
        std::vector<int32_t> parents, plevel;
        for (int r = n_rows - 1; r >= 0; --r) {
            const int par = c->rows_raw[3 * r];
            int x_src, x_kind, x_pidx;
            if (par == node_a || par == node_b) {
                fill_operand(c, par == node_a ? node_b : node_a, &x_src, &x_kind);
                x_pidx = root_p;
                depth[par] = 0;
            } else {
                const int q = c->node_parent[par];
                PHB_REQUIRE(c, q >= 0, PHB_ERR_STATE, "up partials: the given root edge does not match the schedule");
                const int rq = c->node_row[q];
                x_src = c->n_internal + par;
                x_kind = SRC_GLOBAL;
                x_pidx = 2 * rq + (c->rows_raw[3 * rq + 1] == par ? 0 : 1);
                depth[par] = depth[q] + 1;
            }
            parents.insert(parents.end(), {x_src, x_kind, x_pidx});
            for (int i = 0; i < 2; ++i) {
                int src, kind;
                fill_operand(c, c->rows_raw[3 * r + 1 + i], &src, &kind);
                parents.insert(parents.end(), {src, kind, 2 * r + i});
            }
            parents.push_back(c->n_internal + c->rows_raw[3 * r + 1]);
            parents.push_back(c->n_internal + c->rows_raw[3 * r + 2]);
            plevel.push_back(depth[par]);
        }
        std::vector<int32_t> levels;
        if (!c->level_offsets.empty()) {   // level mode: parents grouped by depth (a parent's X is its parent's output)
            const int n = (int)plevel.size();
            std::vector<int> order(n);
            for (int i = 0; i < n; ++i) order[i] = i;
            std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return plevel[x] < plevel[y]; });
            std::vector<int32_t> sorted(parents.size());
            const int n_levels = plevel[order.back()] + 1;
            levels.assign(n_levels + 1, 0);
            for (int i = 0; i < n; ++i) {
                std::copy(parents.begin() + (size_t)order[i] * 11, parents.begin() + (size_t)order[i] * 11 + 11, sorted.begin() + (size_t)i * 11);
                levels[plevel[order[i]] + 1]++;
            }
            for (int l = 0; l < n_levels; ++l) levels[l + 1] += levels[l];
            parents.swap(sorted);
        }
        return mma_run_parent_rows(c, parents, levels);
    }
    for (int r = n_rows - 1; r >= 0; --r) {
        const int par = c->rows_raw[3 * r];
        const int ch[2] = {c->rows_raw[3 * r + 1], c->rows_raw[3 * r + 2]};
        // X: what sits "above" par
        int x_src, x_kind, x_pidx;
        if (par == node_a || par == node_b) {
            fill_operand(c, par == node_a ? node_b : node_a, &x_src, &x_kind);
            x_pidx = root_p;
            depth[par] = 0;
        } else {
            const int q = c->node_parent[par];
            PHB_REQUIRE(c, q >= 0, PHB_ERR_STATE, "up partials: the given root edge does not match the schedule");
            const int rq = c->node_row[q];
            x_src = c->n_internal + par;
            x_kind = SRC_GLOBAL;
            x_pidx = 2 * rq + (c->rows_raw[3 * rq + 1] == par ? 0 : 1);
        }
        for (int i = 0; i < 2; ++i) {
            const int child = ch[i], sib = ch[1 - i];
            OpRow row{};
            row.dst = c->n_internal + child;
            fill_operand(c, sib, &row.src[0], &row.kind[0]);
            row.pidx[0] = 2 * r + (1 - i);
            row.src[1] = x_src;
            row.kind[1] = x_kind;
            row.pidx[1] = x_pidx;
            if (rank(row.kind[0]) > rank(row.kind[1])) {
                std::swap(row.kind[0], row.kind[1]);
                std::swap(row.src[0], row.src[1]);
                std::swap(row.pidx[0], row.pidx[1]);
            }
            depth[child] = depth[par] + 1;
            c->up_rows.push_back(row);
            level_of.push_back(depth[par]);
        }
    }
    const int n_up = (int)c->up_rows.size();
    int mode = c->level_offsets.empty() ? PHB_MODE_TILE : PHB_MODE_LEVEL;
    if (mode == PHB_MODE_LEVEL) {
        std::vector<int> order(n_up);
        for (int i = 0; i < n_up; ++i) order[i] = i;
        std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return level_of[x] < level_of[y]; });
        std::vector<OpRow> sorted(n_up);
        const int n_levels = level_of[order.back()] + 1;
        c->up_levels.assign(n_levels + 1, 0);
        for (int i = 0; i < n_up; ++i) {
            sorted[i] = c->up_rows[order[i]];
            c->up_levels[level_of[order[i]] + 1]++;
        }
        for (int l = 0; l < n_levels; ++l) c->up_levels[l + 1] += c->up_levels[l];
        c->up_rows.swap(sorted);
    }
    PHB_CUDA(c, cudaMemcpyAsync(c->d_up_rows, c->up_rows.data(), (size_t)n_up * sizeof(OpRow), cudaMemcpyHostToDevice,
                                c->stream));
    PHB_CUDA(c, cudaStreamSynchronize(c->stream));
    const RowSet rs{c->d_up_rows, n_up, &c->up_levels};
    return run_rows(c, rs, mode);
}

// far_nodes == nullptr: edge i is the one ABOVE nodes[i] (its other end is the node's up block, or the other root
// child).  far_nodes != nullptr: edge i connects the down-array partials of nodes[i] and far_nodes[i] as they are now -
// the re-rooting sweep, where every partial faces the edge being optimised (phb_branch_derivatives).
int launch_edge_derivatives(Ctx* c, int n_edges, const int32_t* nodes, const double* lengths, int chain_rule,
                            double* out, const int32_t* far_nodes) {
    const int A = c->A, K = c->K;
    const size_t blk = (size_t)K * A * A;
    const bool use_mma = mma_supported(c) && !tuning().disable_mma;
    const int mrows = A == 20 ? 24 : 64;
    // edges per launch: as many as the matrix scratch area (or, for the sum-table kernel, the coefficient table), the
    // block-sum buffer and the grid's y dimension hold - all 2N-3 edges of a 4-state tree go out in one launch
    int cap = use_mma ? (int)std::min<size_t>((c->dmats_doubles - (size_t)A * A) / (3 * (size_t)K * mrows), 1 << 20)
                      : (int)std::min<size_t>(c->dmats_doubles / (3 * blk), 1 << 20);
    // the 4-state sum-table pass keeps the last kListArea doubles of the block-sum buffer for the root-edge list
    constexpr int kListParts = 512, kListArea = 4 * 3 * kListParts;
    cap = std::min(std::min(cap, c->n_nodes), std::min((kPartialCap - kListArea) / 3, 65535));
    if (cap < 1) return c->fail(PHB_ERR_NOMEM, "edge derivatives: scratch area too small");
    double* d_len = c->d_lengths + 2 * (size_t)c->max_rows() + 2;
    std::vector<EdgeDesc> edges;
    for (int start = 0; start < n_edges; start += cap) {
        const int n = std::min(cap, n_edges - start);
        edges.assign(n, EdgeDesc{});
        for (int i = 0; i < n; ++i) {
            const int node = nodes[start + i];
            PHB_REQUIRE(c, node >= 0 && node < c->n_nodes, PHB_ERR_INVALID, "edge derivatives: node id out of range");
            PHB_REQUIRE(c, lengths[start + i] >= 0, PHB_ERR_INVALID, "edge derivatives: negative branch length");
            fill_operand(c, node, &edges[i].src_a, &edges[i].kind_a);
            if (far_nodes != nullptr) {
                const int far = far_nodes[start + i];
                PHB_REQUIRE(c, far >= 0 && far < c->n_nodes && far != node, PHB_ERR_INVALID, "branch derivatives: bad far node");
                fill_operand(c, far, &edges[i].src_b, &edges[i].kind_b);
            } else if (node == c->root_a || node == c->root_b) {
                fill_operand(c, node == c->root_a ? c->root_b : c->root_a, &edges[i].src_b, &edges[i].kind_b);
                if (use_mma && (int)c->st_ready.size() == c->n_nodes && c->st_ready[node]) {
                    edges[i].src_a = c->n_internal + node;   // the root edge's table sits in this root child's up block
                    edges[i].kind_a = SRC_SUMTABLE;
                }
            } else {
                PHB_REQUIRE(c, c->node_parent[node] >= 0, PHB_ERR_INVALID, "edge derivatives: node has no edge above it");
                edges[i].src_b = c->n_internal + node;
                edges[i].kind_b = SRC_GLOBAL;
                if (c->up_sumtable || (use_mma && (int)c->st_ready.size() == c->n_nodes && c->st_ready[node])) {
                    edges[i].src_a = c->n_internal + node;   // the edge's up block holds its sum table
                    edges[i].kind_a = SRC_SUMTABLE;
                }
            }
        }
        PHB_CUDA(c, cudaMemcpyAsync(d_len, lengths + start, (size_t)n * 8, cudaMemcpyHostToDevice, c->stream));
        // pageable sources: both copies are staged before the calls return, so `edges` may be refilled by the next batch
        PHB_CUDA(c, cudaMemcpyAsync(c->d_edges, edges.data(), (size_t)n * sizeof(EdgeDesc), cudaMemcpyHostToDevice, c->stream));
        const EdgeDesc* d_edges = static_cast<const EdgeDesc*>(c->d_edges);
        int n_parts = 0;
        int n_list_done = 0, list_done[4] = {0, 0, 0, 0}, list_parts_done = 0;   // edges reduced from their own block sums
        if (use_mma) {
            // sum-table form on the FP64 tensor cores; the matrix scratch area holds the coefficients and V^T diag(pi)
            MmaDerivArgs m;
            double* d_coef = c->d_dmats;
            double* d_m2 = c->d_dmats + (size_t)cap * 3 * K * mrows;
            deriv_coef_kernel<<<n, 256, 0, c->stream>>>(c->model_evals(), c->model_rates(), c->model_catw(), d_len, A, K, mrows,
                                                        chain_rule, d_coef);
            deriv_m2_kernel<<<1, 256, 0, c->stream>>>(c->model_evecs(), c->model_freqs(), A, d_m2);
            c->launches += 2;
            PHB_CUDA(c, cudaGetLastError());
            m.m1 = c->model_ivecs();
            m.m2 = d_m2;
            m.coef = d_coef;
            m.codes = c->d_codes;
            m.pitch = c->code_pitch;
            m.lut = c->d_lut;
            m.clv = c->d_clv;
            m.scale = c->d_scale;
            m.weights = c->d_weights;
            m.S = c->S;
            m.K = K;
            m.edges = d_edges;
            m.partial_sums = c->d_partial_sums;
            // First derivative pass after a pre-order pass: the kernel leaves each edge's sum table in its up block;
            // later passes read that one block per edge (edge_st_kernel).  Needs 16-byte rows (K A even), a row that
            // fits the register-resident coefficient layouts, and every node at most once per launch.
            const int row_chunks = K * A / 2;
            bool st_ok = far_nodes == nullptr && !tuning().deriv_no_st && (K * A) % 2 == 0 && row_chunks <= 256 &&
                         (int)c->st_ready.size() == c->n_nodes;
            int n_table = 0, n_fresh = 0;
            if (st_ok) {
                std::vector<char> seen(c->n_nodes, 0);
                for (int i = 0; i < n && st_ok; ++i) {
                    const int node = nodes[start + i];
                    if (seen[node]) st_ok = false;
                    seen[node] = 1;
                }
            }
            for (int i = 0; i < n; ++i) {
                if (edges[i].kind_a == SRC_SUMTABLE) ++n_table;
                else ++n_fresh;
            }
            m.st_first_block = st_ok ? c->n_internal : -1;
            m.clv_rw = c->d_clv;
            m.scale_rw = c->d_scale;
            m.st_extra_edge[0] = m.st_extra_edge[1] = -1;
            m.st_extra_block[0] = m.st_extra_block[1] = 0;
            if (st_ok)
                for (int i = 0, x = 0; i < n; ++i) {
                    const int node = nodes[start + i];
                    if ((node == c->root_a || node == c->root_b) && edges[i].kind_a != SRC_SUMTABLE && x < 2) {
                        m.st_extra_edge[x] = i;
                        m.st_extra_block[x] = c->n_internal + node;
                        ++x;
                    }
                }
            int st = PHB_OK;
            if (n_fresh > 0) {
                st = A == 20 ? launch_mma_derivs<20, 3, 5, 4, 4, true>(c, m, n) : launch_mma_derivs<61, 8, 16, 2, 8, false>(c, m, n);
                if (st) return st;
            } else {
                // block sums are laid out for this many parts either way
                int64_t parts = std::max<int64_t>(2, ((int64_t)c->sm_count * 4 + n - 1) / n);
                m.n_parts = (int)std::min<int64_t>(parts, std::min<int64_t>((c->S + 127) / 128, kPartialCap / (3 * n)));
            }
            if (n_table > 0) {
                StArgs t;
                t.coef = d_coef;
                t.clv = c->d_clv;
                t.scale = c->d_scale;
                t.weights = c->d_weights;
                t.edges = reinterpret_cast<const int32_t*>(d_edges);
                t.S = c->S;
                t.K = K;
                t.A = A;
                t.mrows = mrows;
                t.n_parts = m.n_parts;
                t.partial_sums = c->d_partial_sums;
                dim3 grid((unsigned)m.n_parts, (unsigned)n);
                if (row_chunks <= 40) edge_st_kernel<8, 5><<<grid, 128, 0, c->stream>>>(t);
                else if (row_chunks <= 128) edge_st_kernel<32, 4><<<grid, 128, 0, c->stream>>>(t);
                else edge_st_kernel<32, 8><<<grid, 128, 0, c->stream>>>(t);
                c->launches++;
                PHB_CUDA(c, cudaGetLastError());
            }
            if (st_ok)
                for (int i = 0; i < n; ++i)
                    if (edges[i].kind_a != SRC_SUMTABLE &&
                        ((edges[i].kind_b == SRC_GLOBAL && edges[i].src_b >= c->n_internal) || i == m.st_extra_edge[0] ||
                         i == m.st_extra_edge[1]))
                        c->st_ready[nodes[start + i]] = 1;
            n_parts = m.n_parts;
        } else if (dna_supported(c) && (int)c->h_evecs.size() == A * A && (int)c->h_freqs.size() == A &&
                   (c->up_sumtable || !tuning().deriv_matrix_form)) {
            DnaSumArgs q;
            for (int m = 0; m < 4; ++m)
                for (int i = 0; i < 4; ++i) {
                    q.m1[4 * m + i] = c->h_ivecs[4 * m + i];
                    q.m2[4 * m + i] = c->h_evecs[4 * i + m] * c->h_freqs[i];
                }
            double* d_coef = c->d_dmats;
            deriv_coef_kernel<<<n, 64, 0, c->stream>>>(c->model_evals(), c->model_rates(), c->model_catw(), d_len, A, K, 4,
                                                       chain_rule, d_coef);
            c->launches++;
            PHB_CUDA(c, cudaGetLastError());
            q.coef = d_coef;
            q.codes = c->d_codes;
            q.pitch = c->code_pitch;
            q.lut = c->d_lut;
            q.clv = c->d_clv;
            q.scale = c->d_scale;
            q.weights = c->d_weights;
            q.edges = d_edges;
            q.S = c->S;
            // edges of a sum-table pass without a table (the root edge): up to four go out as a list, with their own
            // (much finer) split of the pattern axis and their own block sums at the end of the buffer
            q.n_list = 0;
            int n_plain = 0;
            for (int i = 0; i < n; ++i)
                if (edges[i].kind_a != SRC_SUMTABLE) {
                    if (n_plain < 4) q.list[n_plain] = i;
                    ++n_plain;
                }
            if (far_nodes == nullptr && c->up_sumtable && n_plain <= 4) q.n_list = n_plain;
            const int64_t span = 128 / K;
            int64_t parts = (c->S + span * K - 1) / (span * K);
            const int64_t lim = std::max<int64_t>(1, std::min<int64_t>((kPartialCap - kListArea) / (3 * n), std::max<int64_t>(8, (int64_t)c->sm_count * 16 / n)));
            const int list_parts = (int)std::min<int64_t>(kListParts, parts);
            if (parts > lim) parts = lim;
            double* const list_sums = c->d_partial_sums + (kPartialCap - kListArea);
            q.n_parts = q.n_list > 0 ? list_parts : (int)parts;
            q.partial_sums = q.n_list > 0 ? list_sums : c->d_partial_sums;
            dim3 grid((unsigned)parts, (unsigned)n);
            const dim3 grid_plain((unsigned)q.n_parts, (unsigned)(q.n_list > 0 ? q.n_list : n));
            if (n_plain > 0) {
                switch (K) {
                    case 1: dna_edge_sumtable_kernel<1><<<grid_plain, 128, 0, c->stream>>>(q); break;
                    case 2: dna_edge_sumtable_kernel<2><<<grid_plain, 128, 0, c->stream>>>(q); break;
                    case 4: dna_edge_sumtable_kernel<4><<<grid_plain, 128, 0, c->stream>>>(q); break;
                    default: dna_edge_sumtable_kernel<8><<<grid_plain, 128, 0, c->stream>>>(q); break;
                }
                c->launches++;
                PHB_CUDA(c, cudaGetLastError());
            }
            n_list_done = q.n_list;
            for (int j = 0; j < q.n_list; ++j) list_done[j] = q.list[j];
            list_parts_done = list_parts;
            if (c->up_sumtable && n_plain < n) {
                // every edge but the root edge has its sum table in its up block: those CTAs of the launch above left
                // at once (SRC_SUMTABLE), this launch does their work from one block read per edge
                DnaStArgs t;
                t.coef = d_coef;
                t.clv = c->d_clv;
                t.scale = c->d_scale;
                t.weights = c->d_weights;
                t.blocks = reinterpret_cast<const int32_t*>(d_edges);   // EdgeDesc::src_a, stride 4 ints
                t.S = c->S;
                t.n_parts = (int)parts;
                t.partial_sums = c->d_partial_sums;
                switch (K) {
                    case 1: dna_edge_st_kernel<1><<<grid, 128, 0, c->stream>>>(t); break;
                    case 2: dna_edge_st_kernel<2><<<grid, 128, 0, c->stream>>>(t); break;
                    case 4: dna_edge_st_kernel<4><<<grid, 128, 0, c->stream>>>(t); break;
                    default: dna_edge_st_kernel<8><<<grid, 128, 0, c->stream>>>(t); break;
                }
                c->launches++;
                PHB_CUDA(c, cudaGetLastError());
            }
            n_parts = (int)parts;
        } else {
            DerivArgs p;
            for (int order = 0; order < 3; ++order) {
                int st = launch_build_pmatrices(c, d_len, n, c->d_dmats + (size_t)order * cap * blk, order, chain_rule);
                if (st) return st;
            }
            p.mats = c->d_dmats;
            p.codes = c->d_codes;
            p.pitch = c->code_pitch;
            p.lut = c->d_lut;
            p.clv = c->d_clv;
            p.scale = c->d_scale;
            p.freqs = c->model_freqs();
            p.catw = c->model_catw();
            p.weights = c->d_weights;
            p.S = c->S;
            p.A = A;
            p.K = K;
            p.batch_cap = cap;
            p.edges = d_edges;
            const bool dna = dna_supported(c);
            const int64_t span = dna ? 128 / K : (int64_t)kDerivThreads * kSitesPerThread;
            int64_t parts = (c->S + span - 1) / span;
            // enough CTAs per edge to fill the chip across the whole batch, few enough for the block-sum buffer
            const int64_t lim = std::max<int64_t>(1, std::min<int64_t>(kPartialCap / (3 * n), std::max<int64_t>(8, (int64_t)c->sm_count * 16 / n)));
            if (parts > lim) parts = lim;
            p.n_parts = (int)parts;
            p.partial_sums = c->d_partial_sums;
            dim3 grid((unsigned)parts, (unsigned)n);
            if (dna) {
                switch (K) {
                    case 1: dna_edge_deriv_kernel<1><<<grid, 128, 0, c->stream>>>(p); break;
                    case 2: dna_edge_deriv_kernel<2><<<grid, 128, 0, c->stream>>>(p); break;
                    case 4: dna_edge_deriv_kernel<4><<<grid, 128, 0, c->stream>>>(p); break;
                    default: dna_edge_deriv_kernel<8><<<grid, 128, 0, c->stream>>>(p); break;
                }
            } else {
                const size_t smem = 3 * (size_t)A * A * sizeof(double);
                PHB_CUDA(c, cudaFuncSetAttribute(edge_deriv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                edge_deriv_kernel<<<grid, kDerivThreads, smem, c->stream>>>(p);
            }
            c->launches++;
            PHB_CUDA(c, cudaGetLastError());
            n_parts = (int)parts;
        }
        // the sums of edge i of the call end up at d_result[3 i .. 3 i + 2] (what the stream-ordered form leaves behind)
        double* const d_out = 3 * (size_t)(start + n) <= c->result_doubles ? c->d_result + 3 * (size_t)start : c->d_result;
        int st = launch_final_reduce(c, c->d_partial_sums, n_parts, 3 * n, d_out);
        if (st) return st;
        for (int j = 0; j < n_list_done; ++j) {   // overwrites what the launch above left in the listed edges' slots
            st = launch_final_reduce(c, c->d_partial_sums + (kPartialCap - kListArea) + (size_t)j * 3 * list_parts_done,
                                     list_parts_done, 3, d_out + 3 * (size_t)list_done[j]);
            if (st) return st;
        }
        if (out != nullptr)
            PHB_CUDA(c, cudaMemcpyAsync(out + 3 * (size_t)start, d_out, (size_t)3 * n * 8, cudaMemcpyDeviceToHost, c->stream));
    }
    if (out != nullptr) PHB_CUDA(c, cudaStreamSynchronize(c->stream));
    return PHB_OK;
}

}  // namespace phb
