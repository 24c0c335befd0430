// Felsenstein pruning for an arbitrary state count (protein A=20, codon A=61, binary A=2, ...).
//
// Same arithmetic and storage conventions as clv_dna.cu (reference: the `clv` gufunc,
// /root/reference/phylo_utils/likelihood/numba_likelihood_engine.py:10-46; per-pattern binary
// exponent instead of per-category log scalers).  Here the P.L contraction is a genuine small
// dense product, so the CTA stages everything in shared memory:
//
//   for each category k:  P1[k], P2[k] (A x A each)            -> smem
//                         child tiles  [TS patterns][A]         -> smem (coalesced, tips via LUT)
//                         thread t = pattern t of the tile: out[i] = (P1 row i . La[t]) (P2 row i . Lb[t])
//                         out tile [TS][A]                      -> smem -> coalesced store
//
// P elements are read as warp-wide broadcasts, child rows with an odd row stride (conflict free).
// The pattern maximum over all categories lives in the owning thread; if it falls under 2^-128 the
// thread rescales its K*A outputs in place after the tile has been written.
//
// Also here: the root / mixture / reduction kernel for arbitrary A, and the final deterministic sum.
#include "common.cuh"

namespace phb {

namespace {

struct GenArgs {
    const OpRow* rows;
    int row_begin, row_end;
    const double* pmats;  // [pidx][K][A][A]
    const uint8_t* codes;
    size_t pitch;
    const double* lut;    // [256][A]
    double* clv;
    int32_t* scale;
    int64_t S;
    int64_t n_tiles;
    int A, K, ld;         // ld = padded smem row length (odd)
};

__device__ __forceinline__ void stage_child(const GenArgs& p, int kind, int src, int64_t site0, int k, int ts,
                                            double* tile) {
    const int A = p.A;
    const size_t S = (size_t)p.S;
    if (kind == SRC_TIP) {
        const uint8_t* codes = p.codes + (size_t)src * p.pitch;
        for (int idx = threadIdx.x; idx < ts * A; idx += blockDim.x) {
            const int sl = idx / A, j = idx - sl * A;
            const int64_t s = site0 + sl;
            tile[sl * p.ld + j] = s < p.S ? __ldg(p.lut + (size_t)codes[s] * A + j) : 0.0;
        }
    } else {
        const double* base = p.clv + (size_t)src * S * p.K * A;
        for (int idx = threadIdx.x; idx < ts * A; idx += blockDim.x) {
            const int sl = idx / A, j = idx - sl * A;
            const int64_t s = site0 + sl;
            tile[sl * p.ld + j] = s < p.S ? base[((size_t)s * p.K + k) * A + j] : 0.0;
        }
    }
}

template <bool LEVEL>
__global__ void generic_prune_kernel(const GenArgs p) {
    extern __shared__ double sm[];
    const int A = p.A, K = p.K, ld = p.ld, ts = blockDim.x;
    double* P1 = sm;
    double* P2 = P1 + A * A;
    double* La = P2 + A * A;
    double* Lb = La + (size_t)ts * ld;
    double* Lo = Lb + (size_t)ts * ld;
    const int t = threadIdx.x;
    const size_t S = (size_t)p.S;

    const int64_t items = LEVEL ? (int64_t)(p.row_end - p.row_begin) * p.n_tiles : p.n_tiles;
    for (int64_t it = blockIdx.x; it < items; it += gridDim.x) {
        const int64_t tile = LEVEL ? it % p.n_tiles : it;
        const int r0 = LEVEL ? p.row_begin + (int)(it / p.n_tiles) : p.row_begin;
        const int r1 = LEVEL ? r0 + 1 : p.row_end;
        const int64_t site0 = tile * ts;
        const int64_t s = site0 + t;
        const bool ok = s < p.S;
        for (int r = r0; r < r1; ++r) {
            const OpRow row = p.rows[r];
            double m = 0.0;
            double* out = p.clv + (size_t)row.dst * S * K * A;
            for (int k = 0; k < K; ++k) {
                __syncthreads();  // previous users of the smem tiles are done
                const double* q1 = p.pmats + ((size_t)row.pidx[0] * K + k) * A * A;
                const double* q2 = p.pmats + ((size_t)row.pidx[1] * K + k) * A * A;
                for (int e = t; e < A * A; e += ts) {
                    P1[e] = __ldg(q1 + e);
                    P2[e] = __ldg(q2 + e);
                }
                stage_child(p, row.kind[0] == SRC_TIP ? SRC_TIP : SRC_GLOBAL, row.src[0], site0, k, ts, La);
                stage_child(p, row.kind[1] == SRC_TIP ? SRC_TIP : SRC_GLOBAL, row.src[1], site0, k, ts, Lb);
                __syncthreads();
                const double* la = La + (size_t)t * ld;
                const double* lb = Lb + (size_t)t * ld;
                double* lo = Lo + (size_t)t * ld;
                for (int i = 0; i < A; ++i) {
                    const double* r1p = P1 + i * A;
                    const double* r2p = P2 + i * A;
                    double x = 0.0, y = 0.0;
#pragma unroll 4
                    for (int j = 0; j < A; ++j) {
                        x = fma(r1p[j], la[j], x);
                        y = fma(r2p[j], lb[j], y);
                    }
                    const double o = x * y;
                    lo[i] = o;
                    m = fmax(m, o);
                }
                __syncthreads();
                for (int idx = t; idx < ts * A; idx += ts) {
                    const int sl = idx / A, j = idx - sl * A;
                    const int64_t sg = site0 + sl;
                    if (sg < p.S) out[((size_t)sg * K + k) * A + j] = Lo[sl * ld + j];
                }
            }
            __syncthreads();  // the whole tile (all categories) is in global memory and visible to the block
            int e = 0;
            if (ok) {
                if (row.kind[0] != SRC_TIP) e += p.scale[(size_t)row.src[0] * S + s];
                if (row.kind[1] != SRC_TIP) e += p.scale[(size_t)row.src[1] * S + s];
                const int hi = __double2hiint(m);
                if (hi < kScaleThresholdHi && hi >= 0x00100000) {
                    const int shift = 1023 - (hi >> 20);
                    const double f = pow2i(shift);
                    double* mine = out + (size_t)s * K * A;
                    for (int q = 0; q < K * A; ++q) mine[q] *= f;
                    e -= shift;
                }
                p.scale[(size_t)row.dst * S + s] = e;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
struct GenRootArgs {
    const double* pmats;  // [2][K][A][A]
    const uint8_t* codes;
    size_t pitch;
    const double* lut;
    const double* clv;
    const int32_t* scale;
    const double* freqs;
    const double* catw;
    const double* weights;
    int src[2], kind[2];
    int64_t S;
    int A, K;
    double* pattern_lnl;
    double* cat_lnl;
    double* root_clv;
    int32_t* root_scale;
    double* partial_sums;
};

// thread = pattern.  Two passes over the categories: the first finds the pattern maximum (for the
// exponent), the second forms pi . root per category.  Work is recomputed rather than stored because
// this kernel runs once per evaluation (1 of N-2 node updates) and A can be 61.
__global__ void generic_root_kernel(const GenRootArgs p) {
    extern __shared__ double sm[];
    const int A = p.A, K = p.K;
    double* Pa = sm;               // [A][A] current category
    double* Pb = Pa + A * A;
    double* s_red = Pb + A * A;    // [blockDim/32]
    const size_t S = (size_t)p.S;
    double acc = 0.0;
    const int64_t n_iter = (p.S + blockDim.x - 1) / blockDim.x;
    for (int64_t it = blockIdx.x; it < n_iter; it += gridDim.x) {
        const int64_t s = it * blockDim.x + threadIdx.x;
        const bool ok = s < p.S;
        const size_t ss = ok ? (size_t)s : 0;
        const double* va[2];
        int e = 0;
        for (int c = 0; c < 2; ++c) {
            if (p.kind[c] == SRC_TIP) {
                va[c] = p.lut + (size_t)p.codes[(size_t)p.src[c] * p.pitch + ss] * A;
            } else {
                va[c] = p.clv + ((size_t)p.src[c] * S + ss) * K * A;
                e += p.scale[(size_t)p.src[c] * S + ss];
            }
        }
        double mix = 0.0;
        double m = 0.0;
        for (int pass = 0; pass < 2; ++pass) {
            int shift = 0;
            double f2 = 1.0;
            if (pass == 1) {
                const int hi = __double2hiint(m);
                if (hi < kScaleThresholdHi && hi >= 0x00100000) {
                    shift = 1023 - (hi >> 20);
                    f2 = pow2i(shift);
                    e -= shift;
                }
                if (p.root_scale != nullptr && ok) p.root_scale[ss] = e;
            }
            for (int k = 0; k < K; ++k) {
                __syncthreads();
                for (int q = threadIdx.x; q < A * A; q += blockDim.x) {
                    Pa[q] = p.pmats[(size_t)k * A * A + q];
                    Pb[q] = p.pmats[(size_t)(K + k) * A * A + q];
                }
                __syncthreads();
                const double* la = va[0] + (p.kind[0] == SRC_TIP ? 0 : (size_t)k * A);
                const double* lb = va[1] + (p.kind[1] == SRC_TIP ? 0 : (size_t)k * A);
                double f = 0.0;
                for (int i = 0; i < A; ++i) {
                    double x = 0.0, y = 0.0;
                    for (int j = 0; j < A; ++j) {
                        x = fma(Pa[i * A + j], la[j], x);
                        y = fma(Pb[i * A + j], lb[j], y);
                    }
                    const double o = x * y;
                    if (pass == 0) {
                        m = fmax(m, o);
                    } else {
                        const double os = o * f2;
                        if (p.root_clv != nullptr && ok) p.root_clv[(ss * K + k) * A + i] = os;
                        f = fma(p.freqs[i], os, f);
                    }
                }
                if (pass == 1) {
                    if (p.cat_lnl != nullptr && ok) p.cat_lnl[ss * K + k] = f > 0 ? log(f) + (double)e * kLn2 : -INFINITY;
                    if (f > 0) mix += p.catw[k] * f;
                }
            }
        }
        if (ok) {
            const double lnl = mix > 0 ? log(mix) + (double)e * kLn2 : -INFINITY;
            p.pattern_lnl[ss] = lnl;
            acc += (p.weights ? p.weights[ss] : 1.0) * lnl;
        }
    }
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double tsum = 0;
        for (int w = 0; w < (int)blockDim.x / 32; ++w) tsum += s_red[w];
        p.partial_sums[blockIdx.x] = tsum;
    }
}

// out[o] = sum_i parts[o * n_parts + i], fixed summation order
__global__ void final_reduce_kernel(const double* __restrict__ parts, int n_parts, double* __restrict__ out) {
    __shared__ double s[256];
    const double* mine = parts + (size_t)blockIdx.x * n_parts;
    double acc = 0.0;
    for (int i = threadIdx.x; i < n_parts; i += 256) acc += mine[i];
    s[threadIdx.x] = acc;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[blockIdx.x] = s[0];
}

// largest byte of an array: validates tip codes without a host pass over N x S bytes.
// `head` bytes are handled one by one until the pointer is 16-byte aligned, then 16 at a time.
__global__ void max_code_kernel(const uint8_t* __restrict__ codes, size_t n, size_t head, int* __restrict__ out) {
    const size_t gtid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, gsz = (size_t)gridDim.x * blockDim.x;
    unsigned b = 0;
    for (size_t i = gtid; i < head; i += gsz) b = max(b, (unsigned)codes[i]);
    const size_t n16 = (n - head) / 16;
    const uint4* v = reinterpret_cast<const uint4*>(codes + head);
    unsigned m = 0;
    for (size_t i = gtid; i < n16; i += gsz) {
        const uint4 q = v[i];
        m = __vmaxu4(m, __vmaxu4(__vmaxu4(q.x, q.y), __vmaxu4(q.z, q.w)));
    }
    b = max(b, max(max(m & 0xff, (m >> 8) & 0xff), max((m >> 16) & 0xff, m >> 24)));
    for (size_t i = head + n16 * 16 + gtid; i < n; i += gsz) b = max(b, (unsigned)codes[i]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) b = max(b, __shfl_xor_sync(0xffffffffu, b, o));
    if ((threadIdx.x & 31) == 0 && b > 0) atomicMax(out, (int)b);
}

int tile_sites_for(const Ctx* c, size_t* smem_out, int* ld_out) {
    const int A = c->A;
    const int ld = A | 1;
    const size_t budget = c->smem_optin > 0 ? c->smem_optin : 48 * 1024;
    int ts = 128;
    while (ts > 32 && (2 * (size_t)A * A + 3 * (size_t)ts * ld) * sizeof(double) > budget) ts -= 32;
    // keep at least two CTAs per SM when that is cheap
    if (ts == 128 && (2 * (size_t)A * A + 3 * (size_t)ts * ld) * sizeof(double) * 2 > budget && A > 32) ts = 64;
    *smem_out = (2 * (size_t)A * A + 3 * (size_t)ts * ld) * sizeof(double);
    *ld_out = ld;
    return ts;
}

template <bool LEVEL>
int launch_generic(Ctx* c, const OpRow* d_rows, int row_begin, int row_end) {
    size_t smem;
    int ld;
    const int ts = tile_sites_for(c, &smem, &ld);
    GenArgs a;
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
    a.n_tiles = (c->S + ts - 1) / ts;
    a.A = c->A;
    a.K = c->K;
    a.ld = ld;
    auto kern = generic_prune_kernel<LEVEL>;
    PHB_CUDA(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    PHB_CUDA(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, ts, smem));
    if (per_sm < 1) per_sm = 1;
    const int64_t items = LEVEL ? (int64_t)(row_end - row_begin) * a.n_tiles : a.n_tiles;
    const int64_t cap = (int64_t)c->sm_count * per_sm;
    const int grid = (int)(items < cap ? items : cap);
    if (grid <= 0) return PHB_OK;
    kern<<<grid, ts, smem, c->stream>>>(a);
    c->launches++;
    PHB_CUDA(c, cudaGetLastError());
    return PHB_OK;
}

}  // namespace

int generic_run_rows(Ctx* c, const RowSet& rs, int mode) {
    const int n = rs.n_rows;
    if (n == 0) return PHB_OK;
    if (mode == PHB_MODE_LEVEL) {
        const std::vector<int32_t>& lv = *rs.levels;
        const int n_levels = (int)lv.size() - 1;
        for (int l = 0; l < n_levels; ++l) {
            const int b = lv[l], e = lv[l + 1];
            if (e <= b) continue;
            int st = launch_generic<true>(c, rs.d_rows, b, e);
            if (st != PHB_OK) return st;
        }
        return PHB_OK;
    }
    return launch_generic<false>(c, rs.d_rows, 0, n);
}

int generic_root(Ctx* c, int a, int b, bool want_cat, bool store_root) {
    GenRootArgs p;
    p.pmats = c->d_pmats + (size_t)(2 * c->max_rows()) * c->K * c->A * c->A;
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
    p.A = c->A;
    p.K = c->K;
    p.pattern_lnl = c->d_pattern_lnl;
    p.cat_lnl = want_cat ? c->d_cat_lnl : nullptr;
    p.root_clv = store_root ? c->d_root_clv : nullptr;
    p.root_scale = store_root ? c->d_root_scale : nullptr;
    p.partial_sums = c->d_partial_sums;
    const int threads = 128;
    const int64_t n_iter = (c->S + threads - 1) / threads;
    int64_t grid = (int64_t)c->sm_count * 4;
    if (grid > n_iter) grid = n_iter;
    if (grid > kMaxReduceBlocks) grid = kMaxReduceBlocks;
    if (grid < 1) grid = 1;
    const size_t smem = (2 * (size_t)c->A * c->A + threads / 32) * sizeof(double);
    PHB_CUDA(c, cudaFuncSetAttribute(generic_root_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    generic_root_kernel<<<(int)grid, threads, smem, c->stream>>>(p);
    c->launches++;
    PHB_CUDA(c, cudaGetLastError());
    return launch_final_reduce(c, c->d_partial_sums, (int)grid, 1, c->d_result);
}

int launch_max_code(Ctx* c, const uint8_t* d_codes, size_t n, int* worst) {
    int* d_flag = reinterpret_cast<int*>(c->d_result + 4 * kMaxEdgeBatch - 1);
    PHB_CUDA(c, cudaMemsetAsync(d_flag, 0, sizeof(int), c->stream));
    size_t head = (16 - reinterpret_cast<uintptr_t>(d_codes) % 16) % 16;
    if (head > n) head = n;
    size_t blocks = (n / 16 + 255) / 256;
    if (blocks > (size_t)c->sm_count * 16) blocks = (size_t)c->sm_count * 16;
    if (blocks < 1) blocks = 1;
    max_code_kernel<<<(unsigned)blocks, 256, 0, c->stream>>>(d_codes, n, head, d_flag);
    c->launches++;
    PHB_CUDA(c, cudaGetLastError());
    PHB_CUDA(c, cudaMemcpyAsync(worst, d_flag, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    PHB_CUDA(c, cudaStreamSynchronize(c->stream));
    return PHB_OK;
}

int launch_final_reduce(Ctx* c, const double* d_parts, int n_parts, int n_out, double* d_out) {
    final_reduce_kernel<<<n_out, 256, 0, c->stream>>>(d_parts, n_parts, d_out);
    c->launches++;
    PHB_CUDA(c, cudaGetLastError());
    return PHB_OK;
}

}  // namespace phb
