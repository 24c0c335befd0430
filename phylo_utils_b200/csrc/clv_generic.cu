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
#include <algorithm>

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
// Root finishing: the virtual-root partial has been computed into the root block by the ordinary pruning
// kernels (it is just one more `clv` row: tree_model.py:178-198); what is left of
// compute_likelihood_at_edge (tree_model.py:200-217) is pi-dot per category (lnl_node), the mixture, the log and
// the weighted sum.  One warp per pattern: lanes stride over the states (coalesced), five shuffles per sum.
struct RootFinishArgs {
    const double* root_clv;    // [S][K][A]
    const int32_t* root_scale; // [S]
    const double* freqs;
    const double* catw;
    const double* weights;
    int64_t S;
    int A, K;
    double* pattern_lnl;
    double* cat_lnl;           // [S][K] or null
    double* partial_sums;
};

__global__ void __launch_bounds__(128) root_finish_kernel(const RootFinishArgs p) {
    __shared__ double s_red[4];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int A = p.A, K = p.K;
    double acc = 0.0;
    for (int64_t s = (int64_t)blockIdx.x * 4 + warp; s < p.S; s += (int64_t)gridDim.x * 4) {
        const double* v = p.root_clv + (size_t)s * K * A;
        const double shift = (double)p.root_scale[s] * kLn2;
        double mix = 0.0;
        for (int k = 0; k < K; ++k) {
            double f = 0.0;
            for (int i = lane; i < A; i += 32) f = fma(p.freqs[i], v[k * A + i], f);
            f = warp_sum(f);
            if (p.cat_lnl != nullptr && lane == 0) p.cat_lnl[(size_t)s * K + k] = f > 0 ? log(f) + shift : -INFINITY;
            if (f > 0) mix = fma(p.catw[k], f, mix);
        }
        if (lane == 0) {
            const double lnl = mix > 0 ? log(mix) + shift : -INFINITY;
            p.pattern_lnl[s] = lnl;
            acc += (p.weights ? p.weights[s] : 1.0) * lnl;
        }
    }
    if (lane == 0) s_red[warp] = acc;
    __syncthreads();
    if (threadIdx.x == 0) p.partial_sums[blockIdx.x] = s_red[0] + s_red[1] + s_red[2] + s_red[3];
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

// The same reduction followed by the sum over the ranks of one box, in the same kernel: the rank's total goes straight
// into every peer's exchange buffer (mapped through CUDA IPC: stores over NVLink), then the kernel waits for the peers'
// values in its own buffer and adds them in rank order - every rank ends with the same bits, no collective library call,
// no second launch.  Cell = {value, sequence number}; the number is stored with release semantics at system scope behind
// the value and read with acquire semantics, so a matching number implies a visible value.  Two parities of cells: a
// rank can be at most one exchange ahead of a peer that has not read its cells yet (it needs that peer's value to get
// further).  The wait is bounded (about a minute): a rank that never shows up yields NaN, not a hung GPU.
struct PeerCell {
    double value;
    unsigned long long seq;
};
struct PeerArgs {
    PeerCell* cells[kMaxPeers];
    int rank, world, parity;
    unsigned long long seq;
};

__global__ void final_reduce_peer_kernel(const double* __restrict__ parts, int n_parts, double* __restrict__ out,
                                         const __grid_constant__ PeerArgs a) {
    __shared__ double s[256];
    double acc = 0.0;
    for (int i = threadIdx.x; i < n_parts; i += 256) acc += parts[i];
    s[threadIdx.x] = acc;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o];
        __syncthreads();
    }
    const double mine = s[0];
    __syncthreads();
    const int r = threadIdx.x;
    if (r < a.world) {
        PeerCell* dst = a.cells[r] + a.parity * kMaxPeers + a.rank;
        asm volatile("st.relaxed.sys.global.f64 [%0], %1;" ::"l"(&dst->value), "d"(mine) : "memory");
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(&dst->seq), "l"(a.seq) : "memory");
        const PeerCell* src = a.cells[a.rank] + a.parity * kMaxPeers + r;
        double v = __longlong_as_double(0x7ff8000000000000ll);
#pragma unroll 1
        for (int spin = 0; spin < (1 << 26); ++spin) {
            unsigned long long have;
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(have) : "l"(&src->seq) : "memory");
            if (have == a.seq) {
                asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(&src->value) : "memory");
                break;
            }
            __nanosleep(spin < 4096 ? 20 : 1000);
        }
        s[r] = v;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double total = 0.0;
        for (int i = 0; i < a.world; ++i) total += s[i];
        out[0] = total;
    }
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
    (void)store_root;   // the root partial always lands in the root block
    // the root as one more row: children a, b; P blocks 2*max_rows (+1) = P(0), P(length)
    OpRow row{};
    row.dst = c->root_block;
    const int nodes[2] = {a, b};
    for (int i = 0; i < 2; ++i) {
        if (c->node_tip[nodes[i]] >= 0) {
            row.kind[i] = SRC_TIP;
            row.src[i] = c->node_tip[nodes[i]];
        } else {
            row.kind[i] = SRC_GLOBAL;
            row.src[i] = c->node_slot[nodes[i]];
        }
        row.pidx[i] = 2 * c->max_rows() + i;
    }
    if (row.kind[0] != SRC_TIP && row.kind[1] == SRC_TIP) {   // canonical order: tips first
        std::swap(row.kind[0], row.kind[1]);
        std::swap(row.src[0], row.src[1]);
        std::swap(row.pidx[0], row.pidx[1]);
    }
    OpRow* d_row = c->d_rows + c->max_rows();
    PHB_CUDA(c, cudaMemcpyAsync(d_row, &row, sizeof row, cudaMemcpyHostToDevice, c->stream));
    PHB_CUDA(c, cudaStreamSynchronize(c->stream));   // `row` lives on the stack
    static const std::vector<int32_t> one_level = {0, 1};
    const RowSet rs{d_row, 1, &one_level};
    int st = run_rows(c, rs, PHB_MODE_LEVEL);
    if (st) return st;

    RootFinishArgs p;
    p.root_clv = c->d_root_clv;
    p.root_scale = c->d_root_scale;
    p.freqs = c->model_freqs();
    p.catw = c->model_catw();
    p.weights = c->d_weights;
    p.S = c->S;
    p.A = c->A;
    p.K = c->K;
    p.pattern_lnl = c->d_pattern_lnl;
    p.cat_lnl = want_cat ? c->d_cat_lnl : nullptr;
    p.partial_sums = c->d_partial_sums;
    int64_t grid = (c->S + 3) / 4;
    if (grid > (int64_t)c->sm_count * 16) grid = (int64_t)c->sm_count * 16;
    if (grid < 1) grid = 1;
    root_finish_kernel<<<(int)grid, 128, 0, c->stream>>>(p);
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
    if (n_out == 1 && c->peer.use_now) {
        c->peer.use_now = false;
        if (!c->peer.connected) return c->fail(PHB_ERR_STATE, "phb_peer_sum_next: no peers connected (phb_peer_connect)");
        PeerArgs a;
        for (int r = 0; r < kMaxPeers; ++r) a.cells[r] = static_cast<PeerCell*>(c->peer.cells[r]);
        a.rank = c->peer.rank;
        a.world = c->peer.world;
        a.seq = ++c->peer.epoch;
        a.parity = (int)(a.seq & 1);
        final_reduce_peer_kernel<<<1, 256, 0, c->stream>>>(d_parts, n_parts, d_out, a);
        c->launches++;
        PHB_CUDA(c, cudaGetLastError());
        return PHB_OK;
    }
    final_reduce_kernel<<<n_out, 256, 0, c->stream>>>(d_parts, n_parts, d_out);
    c->launches++;
    PHB_CUDA(c, cudaGetLastError());
    return PHB_OK;
}

}  // namespace phb
