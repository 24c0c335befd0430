// Site-pattern compression on the GPU: characters -> state-set codes -> unique columns, weights, inverse index.
//
// Stands in for alignment_to_numpy's `np.unique(one_hot, axis=1, return_inverse=True, return_counts=True)`
// (/root/reference/phylo_utils/alignment/alignment.py:40-57) and the per-character charmap look-up of
// seq_to_partials (:26-37).  np.unique sorts the columns lexicographically over the taxon-major / state-minor
// flattening of the 0/1 array; codes are ranks of the 0/1 rows in that same order (charmaps.CodeBook), so the
// result is reproduced bit for bit by sorting columns by their byte string of codes:
//
//   1. (optional) chars -> codes through a 256-entry byte table; an unmapped character is reported with its
//      position (the reference raises KeyError from the dict look-up);
//   2. a least-significant-digit radix sort of the column permutation: one stable counting-sort pass per group
//      of taxa, from the last taxon to the first.  A digit packs as many consecutive taxa as fit in 8 bits
//      (two taxa for nucleotide codes).  Each pass = per-block digit histogram, one scan of the (digit, block)
//      counts, stable scatter (warp match_any ranks) - the column data itself never moves, only 4-byte indices;
//   3. boundaries between runs of equal columns (early-exit compare of neighbouring columns), scan ->
//      pattern id; inverse_index[site] = id, weights = run lengths, patterns = first column of each run.
//
// All integer work, bit-exact by construction; HBM/L2-bound on the 4-byte permutation (8 bytes per site and
// pass) plus one random byte gather per site and taxon.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>

#include "common.cuh"

namespace phb {
namespace {

constexpr int kSortThreads = 256;
constexpr int kItems = 8;                          // elements per thread
constexpr int kTile = kSortThreads * kItems;       // elements per block
constexpr int kWarps = kSortThreads / 32;

struct DevBuf {
    void* p = nullptr;
    cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes ? bytes : 1); }
    ~DevBuf() {
        if (p) cudaFree(p);
    }
    template <class T>
    T* as() const {
        return static_cast<T*>(p);
    }
};

int cz_fail(cudaError_t e, const char* where) {
    set_thread_error(std::string(where) + ": " + cudaGetErrorString(e));
    cudaGetLastError();
    return e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver ? PHB_ERR_NO_DEVICE : PHB_ERR_CUDA;
}
#define CZ_CUDA(expr)                                                        \
    do {                                                                     \
        cudaError_t _e = (expr);                                             \
        if (_e != cudaSuccess) return cz_fail(_e, "phb_compress_patterns");  \
    } while (0)

// chars -> codes in place; first_bad = smallest flat index of an unmapped character (or LLONG_MAX)
__global__ void encode_kernel(uint8_t* __restrict__ data, size_t n, const uint8_t* __restrict__ table,
                              unsigned long long* __restrict__ first_bad) {
    __shared__ uint8_t s_table[256];
    if (threadIdx.x < 256) s_table[threadIdx.x] = table[threadIdx.x];
    __syncthreads();
    unsigned long long bad = ~0ull;
    const size_t n16 = n / 16;
    uint4* v = reinterpret_cast<uint4*>(data);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) {
        uint4 w = v[i];
        unsigned* parts = reinterpret_cast<unsigned*>(&w);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            unsigned out = 0;
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const unsigned code = s_table[(parts[j] >> (8 * b)) & 0xff];
                if (code == 255u) bad = min(bad, (unsigned long long)(i * 16 + j * 4 + b));
                out |= code << (8 * b);
            }
            parts[j] = out;
        }
        v[i] = w;
    }
    for (size_t i = n16 * 16 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const unsigned code = s_table[data[i]];
        if (code == 255u) bad = min(bad, (unsigned long long)i);
        data[i] = (uint8_t)code;
    }
    if (bad != ~0ull) atomicMin(first_bad, bad);
}

__global__ void max_byte_kernel(const uint8_t* __restrict__ data, size_t n, unsigned* __restrict__ out) {
    unsigned m = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) m = max(m, (unsigned)data[i]);
    for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m) atomicMax(out, m);
}

__global__ void iota_kernel(int32_t* __restrict__ perm, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) perm[i] = (int32_t)i;
}

// the digit of column `site` in a pass over taxa [t0, t0 + nt): first taxon most significant
__device__ __forceinline__ unsigned digit_of(const uint8_t* __restrict__ codes, int64_t S, int t0, int nt, int bits, int32_t site) {
    unsigned d = 0;
    for (int j = 0; j < nt; ++j) d = (d << bits) | codes[(size_t)(t0 + j) * S + site];
    return d;
}

// counts[digit][block]
__global__ void __launch_bounds__(kSortThreads) hist_kernel(const uint8_t* __restrict__ codes, int64_t S, int t0, int nt,
                                                            int bits, const int32_t* __restrict__ perm, int n_bins,
                                                            int32_t* __restrict__ counts) {
    __shared__ int s_hist[kWarps][256];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int b = lane; b < n_bins; b += 32) s_hist[warp][b] = 0;
    __syncwarp();
    const int64_t base = (int64_t)blockIdx.x * kTile + warp * (kTile / kWarps);
#pragma unroll
    for (int r = 0; r < kItems; ++r) {
        const int64_t i = base + r * 32 + lane;
        const bool valid = i < S;
        const unsigned d = valid ? digit_of(codes, S, t0, nt, bits, perm[i]) : 0xffffffffu;
        const unsigned peers = __match_any_sync(0xffffffffu, d);
        if (valid && lane == __ffs(peers) - 1) s_hist[warp][d] += __popc(peers);
        __syncwarp();
    }
    __syncthreads();
    for (int b = threadIdx.x; b < n_bins; b += kSortThreads) {
        int t = 0;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) t += s_hist[w][b];
        counts[(size_t)b * gridDim.x + blockIdx.x] = t;
    }
}

// in-place exclusive scan of n int32 values by ONE block (n is a few hundred thousand at most per call);
// total -> *total_out if given
__global__ void __launch_bounds__(1024) scan_small_kernel(int32_t* __restrict__ data, int64_t n, int32_t* __restrict__ total_out) {
    __shared__ int s_warp[32];
    __shared__ int s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int64_t base = 0; base < n; base += 1024 * 4) {
        // four consecutive values per thread
        const int64_t i0 = base + (int64_t)threadIdx.x * 4;
        int v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = i0 + j < n ? data[i0 + j] : 0;
        const int mine = v[0] + v[1] + v[2] + v[3];
        int inc = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == 31) s_warp[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            int w = s_warp[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= o) w += t;
            }
            s_warp[lane] = w;   // inclusive over warps
        }
        __syncthreads();
        int run = s_carry + (warp ? s_warp[warp - 1] : 0) + inc - mine;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (i0 + j < n) data[i0 + j] = run;
            run += v[j];
        }
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = run;
        __syncthreads();
    }
    if (threadIdx.x == 0 && total_out) *total_out = s_carry;
}

// stable scatter of one pass: offsets[digit][block] = first output position of this block's elements with that digit
__global__ void __launch_bounds__(kSortThreads) scatter_kernel(const uint8_t* __restrict__ codes, int64_t S, int t0, int nt,
                                                               int bits, const int32_t* __restrict__ perm_in,
                                                               int32_t* __restrict__ perm_out, int n_bins,
                                                               const int32_t* __restrict__ offsets) {
    __shared__ int s_cnt[kWarps][256];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int b = lane; b < n_bins; b += 32) s_cnt[warp][b] = 0;
    __syncwarp();
    const int64_t base = (int64_t)blockIdx.x * kTile + warp * (kTile / kWarps);
    int32_t site[kItems];
    unsigned dig[kItems];
#pragma unroll
    for (int r = 0; r < kItems; ++r) {
        const int64_t i = base + r * 32 + lane;
        const bool valid = i < S;
        site[r] = valid ? perm_in[i] : 0;
        dig[r] = valid ? digit_of(codes, S, t0, nt, bits, site[r]) : 0xffffffffu;
        const unsigned peers = __match_any_sync(0xffffffffu, dig[r]);
        if (valid && lane == __ffs(peers) - 1) s_cnt[warp][dig[r]] += __popc(peers);
        __syncwarp();
    }
    __syncthreads();
    // per digit: running start over the warps of this block (warps hold consecutive sub-ranges)
    for (int b = threadIdx.x; b < n_bins; b += kSortThreads) {
        int run = offsets[(size_t)b * gridDim.x + blockIdx.x];
#pragma unroll
        for (int w = 0; w < kWarps; ++w) {
            const int t = s_cnt[w][b];
            s_cnt[w][b] = run;
            run += t;
        }
    }
    __syncthreads();
    const unsigned lt = (1u << lane) - 1;
#pragma unroll
    for (int r = 0; r < kItems; ++r) {
        const int64_t i = base + r * 32 + lane;
        const bool valid = i < S;
        const unsigned peers = __match_any_sync(0xffffffffu, dig[r]);
        if (valid) perm_out[s_cnt[warp][dig[r]] + __popc(peers & lt)] = site[r];
        __syncwarp();
        if (valid && lane == __ffs(peers) - 1) s_cnt[warp][dig[r]] += __popc(peers);
        __syncwarp();
    }
}

// flag[i] = 1 when sorted column i differs from sorted column i-1 (flag[0] = 1)
__global__ void boundary_kernel(const uint8_t* __restrict__ codes, int64_t S, int n_tips, const int32_t* __restrict__ perm,
                                int32_t* __restrict__ flag) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= S) return;
    int f = 1;
    if (i > 0) {
        const int32_t a = perm[i], b = perm[i - 1];
        f = 0;
        for (int t = 0; t < n_tips; ++t)
            if (codes[(size_t)t * S + a] != codes[(size_t)t * S + b]) {
                f = 1;
                break;
            }
    }
    flag[i] = f;
}

// large exclusive scan: per-block sums -> scan_small -> add back
__global__ void __launch_bounds__(kSortThreads) block_sum_kernel(const int32_t* __restrict__ in, int64_t n, int32_t* __restrict__ sums) {
    __shared__ int s_w[kWarps];
    const int64_t base = (int64_t)blockIdx.x * kTile;
    int t = 0;
#pragma unroll
    for (int r = 0; r < kItems; ++r) {
        const int64_t i = base + r * kSortThreads + threadIdx.x;
        if (i < n) t += in[i];
    }
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = t;
    __syncthreads();
    if (threadIdx.x == 0) {
        int s = 0;
        for (int w = 0; w < kWarps; ++w) s += s_w[w];
        sums[blockIdx.x] = s;
    }
}

// id[i] = (number of boundaries at positions <= i) - 1; block_off = exclusive scan of the block sums
__global__ void __launch_bounds__(kSortThreads) pattern_id_kernel(const int32_t* __restrict__ flag, int64_t n,
                                                                  const int32_t* __restrict__ block_off,
                                                                  int32_t* __restrict__ id) {
    __shared__ int s_w[kWarps];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t i0 = (int64_t)blockIdx.x * kTile + (int64_t)threadIdx.x * kItems;   // kItems consecutive values per thread
    int v[kItems], mine = 0;
#pragma unroll
    for (int j = 0; j < kItems; ++j) {
        v[j] = i0 + j < n ? flag[i0 + j] : 0;
        mine += v[j];
    }
    int inc = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) s_w[warp] = inc;
    __syncthreads();
    int before = block_off[blockIdx.x] + inc - mine;
    for (int w = 0; w < warp; ++w) before += s_w[w];
#pragma unroll
    for (int j = 0; j < kItems; ++j) {
        before += v[j];
        if (i0 + j < n) id[i0 + j] = before - 1;
    }
}

// inverse_index[site] = pattern id; start[id] = first sorted position of the run
__global__ void inverse_kernel(const int32_t* __restrict__ perm, const int32_t* __restrict__ flag,
                               const int32_t* __restrict__ id, int64_t S, int64_t* __restrict__ inverse,
                               int32_t* __restrict__ start) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= S) return;
    inverse[perm[i]] = id[i];
    if (flag[i]) start[id[i]] = (int32_t)i;
}

__global__ void weights_kernel(const int32_t* __restrict__ start, int64_t n_pat, int64_t S, int64_t* __restrict__ weights) {
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_pat) return;
    const int64_t next = g + 1 < n_pat ? start[g + 1] : S;
    weights[g] = next - start[g];
}

// patterns[t][g] = codes[t][perm[start[g]]]
__global__ void gather_patterns_kernel(const uint8_t* __restrict__ codes, int64_t S, const int32_t* __restrict__ perm,
                                       const int32_t* __restrict__ start, int64_t n_pat, int n_tips,
                                       uint8_t* __restrict__ out) {
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_pat) return;
    const int32_t site = perm[start[g]];
    for (int t = blockIdx.y; t < n_tips; t += gridDim.y) out[(size_t)t * n_pat + g] = codes[(size_t)t * S + site];
}

// PHB_COMPRESS_TIMING=1: print the wall time of each phase of phb_compress_patterns to stderr (profiling aid)
struct PhaseTimer {
    bool on;
    cudaStream_t stream;
    double t0;
    static double now() {
        timespec ts;
        clock_gettime(CLOCK_MONOTONIC, &ts);
        return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
    }
    explicit PhaseTimer(cudaStream_t s) : on(tuning().compress_timing), stream(s), t0(now()) {}
    void mark(const char* what) {
        if (!on) return;
        cudaStreamSynchronize(stream);
        const double t = now();
        fprintf(stderr, "[phb_compress_patterns] %-28s %9.3f ms\n", what, t - t0);
        t0 = t;
    }
};

inline int grid_for(int64_t n, int threads) { return (int)std::min<int64_t>((n + threads - 1) / threads, 1 << 30); }

}  // namespace
}  // namespace phb

using namespace phb;

extern "C" int phb_compress_patterns(int device, const uint8_t* data, const uint8_t* byte_table, int n_tips,
                                     int64_t n_sites, uint8_t* patterns_out, int64_t* weights_out,
                                     int64_t* inverse_out, int64_t* n_patterns_out, int64_t* bad_index_out) {
    if (!data || !patterns_out || !weights_out || !inverse_out || !n_patterns_out || n_tips < 1 || n_sites < 0) {
        set_thread_error("phb_compress_patterns: bad argument");
        return PHB_ERR_INVALID;
    }
    if (n_sites >= ((int64_t)1 << 31) - kTile) {
        set_thread_error("phb_compress_patterns: more than 2^31 sites");
        return PHB_ERR_UNSUPPORTED;
    }
    if (bad_index_out) *bad_index_out = -1;
    *n_patterns_out = 0;
    if (n_sites == 0) return PHB_OK;
    int n_dev = 0;
    cudaError_t e = cudaGetDeviceCount(&n_dev);
    if (e != cudaSuccess || n_dev == 0) {
        set_thread_error("phb_compress_patterns: no CUDA device available; this engine has no CPU fallback");
        cudaGetLastError();
        return PHB_ERR_NO_DEVICE;
    }
    if (device < 0 || device >= n_dev) {
        set_thread_error("phb_compress_patterns: device index out of range");
        return PHB_ERR_INVALID;
    }
    CZ_CUDA(cudaSetDevice(device));
    const int64_t S = n_sites;
    const size_t n_bytes = (size_t)n_tips * S;
    const int n_blocks = (int)((S + kTile - 1) / kTile);
    DevBuf d_codes, d_perm_a, d_perm_b, d_counts, d_flag, d_id, d_start, d_sums, d_inverse, d_weights, d_patterns, d_small;
    CZ_CUDA(d_codes.alloc(n_bytes + 16));
    CZ_CUDA(d_perm_a.alloc((size_t)S * 4));
    CZ_CUDA(d_perm_b.alloc((size_t)S * 4));
    CZ_CUDA(d_counts.alloc((size_t)256 * n_blocks * 4));
    CZ_CUDA(d_small.alloc(1024));
    cudaStream_t stream = nullptr;
    PhaseTimer timer(stream);
    timer.mark("allocations");
    CZ_CUDA(cudaMemcpyAsync(d_codes.p, data, n_bytes, cudaMemcpyHostToDevice, stream));
    timer.mark("host -> device");
    unsigned long long* d_bad = d_small.as<unsigned long long>();
    unsigned* d_max = reinterpret_cast<unsigned*>(d_small.as<uint8_t>() + 8);
    int32_t* d_total = reinterpret_cast<int32_t*>(d_small.as<uint8_t>() + 16);
    uint8_t* d_table = d_small.as<uint8_t>() + 256;
    CZ_CUDA(cudaMemsetAsync(d_small.p, 0xff, 8, stream));
    CZ_CUDA(cudaMemsetAsync(d_small.as<uint8_t>() + 8, 0, 16, stream));
    const int wide_grid = std::min(grid_for((int64_t)(n_bytes / 16 + 1), 256), 148 * 16);
    if (byte_table) {
        CZ_CUDA(cudaMemcpyAsync(d_table, byte_table, 256, cudaMemcpyHostToDevice, stream));
        encode_kernel<<<wide_grid, 256, 0, stream>>>(d_codes.as<uint8_t>(), n_bytes, d_table, d_bad);
        CZ_CUDA(cudaGetLastError());
    }
    max_byte_kernel<<<wide_grid, 256, 0, stream>>>(d_codes.as<uint8_t>(), n_bytes, d_max);
    CZ_CUDA(cudaGetLastError());
    unsigned long long h_bad = 0;
    unsigned h_max = 0;
    CZ_CUDA(cudaMemcpyAsync(&h_bad, d_bad, 8, cudaMemcpyDeviceToHost, stream));
    CZ_CUDA(cudaMemcpyAsync(&h_max, d_max, 4, cudaMemcpyDeviceToHost, stream));
    CZ_CUDA(cudaStreamSynchronize(stream));
    timer.mark("byte table, validation");
    if (h_bad != ~0ull) {
        if (bad_index_out) *bad_index_out = (int64_t)h_bad;
        set_thread_error("phb_compress_patterns: a character is not in the byte table");
        return PHB_ERR_INVALID;
    }
    // digit geometry: as many taxa per pass as fit in 8 bits
    int bits = 1;
    while ((1u << bits) <= h_max) ++bits;
    const int per_pass = std::max(1, 8 / bits);
    const uint8_t* codes = d_codes.as<uint8_t>();
    int32_t* perm_in = d_perm_a.as<int32_t>();
    int32_t* perm_out = d_perm_b.as<int32_t>();
    iota_kernel<<<grid_for(S, 256), 256, 0, stream>>>(perm_in, S);
    CZ_CUDA(cudaGetLastError());
    for (int t_end = n_tips; t_end > 0; t_end -= per_pass) {
        const int t0 = std::max(0, t_end - per_pass), nt = t_end - t0;
        const int n_bins = 1 << (bits * nt);
        hist_kernel<<<n_blocks, kSortThreads, 0, stream>>>(codes, S, t0, nt, bits, perm_in, n_bins, d_counts.as<int32_t>());
        scan_small_kernel<<<1, 1024, 0, stream>>>(d_counts.as<int32_t>(), (int64_t)n_bins * n_blocks, nullptr);
        scatter_kernel<<<n_blocks, kSortThreads, 0, stream>>>(codes, S, t0, nt, bits, perm_in, perm_out, n_bins,
                                                               d_counts.as<int32_t>());
        std::swap(perm_in, perm_out);
    }
    CZ_CUDA(cudaGetLastError());
    timer.mark("radix sort passes");
    // runs of equal columns
    CZ_CUDA(d_flag.alloc((size_t)S * 4));
    CZ_CUDA(d_id.alloc((size_t)S * 4));
    CZ_CUDA(d_start.alloc((size_t)S * 4));
    CZ_CUDA(d_sums.alloc((size_t)n_blocks * 4));
    CZ_CUDA(d_inverse.alloc((size_t)S * 8));
    boundary_kernel<<<grid_for(S, 256), 256, 0, stream>>>(codes, S, n_tips, perm_in, d_flag.as<int32_t>());
    block_sum_kernel<<<n_blocks, kSortThreads, 0, stream>>>(d_flag.as<int32_t>(), S, d_sums.as<int32_t>());
    scan_small_kernel<<<1, 1024, 0, stream>>>(d_sums.as<int32_t>(), n_blocks, d_total);
    pattern_id_kernel<<<n_blocks, kSortThreads, 0, stream>>>(d_flag.as<int32_t>(), S, d_sums.as<int32_t>(), d_id.as<int32_t>());
    inverse_kernel<<<grid_for(S, 256), 256, 0, stream>>>(perm_in, d_flag.as<int32_t>(), d_id.as<int32_t>(), S,
                                                         d_inverse.as<int64_t>(), d_start.as<int32_t>());
    CZ_CUDA(cudaGetLastError());
    int32_t n_pat = 0;
    CZ_CUDA(cudaMemcpyAsync(&n_pat, d_total, 4, cudaMemcpyDeviceToHost, stream));
    CZ_CUDA(cudaStreamSynchronize(stream));
    timer.mark("runs, ids, inverse index");
    CZ_CUDA(d_weights.alloc((size_t)n_pat * 8));
    CZ_CUDA(d_patterns.alloc((size_t)n_tips * n_pat));
    weights_kernel<<<grid_for(n_pat, 256), 256, 0, stream>>>(d_start.as<int32_t>(), n_pat, S, d_weights.as<int64_t>());
    dim3 ggrid(grid_for(n_pat, 256), std::min(n_tips, 65535));
    gather_patterns_kernel<<<ggrid, 256, 0, stream>>>(codes, S, perm_in, d_start.as<int32_t>(), n_pat, n_tips,
                                                     d_patterns.as<uint8_t>());
    CZ_CUDA(cudaGetLastError());
    timer.mark("weights, pattern gather");
    CZ_CUDA(cudaMemcpyAsync(patterns_out, d_patterns.p, (size_t)n_tips * n_pat, cudaMemcpyDeviceToHost, stream));
    CZ_CUDA(cudaMemcpyAsync(weights_out, d_weights.p, (size_t)n_pat * 8, cudaMemcpyDeviceToHost, stream));
    CZ_CUDA(cudaMemcpyAsync(inverse_out, d_inverse.p, (size_t)S * 8, cudaMemcpyDeviceToHost, stream));
    CZ_CUDA(cudaStreamSynchronize(stream));
    timer.mark("device -> host");
    *n_patterns_out = n_pat;
    return PHB_OK;
}
