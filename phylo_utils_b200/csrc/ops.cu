// Stand-alone operators with the reference gufuncs' exact semantics
// (/root/reference/phylo_utils/likelihood/numba_likelihood_engine.py): host arrays in, host arrays
// out, natural-log per-(pattern, category) scalers, rescale by the category maximum.  They exist
// so that code written against `clv / lnl_node / lnl_branch / lnl_branch_derivs` keeps working and so
// that the tests can compare operator by operator; the tree path (api.cu) never goes through them.
#include <cmath>

#include "common.cuh"

namespace phb {
namespace {

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

int op_fail(cudaError_t e, const char* where) {
    set_thread_error(std::string(where) + ": " + cudaGetErrorString(e));
    cudaGetLastError();
    return e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver ? PHB_ERR_NO_DEVICE : PHB_ERR_CUDA;
}
#define OP_CUDA(expr, where)                                  \
    do {                                                      \
        cudaError_t _e = (expr);                              \
        if (_e != cudaSuccess) return op_fail(_e, where);     \
    } while (0)

// fp64 throughput probes (phb_op_fp64_peak): eight independent dependency chains per thread
__global__ void __launch_bounds__(256) fp64_fma_probe(double* out, int iters) {
    double v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = 1.0 + 1e-9 * (threadIdx.x + j);
    const double m = 1.0 - 1e-12, c = 1e-12;
#pragma unroll 1
    for (int i = 0; i < iters; i += 16)          // 128 DFMA per trip: loop control is ~2 % of the instructions
#pragma unroll
        for (int u = 0; u < 16; ++u)
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = fma(v[j], m, c);
    double s = 0.0;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += v[j];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256) fp64_mma_probe(double* out, int iters) {
    double d[8][2];
#pragma unroll
    for (int j = 0; j < 8; ++j) d[j][0] = d[j][1] = 0.0;
    const double a = 1.0 + 1e-9 * threadIdx.x, b = 1e-3;
#pragma unroll 1
    for (int i = 0; i < iters; i += 8)
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
            for (int j = 0; j < 8; ++j)
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(d[j][0]), "+d"(d[j][1])
                         : "d"(a), "d"(b));
    double s = 0.0;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += d[j][0] + d[j][1];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// numba_likelihood_engine.py:14-46, one thread per (pattern, category)
__global__ void op_clv_kernel(int64_t S, int K, int A, const double* __restrict__ p1, const double* __restrict__ p2,
                              const double* __restrict__ clv1, const double* __restrict__ clv2,
                              const double* __restrict__ sa, const double* __restrict__ sb,
                              double* __restrict__ s_out, double* __restrict__ out) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= S * K) return;
    const int k = (int)(idx % K);
    const double* P1 = p1 + (size_t)k * A * A;
    const double* P2 = p2 + (size_t)k * A * A;
    const double* a = clv1 + (size_t)idx * A;
    const double* b = clv2 + (size_t)idx * A;
    double* o = out + (size_t)idx * A;
    double m = -INFINITY;
    for (int i = 0; i < A; ++i) {
        double x = 0.0, y = 0.0;
        for (int j = 0; j < A; ++j) {
            x = fma(P1[i * A + j], a[j], x);
            y = fma(P2[i * A + j], b[j], y);
        }
        const double v = x * y;
        o[i] = v;
        m = fmax(m, v);
    }
    if (m < kScaleThreshold && m > 0) {
        s_out[idx] = sa[idx] + sb[idx] + log(m);
        for (int i = 0; i < A; ++i) o[i] /= m;
    } else {
        s_out[idx] = sa[idx] + sb[idx];
    }
}

// :82-87
__global__ void op_lnl_node_kernel(int64_t n, int A, const double* __restrict__ pi, const double* __restrict__ partials,
                                   const double* __restrict__ scale, double* __restrict__ out) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n) return;
    const double* v = partials + (size_t)idx * A;
    double f = 0.0;
    for (int i = 0; i < A; ++i) f = fma(v[i], pi[i], f);
    out[idx] = f > 0 ? log(f) + scale[idx] : -INFINITY;
}

// :49-79, one thread per pattern; nd = 0 (lnl_branch) or 2 (lnl_branch_derivs)
__global__ void op_lnl_branch_kernel(int64_t S, int A, int nd, const double* __restrict__ probs,
                                     const double* __restrict__ pi, const double* __restrict__ pa,
                                     const double* __restrict__ pb, const double* __restrict__ sa,
                                     const double* __restrict__ sb, double* __restrict__ out) {
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    const double* a = pa + (size_t)s * A;
    const double* b = pb + (size_t)s * A;
    double f[3] = {0.0, 0.0, 0.0};
    for (int d = 0; d <= nd; ++d) {
        const double* P = probs + (size_t)d * A * A;
        double acc = 0.0;
        for (int i = 0; i < A; ++i) {
            double x = 0.0;
            for (int j = 0; j < A; ++j) x = fma(P[i * A + j], a[j], x);
            acc = fma(x * b[i], pi[i], acc);
        }
        f[d] = acc;
    }
    double* o = out + (size_t)s * (nd + 1);
    o[0] = log(f[0]) + sa[s] + sb[s];
    if (nd == 2) {
        o[1] = f[1] / f[0];
        o[2] = (f[2] * f[0] - f[1] * f[1]) / (f[0] * f[0]);
    }
}

int pick_device(int device) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        set_thread_error(std::string("no CUDA device available (") + cudaGetErrorString(e) +
                         "); this engine has no CPU fallback");
        cudaGetLastError();
        return PHB_ERR_NO_DEVICE;
    }
    if (device < 0 || device >= n) {
        set_thread_error("device index out of range");
        return PHB_ERR_INVALID;
    }
    OP_CUDA(cudaSetDevice(device), "cudaSetDevice");
    return PHB_OK;
}

}  // namespace
}  // namespace phb

using namespace phb;

extern "C" {

int phb_op_clv(int device, int64_t S, int K, int A, const double* p1, const double* p2, const double* clv1,
               const double* clv2, const double* scaler_a, const double* scaler_b, double* cml_scaler, double* out) {
    if (S < 0 || K < 1 || A < 1 || !p1 || !p2 || (S > 0 && (!clv1 || !clv2 || !scaler_a || !scaler_b || !cml_scaler || !out))) {
        set_thread_error("phb_op_clv: bad argument");
        return PHB_ERR_INVALID;
    }
    if (S == 0) return PHB_OK;
    int st = pick_device(device);
    if (st) return st;
    const size_t nP = (size_t)K * A * A * 8, nL = (size_t)S * K * A * 8, nS = (size_t)S * K * 8;
    DevBuf dp1, dp2, d1, d2, dsa, dsb, dso, dout;
    OP_CUDA(dp1.alloc(nP), "phb_op_clv alloc");
    OP_CUDA(dp2.alloc(nP), "phb_op_clv alloc");
    OP_CUDA(d1.alloc(nL), "phb_op_clv alloc");
    OP_CUDA(d2.alloc(nL), "phb_op_clv alloc");
    OP_CUDA(dsa.alloc(nS), "phb_op_clv alloc");
    OP_CUDA(dsb.alloc(nS), "phb_op_clv alloc");
    OP_CUDA(dso.alloc(nS), "phb_op_clv alloc");
    OP_CUDA(dout.alloc(nL), "phb_op_clv alloc");
    OP_CUDA(cudaMemcpy(dp1.p, p1, nP, cudaMemcpyHostToDevice), "phb_op_clv h2d");
    OP_CUDA(cudaMemcpy(dp2.p, p2, nP, cudaMemcpyHostToDevice), "phb_op_clv h2d");
    OP_CUDA(cudaMemcpy(d1.p, clv1, nL, cudaMemcpyHostToDevice), "phb_op_clv h2d");
    OP_CUDA(cudaMemcpy(d2.p, clv2, nL, cudaMemcpyHostToDevice), "phb_op_clv h2d");
    OP_CUDA(cudaMemcpy(dsa.p, scaler_a, nS, cudaMemcpyHostToDevice), "phb_op_clv h2d");
    OP_CUDA(cudaMemcpy(dsb.p, scaler_b, nS, cudaMemcpyHostToDevice), "phb_op_clv h2d");
    const int threads = 128;
    const int64_t blocks = (S * K + threads - 1) / threads;
    op_clv_kernel<<<(unsigned)blocks, threads>>>(S, K, A, dp1.as<double>(), dp2.as<double>(), d1.as<double>(),
                                                 d2.as<double>(), dsa.as<double>(), dsb.as<double>(),
                                                 dso.as<double>(), dout.as<double>());
    OP_CUDA(cudaGetLastError(), "phb_op_clv launch");
    OP_CUDA(cudaMemcpy(out, dout.p, nL, cudaMemcpyDeviceToHost), "phb_op_clv d2h");
    OP_CUDA(cudaMemcpy(cml_scaler, dso.p, nS, cudaMemcpyDeviceToHost), "phb_op_clv d2h");
    return PHB_OK;
}

int phb_op_lnl_node(int device, int64_t S, int K, int A, const double* pi, const double* partials,
                    const double* scale, double* out) {
    if (S < 0 || K < 1 || A < 1 || !pi || (S > 0 && (!partials || !scale || !out))) {
        set_thread_error("phb_op_lnl_node: bad argument");
        return PHB_ERR_INVALID;
    }
    if (S == 0) return PHB_OK;
    int st = pick_device(device);
    if (st) return st;
    const size_t nL = (size_t)S * K * A * 8, nS = (size_t)S * K * 8;
    DevBuf dpi, dl, ds, dout;
    OP_CUDA(dpi.alloc((size_t)A * 8), "phb_op_lnl_node alloc");
    OP_CUDA(dl.alloc(nL), "phb_op_lnl_node alloc");
    OP_CUDA(ds.alloc(nS), "phb_op_lnl_node alloc");
    OP_CUDA(dout.alloc(nS), "phb_op_lnl_node alloc");
    OP_CUDA(cudaMemcpy(dpi.p, pi, (size_t)A * 8, cudaMemcpyHostToDevice), "phb_op_lnl_node h2d");
    OP_CUDA(cudaMemcpy(dl.p, partials, nL, cudaMemcpyHostToDevice), "phb_op_lnl_node h2d");
    OP_CUDA(cudaMemcpy(ds.p, scale, nS, cudaMemcpyHostToDevice), "phb_op_lnl_node h2d");
    const int threads = 128;
    const int64_t blocks = (S * K + threads - 1) / threads;
    op_lnl_node_kernel<<<(unsigned)blocks, threads>>>(S * K, A, dpi.as<double>(), dl.as<double>(), ds.as<double>(),
                                                      dout.as<double>());
    OP_CUDA(cudaGetLastError(), "phb_op_lnl_node launch");
    OP_CUDA(cudaMemcpy(out, dout.p, nS, cudaMemcpyDeviceToHost), "phb_op_lnl_node d2h");
    return PHB_OK;
}

int phb_op_lnl_branch(int device, int64_t S, int A, int n_derivs, const double* probs, const double* pi,
                      const double* partials_a, const double* partials_b, const double* scale_a,
                      const double* scale_b, double* out) {
    if (S < 0 || A < 1 || (n_derivs != 0 && n_derivs != 2) || !probs || !pi ||
        (S > 0 && (!partials_a || !partials_b || !scale_a || !scale_b || !out))) {
        set_thread_error("phb_op_lnl_branch: bad argument");
        return PHB_ERR_INVALID;
    }
    if (S == 0) return PHB_OK;
    int st = pick_device(device);
    if (st) return st;
    const size_t nP = (size_t)(n_derivs + 1) * A * A * 8, nL = (size_t)S * A * 8, nS = (size_t)S * 8;
    const size_t nO = (size_t)S * (n_derivs + 1) * 8;
    DevBuf dp, dpi, da, db, dsa, dsb, dout;
    OP_CUDA(dp.alloc(nP), "phb_op_lnl_branch alloc");
    OP_CUDA(dpi.alloc((size_t)A * 8), "phb_op_lnl_branch alloc");
    OP_CUDA(da.alloc(nL), "phb_op_lnl_branch alloc");
    OP_CUDA(db.alloc(nL), "phb_op_lnl_branch alloc");
    OP_CUDA(dsa.alloc(nS), "phb_op_lnl_branch alloc");
    OP_CUDA(dsb.alloc(nS), "phb_op_lnl_branch alloc");
    OP_CUDA(dout.alloc(nO), "phb_op_lnl_branch alloc");
    OP_CUDA(cudaMemcpy(dp.p, probs, nP, cudaMemcpyHostToDevice), "phb_op_lnl_branch h2d");
    OP_CUDA(cudaMemcpy(dpi.p, pi, (size_t)A * 8, cudaMemcpyHostToDevice), "phb_op_lnl_branch h2d");
    OP_CUDA(cudaMemcpy(da.p, partials_a, nL, cudaMemcpyHostToDevice), "phb_op_lnl_branch h2d");
    OP_CUDA(cudaMemcpy(db.p, partials_b, nL, cudaMemcpyHostToDevice), "phb_op_lnl_branch h2d");
    OP_CUDA(cudaMemcpy(dsa.p, scale_a, nS, cudaMemcpyHostToDevice), "phb_op_lnl_branch h2d");
    OP_CUDA(cudaMemcpy(dsb.p, scale_b, nS, cudaMemcpyHostToDevice), "phb_op_lnl_branch h2d");
    const int threads = 128;
    const int64_t blocks = (S + threads - 1) / threads;
    op_lnl_branch_kernel<<<(unsigned)blocks, threads>>>(S, A, n_derivs, dp.as<double>(), dpi.as<double>(),
                                                        da.as<double>(), db.as<double>(), dsa.as<double>(),
                                                        dsb.as<double>(), dout.as<double>());
    OP_CUDA(cudaGetLastError(), "phb_op_lnl_branch launch");
    OP_CUDA(cudaMemcpy(out, dout.p, nO, cudaMemcpyDeviceToHost), "phb_op_lnl_branch d2h");
    return PHB_OK;
}

int phb_op_pmatrices(int device, int A, int n, const double* evecs, const double* evals, const double* ivecs,
                     const double* times, int order, double* out) {
    if (A < 1 || A > 64 || n < 0 || order < 0 || order > 2 || !evecs || !evals || !ivecs || (n > 0 && (!times || !out))) {
        set_thread_error("phb_op_pmatrices: bad argument");
        return PHB_ERR_INVALID;
    }
    if (n == 0) return PHB_OK;
    int st = pick_device(device);
    if (st) return st;
    const size_t AA = (size_t)A * A * 8;
    DevBuf dv, dl, di, dt, dout;
    OP_CUDA(dv.alloc(AA), "phb_op_pmatrices alloc");
    OP_CUDA(dl.alloc((size_t)A * 8), "phb_op_pmatrices alloc");
    OP_CUDA(di.alloc(AA), "phb_op_pmatrices alloc");
    OP_CUDA(dt.alloc((size_t)n * 8), "phb_op_pmatrices alloc");
    OP_CUDA(dout.alloc((size_t)n * AA), "phb_op_pmatrices alloc");
    OP_CUDA(cudaMemcpy(dv.p, evecs, AA, cudaMemcpyHostToDevice), "phb_op_pmatrices h2d");
    OP_CUDA(cudaMemcpy(dl.p, evals, (size_t)A * 8, cudaMemcpyHostToDevice), "phb_op_pmatrices h2d");
    OP_CUDA(cudaMemcpy(di.p, ivecs, AA, cudaMemcpyHostToDevice), "phb_op_pmatrices h2d");
    OP_CUDA(cudaMemcpy(dt.p, times, (size_t)n * 8, cudaMemcpyHostToDevice), "phb_op_pmatrices h2d");
    OP_CUDA(launch_pmatrix_raw(nullptr, dv.as<double>(), dl.as<double>(), di.as<double>(), nullptr, dt.as<double>(),
                               dout.as<double>(), A, 1, n, order, 0),
            "phb_op_pmatrices launch");
    OP_CUDA(cudaMemcpy(out, dout.p, (size_t)n * AA, cudaMemcpyDeviceToHost), "phb_op_pmatrices d2h");
    return PHB_OK;
}

// Measured fp64 ceilings of this GPU, the denominators bench.py's roofline uses for the compute-bound kernels
// (MEASURED_PEAKS.json has HBM and bf16 figures only).  kind 0: independent DFMA chains on every SM (vector pipe);
// kind 1: independent DMMA m8n8k4 chains (fp64 tensor pipe).  Best of five launches, CUDA events.
int phb_op_fp64_peak(int device, int kind, double* tflops) {
    if (tflops == nullptr || kind < 0 || kind > 3) {
        set_thread_error("phb_op_fp64_peak: bad argument");
        return PHB_ERR_INVALID;
    }
    int st = pick_device(device);
    if (st) return st;
    cudaDeviceProp prop;
    OP_CUDA(cudaGetDeviceProperties(&prop, device), "phb_op_fp64_peak");
    // kinds 2 / 3 (tuning aid): the DMMA probe with only 8 / 16 warps per SM - the occupancy of the 61-state pruning
    // kernel and twice that - to see how many warps it takes to keep the fp64 tensor pipe busy
    const int blocks = prop.multiProcessorCount * (kind == 2 ? 1 : (kind == 3 ? 2 : 4)), threads = 256, iters = 1 << 13;
    DevBuf out;
    OP_CUDA(out.alloc((size_t)blocks * threads * 8), "phb_op_fp64_peak alloc");
    cudaEvent_t a, b;
    OP_CUDA(cudaEventCreate(&a), "phb_op_fp64_peak");
    OP_CUDA(cudaEventCreate(&b), "phb_op_fp64_peak");
    // flops per thread and iteration: 8 chains x 2 (DFMA), or 8 chains x 512 / 32 (one m8n8k4 per warp = 512 flops)
    const double flops = (double)blocks * threads * iters * (kind == 0 ? 16.0 : 8.0 * 512.0 / 32.0);
    float best = 0.f;
    for (int rep = 0; rep < 6; ++rep) {
        cudaEventRecord(a);
        if (kind == 0) fp64_fma_probe<<<blocks, threads, 0>>>(out.as<double>(), iters);
        else fp64_mma_probe<<<blocks, threads>>>(out.as<double>(), iters);
        cudaEventRecord(b);
        OP_CUDA(cudaEventSynchronize(b), "phb_op_fp64_peak run");
        float ms = 0.f;
        cudaEventElapsedTime(&ms, a, b);
        if (rep > 0 && (best == 0.f || ms < best)) best = ms;   // launch 0 is the warm-up
    }
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    OP_CUDA(cudaGetLastError(), "phb_op_fp64_peak launch");
    *tflops = flops / (best * 1e-3) / 1e12;
    return PHB_OK;
}

}  // extern "C"
