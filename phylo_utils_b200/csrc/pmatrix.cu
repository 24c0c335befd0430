// Batched transition-probability matrices.
//
//   P[m][k] = V . diag( g(lambda) . exp(lambda . t_m . r_k) ) . V^-1
//
// for every branch length t_m (all edges of the schedule, plus the root edge) and every rate
// category r_k in ONE launch; g = 1 for P, lambda (or lambda r_k) for dP/dt, its square for
// d2P/dt2.  This is the device form of Eigen.exp / Eigen.fn_apply behind Model.p, dp_dt, d2p_dt2
// (/root/reference/phylo_utils/substitution_models/abstract.py:49-77, 99-122), which the reference
// calls 2(N-2) times per evaluation from Python (tree_model.py:168-169).
//
// The matrices are tiny (A = 4, 20, 61): one CTA per (matrix, category), the scaled eigenvector
// matrix staged in shared memory, one output element per thread-iteration.  Traffic is KBs; the
// kernel only has to stay off the critical path of the pruning kernels that follow it in-stream.
#include "common.cuh"

namespace phb {

__global__ void pmatrix_kernel(const double* __restrict__ evecs, const double* __restrict__ evals,
                               const double* __restrict__ ivecs, const double* __restrict__ rates,
                               const double* __restrict__ lengths, double* __restrict__ out, int A, int K,
                               int order, int chain_rule) {
    extern __shared__ double sm[];
    double* scaledV = sm;          // [A][A]  V[i][m] * g_m * exp(lambda_m s)
    double* fac = sm + A * A;      // [A]
    const int m = blockIdx.x, k = blockIdx.y;
    const double r = rates ? rates[k] : 1.0;
    const double s = lengths[m] * r;  // same association as the reference: exp(evals * (t * rate))
    for (int j = threadIdx.x; j < A; j += blockDim.x) {
        const double lam = evals[j];
        double g = 1.0;
        const double base = chain_rule ? lam * r : lam;
        if (order >= 1) g = base;
        if (order >= 2) g = base * base;
        fac[j] = g * exp(lam * s);
    }
    __syncthreads();
    for (int e = threadIdx.x; e < A * A; e += blockDim.x) scaledV[e] = evecs[e] * fac[e % A];
    __syncthreads();
    double* dst = out + ((size_t)m * K + k) * A * A;
    for (int e = threadIdx.x; e < A * A; e += blockDim.x) {
        const int i = e / A, j = e % A;
        double acc = 0.0;
        for (int q = 0; q < A; ++q) acc += scaledV[i * A + q] * ivecs[q * A + j];
        dst[e] = acc;
    }
}

cudaError_t launch_pmatrix_raw(cudaStream_t stream, const double* evecs, const double* evals, const double* ivecs,
                               const double* rates, const double* d_lengths, double* d_out, int A, int K, int n_mats,
                               int order, int chain_rule) {
    if (n_mats <= 0) return cudaSuccess;
    const int threads = A * A >= 256 ? 256 : (A * A >= 64 ? 128 : 32);
    const size_t smem = (size_t)(A * A + A) * sizeof(double);
    dim3 grid(n_mats, K);
    pmatrix_kernel<<<grid, threads, smem, stream>>>(evecs, evals, ivecs, rates, d_lengths, d_out, A, K, order,
                                                    chain_rule);
    return cudaGetLastError();
}

// T[m][k][code][i] = sum_j P[m][k][i][j] . lut[code][j]   (4-state models; one CTA per P block m)
__global__ void tip_table_kernel(const double* __restrict__ pmats, const double* __restrict__ lut, int n_codes, int nc,
                                 int K, int first_mat, double* __restrict__ out) {
    const int m = first_mat + blockIdx.x;
    for (int idx = threadIdx.x; idx < K * nc * 4; idx += blockDim.x) {
        const int i = idx & 3, code = (idx >> 2) % nc, k = idx / (4 * nc);
        double acc = 0.0;
        if (code < n_codes) {
            const double* P = pmats + ((size_t)m * K + k) * 16 + i * 4;
            const double* v = lut + code * 4;
            acc = P[0] * v[0];
            acc = fma(P[1], v[1], acc);
            acc = fma(P[2], v[2], acc);
            acc = fma(P[3], v[3], acc);
        }
        out[((size_t)m * K + k) * nc * 4 + code * 4 + i] = acc;
    }
}

// R[m][k] = upper triangle of diag(pi) P[m][k], packed [r00 r01 r02 r03 r11 r12 r13 r22 r23 r33] (4-state reversible models)
__global__ void sym_block_kernel(const double* __restrict__ pmats, const double* __restrict__ freqs, int K, int first_mat,
                                 double* __restrict__ out) {
    const int m = first_mat + blockIdx.x;
    for (int idx = threadIdx.x; idx < K * 10; idx += blockDim.x) {
        const int k = idx / 10, e = idx - 10 * k;
        const int i = e < 4 ? 0 : (e < 7 ? 1 : (e < 9 ? 2 : 3));
        const int j = e < 4 ? e : (e < 7 ? e - 3 : (e < 9 ? e - 5 : 3));
        out[((size_t)m * K + k) * 10 + e] = freqs[i] * pmats[((size_t)m * K + k) * 16 + i * 4 + j];
    }
}

// any state count: T[m][k][code][i], grid (n_mats, K)
__global__ void tip_table_generic_kernel(const double* __restrict__ pmats, const double* __restrict__ lut, int n_codes,
                                         int nc, int A, int K, int first_mat, double* __restrict__ out) {
    const int m = first_mat + blockIdx.x, k = blockIdx.y;
    const double* P = pmats + ((size_t)m * K + k) * A * A;
    double* dst = out + ((size_t)m * K + k) * nc * A;
    for (int idx = threadIdx.x; idx < nc * A; idx += blockDim.x) {
        const int code = idx / A, i = idx - code * A;
        double acc = 0.0;
        if (code < n_codes)
            for (int j = 0; j < A; ++j) acc = fma(P[i * A + j], lut[code * A + j], acc);
        dst[idx] = acc;
    }
}

int launch_tip_tables(Ctx* c, int first_mat, int n_mats) {
    if (n_mats <= 0) return PHB_OK;
    if (c->A != 4) {
        if (tip_tables_usable(c)) {
            dim3 grid(n_mats, c->K);
            tip_table_generic_kernel<<<grid, 256, 0, c->stream>>>(c->d_pmats, c->d_lut, c->n_codes, tip_table_rows(c), c->A,
                                                                  c->K, first_mat, c->d_tiptab);
            c->launches++;
            PHB_CUDA(c, cudaGetLastError());
        }
        return launch_mma_images(c, first_mat, n_mats);   // padded staging images for the DMMA kernels (61 states)
    }
    if (c->d_rmats != nullptr && c->reversible) {
        sym_block_kernel<<<n_mats, 64, 0, c->stream>>>(c->d_pmats, c->model_freqs(), c->K, first_mat, c->d_rmats);
        c->launches++;
        PHB_CUDA(c, cudaGetLastError());
    }
    if (!tip_tables_usable(c)) return PHB_OK;
    tip_table_kernel<<<n_mats, 128, 0, c->stream>>>(c->d_pmats, c->d_lut, c->n_codes, tip_table_rows(c), c->K, first_mat,
                                                    c->d_tiptab);
    c->launches++;
    PHB_CUDA(c, cudaGetLastError());
    return PHB_OK;
}

int launch_build_pmatrices(Ctx* c, const double* d_lengths, int n_mats, double* d_out, int order, int chain_rule) {
    if (n_mats <= 0) return PHB_OK;
    c->launches++;
    PHB_CUDA(c, launch_pmatrix_raw(c->stream, c->model_evecs(), c->model_evals(), c->model_ivecs(), c->model_rates(),
                                   d_lengths, d_out, c->A, c->K, n_mats, order, chain_rule));
    return PHB_OK;
}

}  // namespace phb
