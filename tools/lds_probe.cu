// Shared-memory wavefront probe: how many data-pipe cycles does one warp-wide LDS.128 cost for a given address pattern?
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/lds_probe tools/lds_probe.cu && tools/lds_probe
//
// One CTA of 16 warps on one SM; every lane issues UNROLL independent 128-bit loads per iteration from
// base + offset[lane] (+ a per-load stride that keeps the bank pattern), the sum of everything loaded is kept so that
// nothing is optimised away.  cycles / (warps * loads) under saturation = wavefronts per instruction.
// The patterns are the ones the 4-state walks use or could use.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

constexpr int WARPS = 16, ITERS = 2000, UNROLL = 8, SMEM = 48 * 1024;

__global__ void probe(const int* __restrict__ offs, int stride, long long* cycles, double* sink) {
    extern __shared__ __align__(128) unsigned char smem[];
    for (int i = threadIdx.x; i < SMEM / 8; i += blockDim.x) reinterpret_cast<double*>(smem)[i] = 1e-3 * i;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const unsigned char* base = smem + offs[lane];
    unsigned acc = 0;   // one integer op per load: the loop must be bound by the data pipe, not by an arithmetic pipe
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
        unsigned v[UNROLL], w[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            // + a multiple of 128 bytes that changes with the iteration: same banks, and ptxas cannot hoist the load
            const unsigned a = (unsigned)__cvta_generic_to_shared(base + ((u * stride) & 4095) + ((it & 31) << 7));
            unsigned x, y, z;
            asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v[u]), "=r"(x), "=r"(y), "=r"(z) : "r"(a) : "memory");
            v[u] ^= x;   // all four words are used, or ptxas narrows the load
            w[u] = y ^ z;
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) acc ^= v[u] ^ w[u];
    }
    __syncthreads();
    const long long t1 = clock64();
    if (threadIdx.x == 0) *cycles = t1 - t0;
    sink[blockIdx.x * blockDim.x + threadIdx.x] = (double)acc;
}

int main() {
    int* d_offs;
    long long* d_cycles;
    double* d_sink;
    cudaMalloc(&d_offs, 32 * sizeof(int));
    cudaMalloc(&d_cycles, sizeof(long long));
    cudaMalloc(&d_sink, WARPS * 32 * sizeof(double));
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    // a fixed "random" assignment of the four unambiguous codes to 32 patterns (+ one gap)
    const int code32[32] = {0, 3, 1, 2, 2, 0, 3, 1, 1, 1, 0, 3, 2, 3, 0, 2, 3, 0, 2, 1, 4, 2, 1, 3, 0, 0, 3, 2, 1, 3, 2, 0};
    struct Pattern {
        const char* name;
        int stride;
        int (*off)(int lane, const int* code);
    };
    const Pattern pats[] = {
        {"full broadcast (today's P block read)", 16, [](int, const int*) { return 0; }},
        {"32 distinct 16-byte words, conflict free (parked operand read)", 512, [](int l, const int*) { return 16 * l; }},
        {"P block per category, k = lane & 3, 80-byte blocks (SYM)", 16, [](int l, const int*) { return (l & 3) * 80; }},
        {"P block per category, k = lane >> 3, 80-byte blocks (SYM)", 16, [](int l, const int*) { return (l >> 3) * 80; }},
        {"P block per category, k = lane & 3, 128-byte blocks", 16, [](int l, const int*) { return (l & 3) * 128; }},
        {"P block per category, k = lane >> 3, 128-byte blocks", 16, [](int l, const int*) { return (l >> 3) * 128; }},
        {"tip table today: row = code of the lane's pattern, 32-byte rows", 256, [](int l, const int* c) { return c[l] * 32; }},
        {"tip table, k = lane >> 3, [k][code] 32-byte rows, 8 patterns", 16, [](int l, const int* c) { return (l >> 3) * 256 + c[l & 7] * 32; }},
        {"tip table, k = lane & 3, [code][k] 128-byte rows, 8 patterns", 16, [](int l, const int* c) { return c[l >> 2] * 128 + (l & 3) * 32; }},
        {"tip table, k = lane & 3, [code][k] 144-byte rows, 8 patterns", 16, [](int l, const int* c) { return c[l >> 2] * 144 + (l & 3) * 32; }},
        {"tip table, k = lane >> 3, [code][k] 144-byte rows, 8 patterns", 16, [](int l, const int* c) { return c[l & 7] * 144 + (l >> 3) * 32; }},
        {"tip table, k = lane & 3, [code][k] 160-byte rows, 8 patterns", 16, [](int l, const int* c) { return c[l >> 2] * 160 + (l & 3) * 32; }},
        {"sanity: 32 lanes on one bank, 128 bytes apart (32-way conflict)", 0, [](int l, const int*) { return l * 128; }},
        {"sanity: 2-way conflict (lanes l and l + 16 share banks)", 0, [](int l, const int*) { return (l & 15) * 16 + (l >> 4) * 256; }},
        {"two distinct addresses, lane & 1", 16, [](int l, const int*) { return (l & 1) * 64; }},
        {"two distinct addresses, lane >> 4", 16, [](int l, const int*) { return (l >> 4) * 64; }},
        {"eight distinct addresses, lane & 7, conflict free", 0, [](int l, const int*) { return (l & 7) * 16; }},
        {"eight distinct addresses, lane >> 2, conflict free", 0, [](int l, const int*) { return (l >> 2) * 16; }},
    };
    for (const Pattern& p : pats) {
        int offs[32];
        for (int l = 0; l < 32; ++l) offs[l] = p.off(l, code32);
        cudaMemcpy(d_offs, offs, sizeof(offs), cudaMemcpyHostToDevice);
        long long best = 1ll << 60;
        for (int rep = 0; rep < 3; ++rep) {
            probe<<<1, 32 * WARPS, SMEM>>>(d_offs, p.stride, d_cycles, d_sink);
            long long cyc = 0;
            cudaMemcpy(&cyc, d_cycles, sizeof(cyc), cudaMemcpyDeviceToHost);
            if (cyc < best) best = cyc;
        }
        const cudaError_t err = cudaGetLastError();
        printf("{\"pattern\": \"%s\", \"cycles_per_lds128\": %.3f%s}\n", p.name, (double)best / ((double)WARPS * ITERS * UNROLL),
               err == cudaSuccess ? "" : ", \"error\": true");
    }
    return 0;
}
