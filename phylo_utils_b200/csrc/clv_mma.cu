// Felsenstein pruning on the FP64 tensor cores (DMMA) for the two large alphabets:
// amino acids (A = 20) and sense codons (A = 61).
//
// With 20 or 61 states, P . L is a real dense contraction (BASELINE north_star, subsystem 3): for one
// category and one child
//
//     X[i][s] = sum_j P[i][j] . L[s][j]              i, j = states, s = pattern
//
// is a (A x A) . (A x N) product over a tile of N patterns.  It is issued as
// mma.sync.aligned.m8n8k4.row.col.f64 (SASS DMMA; tcgen05 has no FP64 kind):
//     A operand = P       (row-major, 8 states x 4 states per instruction)
//     B operand = L^T     ("column-major": the 4 states of one pattern are contiguous)
//     C/D       = X       (8 states x 8 patterns, two doubles per lane)
// States are padded to MT*8 rows and KS*4 columns with zeros (20 -> 24 x 20, 61 -> 64 x 64).
//
// A CTA owns a tile of patterns; each warp owns NT * 8 of them and keeps both children's accumulators in
// registers, multiplies them element-wise (same fragment layout), finds the per-pattern maximum with three
// shuffles, and writes the result through its own rows of the shared child tile back to global memory with
// coalesced stores.  P[k] for the two children is staged in shared memory once per (row, category) and
// shared by all warps; all shared arrays use a row pitch = 4 (mod 16) doubles so that the 8 x 4 fragment
// loads are bank-conflict free.  Semantics (reference `clv`, numba_likelihood_engine.py:10-46, with the
// per-pattern binary exponent of clv_dna.cu) and data layout are those of clv_generic.cu, which remains the
// fallback for every other state count.
#include <cstdlib>

#include "common.cuh"

namespace phb {

namespace {

// A row as the kernel runs it: up to three operands, up to two outputs.
//   n_ops == 2: dst[0] = op0 * op1                        (a pruning row, what an OpRow says)
//   n_ops == 3: dst[0] = op0 * op1,  dst[1] = op0 * op2    (a pre-order parent step: op0 = what sits above the parent,
//               op1 / op2 = its children; three products where two separate rows take four, X staged once)
struct MmaRow {
    int32_t src[3], kind[3], pidx[3];
    int32_t dst[2];
    int32_t n_ops;
};
static_assert(sizeof(MmaRow) == 48, "MmaRow must stay 48 bytes");

struct MmaArgs {
    const OpRow* rows;
    const MmaRow* frows;  // non-null: the launch runs these rows instead of `rows`
    int row_begin, row_end;
    const double* pmats;  // [pidx][K][A][A]
    const double* tiptab; // [pidx][K][nc][A] = P . lut[code], or null
    // odd state counts (61): zero-padded staging images [pidx][K][2][MROWS][LDP] of the P blocks (0) and the tip
    // tables (1), 16-byte aligned and contiguous, so that a block is staged by a flat 16-byte copy; or null
    const double* pimg;
    int nc;
    const uint8_t* codes;
    size_t pitch;
    const double* lut;    // [256][A]
    double* clv;
    int32_t* scale;
    int64_t S;
    int64_t n_tiles;
    int A, K;
};

__device__ __forceinline__ void dmma(double (&d)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d[0]), "+d"(d[1])
                 : "d"(a), "d"(b));
}

constexpr int pad_pitch(int cols) {   // smallest pitch >= cols with pitch % 16 == 4
    int p = cols;
    while (p % 16 != 4) ++p;
    return p;
}

__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gmem_src) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit_all() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// A row is processed as 2K "phases" (category k, child c).  While the warps issue the DMMAs of phase ph out of
// one P buffer and one child-row buffer, the P block and the child rows of phase ph+1 stream into the other
// pair with cp.async (8-byte pieces: rows of 20 or 61 doubles are only 8-byte aligned).  Padding rows and
// columns are zeroed once and never touched again.
// KK: the category count when it is known at compile time (4: every configuration BASELINE.json names), 0 = p.K.  With it
// the byte distance between two patterns' rows of a block, K * A * 8, is a constant, and the per-row copies below become
// one instruction with an immediate offset each; as a run-time value every copy paid a 64-bit multiply-add chain (the
// copy loops were 37 % of the 20-state kernel's instructions: profiles/r02i_mma_sass_profile.txt).
template <int AA, int MT, int KS, int NT, int WARPS, bool LEVEL, bool FUSED, int KK = 0>
__global__ void __launch_bounds__(WARPS * 32) mma_prune_kernel(const MmaArgs p) {
    constexpr int MROWS = MT * 8, KCOLS = KS * 4;
    constexpr int LDP = pad_pitch(KCOLS);
    constexpr int LDL = pad_pitch(KCOLS);             // child / output rows hold A <= KCOLS states
    constexpr int TS = WARPS * NT * 8, WR = NT * 8;   // patterns per CTA / per warp
    extern __shared__ double sm[];
    double* Pbuf = sm;                               // [2][MROWS][LDP]
    double* Lbuf = Pbuf + 2 * MROWS * LDP;           // [2][TS][LDL]
    unsigned char* s_codes = reinterpret_cast<unsigned char*>(Lbuf + 2 * TS * LDL);   // [3 operands][TS]
    int* s_exp = reinterpret_cast<int*>(s_codes + 4 * TS);   // [3 operands][TS] exponents of the operands' patterns
    int* s_shift = s_exp + 3 * TS;                           // [TS] binary shift a pattern of the output is rescaled by
    static_assert(WR <= 32, "one pattern per lane in the per-pattern bookkeeping");
    constexpr int A = AA;   // compile-time: the staging loops divide by it
    const int K = KK != 0 ? KK : p.K;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int fr = lane >> 2, fc = lane & 3;          // fragment row / column
    const size_t S = (size_t)p.S;
    for (int e = threadIdx.x; e < 2 * MROWS * LDP + 2 * TS * LDL; e += WARPS * 32) sm[e] = 0.0;
    __syncthreads();
    // Rows of an odd number of doubles (61 states) start on a 16-byte boundary only every second row.  With an even
    // category count every row of a phase shares the parity of k: shifted by one double in shared memory (element j of
    // a row sits in column j + sh, sh = k & 1), both ends of every copy are 16-byte aligned, and a row moves as 30
    // 16-byte pieces and one 8-byte piece - lane = piece, no index arithmetic - instead of 61 8-byte pieces that each
    // cost a division (profiles/r02a_cfg4_mma.txt: 69 % of the kernel's instructions were in the three copy loops).
    constexpr bool ODD = (A % 2) == 1;
    constexpr int HALF = (A - 1) / 2;                 // 16-byte pieces of an odd row
    static_assert(!ODD || (HALF + 2 <= 32 && A + 1 <= LDL), "one piece per lane; room for the shift");
    const bool fast = ODD && (p.K % 2 == 0) && p.pimg != nullptr;
    // A row = HALF 16-byte pieces (lane = piece, lanes 0 .. HALF-1, first element 2 lane + sh) + ONE 8-byte piece
    // (element A-1 of an unshifted row, element 0 of a shifted one).  The 8-byte pieces of a warp's rows are moved by one
    // instruction, lane = row - not by one lane walking the rows alone while 31 wait.
    static_assert(!ODD || NT * 8 <= 16, "lane = row for the 8-byte pieces (rows 0..15), lane - 16 = row for the zeroing");

    // a P block or tip table -> the staging buffer: rows of A doubles, contiguous in shared memory too when LDP == A
    // (20 states: one 16-byte copy per two doubles instead of two 8-byte ones)
    auto stage_image = [&](double* Pd, const double* img) {
        for (int e = threadIdx.x; e < MROWS * LDP / 2; e += WARPS * 32) {
            const unsigned d = (unsigned)__cvta_generic_to_shared(Pd + 2 * e);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(img + 2 * e) : "memory");
        }
    };
    auto stage_matrix = [&](double* Pd, const double* q, int n_rows) {
        if (LDP == A && A % 2 == 0) {
            for (int e = threadIdx.x; e < n_rows * A / 2; e += WARPS * 32) {
                const unsigned d = (unsigned)__cvta_generic_to_shared(Pd + 2 * e);
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(q + 2 * e) : "memory");
            }
        } else {
            for (int e = threadIdx.x; e < n_rows * A; e += WARPS * 32) {
                const int i = e / A, j = e - i * A;
                cp_async8(Pd + i * LDP + j, q + e);
            }
        }
    };

    const int64_t items = LEVEL ? (int64_t)(p.row_end - p.row_begin) * p.n_tiles : p.n_tiles;
    for (int64_t it = blockIdx.x; it < items; it += gridDim.x) {
        const int64_t tile = LEVEL ? it % p.n_tiles : it;
        const int r0 = LEVEL ? p.row_begin + (int)(it / p.n_tiles) : p.row_begin;
        const int r1 = LEVEL ? r0 + 1 : p.row_end;
        const int64_t wsite0 = tile * TS + (int64_t)warp * WR;   // first pattern of this warp
        // the descriptor of row r + 1 is fetched while row r runs: its latency, and that of the codes / exponents that
        // depend on it, no longer sit in front of every row
        auto load_row = [&](int r) {
            MmaRow d;
            if (FUSED) {
                d = p.frows[r];
            } else {
                const OpRow o = p.rows[r];
                for (int c = 0; c < 2; ++c) {
                    d.src[c] = o.src[c];
                    d.kind[c] = o.kind[c];
                    d.pidx[c] = o.pidx[c];
                }
                d.src[2] = d.pidx[2] = 0;
                d.kind[2] = SRC_TIP;
                d.dst[0] = d.dst[1] = o.dst;
                d.n_ops = 2;
            }
            return d;
        };
        MmaRow next_row = load_row(r0);
        for (int r = r0; r < r1; ++r) {
            const MmaRow row = next_row;
            if (r + 1 < r1) next_row = load_row(r + 1);
            const int n_ops = FUSED ? 3 : 2;
            // selects instead of indexed reads: the row stays in registers
            auto src_of = [&](int c) { return c == 0 ? row.src[0] : (c == 1 ? row.src[1] : row.src[2]); };
            auto kind_of = [&](int c) { return c == 0 ? row.kind[0] : (c == 1 ? row.kind[1] : row.kind[2]); };
            auto pidx_of = [&](int c) { return c == 0 ? row.pidx[0] : (c == 1 ? row.pidx[1] : row.pidx[2]); };
            // Pruning rows: a tip operand's contribution (row `code` of its P.lut table) is gathered STRAIGHT from the
            // table in global memory into the accumulator registers - no staging copy, no barrier, no phase of its own;
            // the loads fly during the product phase of the other operand.  Only internal operands have phases:
            // 2K for an internal x internal row, K for a tip x internal row, none for a cherry (half of all operands of
            // a tree are tips: the phase count, and with it the barriers and the table staging traffic, is halved).
            // (61 states, where a table is 30 kB; at 20 states staging 3.8 kB costs less than the exposed L2 latency of the
            // gathers: cfg3 15.98 vs 16.42 ms)
            const bool d0 = !FUSED && A > 32 && kind_of(0) == SRC_TIP && p.tiptab != nullptr;
            const bool d1 = !FUSED && A > 32 && kind_of(1) == SRC_TIP && p.tiptab != nullptr;
            const int n_phases = FUSED ? 3 * K : ((d0 ? 0 : 1) + (d1 ? 0 : 1)) * K;
            auto next_phase = [&](int k, int c, int& kn, int& cn) {
                if (FUSED || (!d0 && !d1)) {
                    cn = c + 1 == n_ops ? 0 : c + 1;
                    kn = cn == 0 ? k + 1 : k;
                } else {          // one internal operand: the same operand of the next category
                    cn = c;
                    kn = k + 1;
                }
            };

            // tip codes and exponents of this warp's patterns (one per lane), same for every category.  The loads are
            // issued here and parked in shared memory once the first operand copies are under way: one latency, not two.
            const bool lane_ok = lane < WR && wsite0 + lane < p.S;
            int code_reg[3] = {0, 0, 0}, exp_reg[3] = {0, 0, 0};
#pragma unroll
            for (int c = 0; c < 3; ++c)
                if (c < n_ops && lane_ok) {
                    if (kind_of(c) == SRC_TIP) code_reg[c] = p.codes[(size_t)src_of(c) * p.pitch + wsite0 + lane];
                    else exp_reg[c] = p.scale[(size_t)src_of(c) * S + wsite0 + lane];
                }
            auto park_codes = [&]() {
                if (lane < WR) {
#pragma unroll
                    for (int c = 0; c < 3; ++c)
                        if (c < n_ops) {
                            s_codes[c * TS + warp * WR + lane] = (unsigned char)code_reg[c];
                            s_exp[c * TS + warp * WR + lane] = exp_reg[c];
                        }
                }
                __syncwarp();
            };

            // phase ph = (category k, operand c), in the order (k, 0), (k, 1) [, (k, 2)]; buffers alternate with ph
            auto prefetch = [&](int ph, int k, int c) {
                const int buf = ph & 1;
                double* Pd = Pbuf + (size_t)buf * MROWS * LDP;
                const int sh = fast ? (k & 1) : 0;
                if (kind_of(c) == SRC_TIP && p.tiptab != nullptr) {
                    // a tip child needs no product at all: its contribution is row `code` of T = P . lut, staged
                    // where the P block would go (row = code, n_codes <= MROWS rows)
                    if (fast) stage_image(Pd, p.pimg + (((size_t)pidx_of(c) * K + k) * 2 + 1) * (MROWS * LDP));
                    else stage_matrix(Pd, p.tiptab + ((size_t)pidx_of(c) * K + k) * p.nc * A, p.nc);
                    cp_async_commit_all();
                    return;
                }
                if (fast) stage_image(Pd, p.pimg + (((size_t)pidx_of(c) * K + k) * 2) * (MROWS * LDP));
                else stage_matrix(Pd, p.pmats + ((size_t)pidx_of(c) * K + k) * A * A, A);
                double* Ld = Lbuf + ((size_t)buf * TS + (size_t)warp * WR) * LDL;
                if (kind_of(c) == SRC_TIP) {
                    for (int e = lane; e < WR * A; e += 32) {
                        const int n = e / A, j = e - n * A;
                        Ld[n * LDL + j + sh] = __ldg(p.lut + (size_t)s_codes[c * TS + warp * WR + n] * A + j);
                    }
                    if (fast && sh == 0 && lane < WR) Ld[lane * LDL + A] = 0.0;   // column A may hold element A-1 of a shifted row
                } else if (fast) {
                    const int n_valid = (int)min((int64_t)WR, p.S - wsite0);
                    const double* g = p.clv + (((size_t)src_of(c) * S + wsite0) * K + k) * A;
                    if (lane < HALF) {
                        const int j0 = 2 * lane + sh;
                        const unsigned sdst = (unsigned)__cvta_generic_to_shared(Ld + j0 + sh);
                        const double* q = g + j0;
                        asm volatile("" : "+l"(q));   // one base register: ptxas otherwise rebuilds the address from the kernel parameters inside every predicated copy
#pragma unroll
                        for (int n = 0; n < WR; ++n)
                            if (n < n_valid)
                                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sdst + n * (LDL * 8)), "l"(q + (size_t)n * (K * A)) : "memory");
                    }
                    if (lane < n_valid) {                     // the 8-byte piece of row `lane`
                        const int j8 = sh == 0 ? A - 1 : 0;
                        cp_async8(Ld + lane * LDL + j8 + sh, g + (size_t)lane * (K * A) + j8);
                    } else if (sh == 0 && lane >= 16 && lane - 16 < n_valid) {
                        Ld[(lane - 16) * LDL + A] = 0.0;      // column A may hold element A-1 of a shifted row
                    }
                } else {
                    // rows of A doubles are 16-byte aligned when A is even (A = 20: ten 16-byte pieces per row)
                    constexpr int PB = (A % 2 == 0) ? 16 : 8, PIECES = A * 8 / PB;
                    const int n_valid = (int)min((int64_t)WR, p.S - wsite0);
                    const unsigned char* g =
                        reinterpret_cast<const unsigned char*>(p.clv + (((size_t)src_of(c) * S + wsite0) * K + k) * A);
                    const unsigned sdst = (unsigned)__cvta_generic_to_shared(Ld);
                    if (PB == 16 && PIECES <= 16) {
                        // a lane keeps its piece and walks the rows (32 / PIECES rows per trip): no index division per copy
                        constexpr int RPT = 32 / PIECES > 0 ? 32 / PIECES : 1, TRIPS = (WR + RPT - 1) / RPT;   // (PIECES > 32: this branch is not taken)
                        const int r0 = lane / PIECES, piece = lane - r0 * PIECES;
                        if (r0 < RPT) {
                            const unsigned d0 = sdst + r0 * (LDL * 8) + piece * 16;
                            const unsigned char* q0 = g + (size_t)r0 * (K * A * 8) + piece * 16;
                            asm volatile("" : "+l"(q0));
#pragma unroll
                            for (int i = 0; i < TRIPS; ++i)
                                if (r0 + RPT * i < n_valid)
                                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d0 + i * (RPT * LDL * 8)),
                                                 "l"(q0 + (size_t)i * (RPT * K * A * 8)) : "memory");
                        }
                    } else {
                        for (int e = lane; e < n_valid * PIECES; e += 32) {
                            const int n = e / PIECES, piece = e - n * PIECES;
                            const unsigned d = sdst + n * (LDL * 8) + piece * PB;
                            const unsigned char* q = g + (size_t)n * (K * A * 8) + piece * PB;
                            if (PB == 16) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(q) : "memory");
                            else asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(q) : "memory");
                        }
                    }
                }
                cp_async_commit_all();
            };

            int mx[2][NT][2];   // per output: high word of the per-pattern maximum (partials are >= 0: the high word orders them)
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) mx[0][nt][0] = mx[0][nt][1] = mx[1][nt][0] = mx[1][nt][1] = 0;
            double acc0[MT][NT][2], acc1[MT][NT][2];
            // DMMAs of one phase: acc = P[buf] . (this warp's child rows in buf)^T
            auto run_phase = [&](int ph, int k, int c, double (&acc)[MT][NT][2]) {
                const int buf = ph & 1;
                cp_async_wait_all();
                __syncthreads();      // phase ph's operands have landed; everybody has left phase ph-1
                if (ph + 1 < n_phases) {
                    int kn, cn;
                    next_phase(k, c, kn, cn);
                    prefetch(ph + 1, kn, cn);
                }
                const double* Pd = Pbuf + (size_t)buf * MROWS * LDP;
                const double* myLr = Lbuf + ((size_t)buf * TS + (size_t)warp * WR) * LDL + (fast ? (k & 1) : 0);
                if (kind_of(c) == SRC_TIP && p.tiptab != nullptr) {
                    // gather in fragment layout: acc[mt][nt][q] = T[code(pattern nt*8 + 2fc + q)][state mt*8 + fr]
#pragma unroll
                    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                        for (int q = 0; q < 2; ++q) {
                            const int code = s_codes[c * TS + warp * WR + nt * 8 + 2 * fc + q];
#pragma unroll
                            for (int mt = 0; mt < MT; ++mt) acc[mt][nt][q] = Pd[code * LDP + mt * 8 + fr];
                        }
                    return;
                }
#pragma unroll
                for (int mt = 0; mt < MT; ++mt)
#pragma unroll
                    for (int nt = 0; nt < NT; ++nt) acc[mt][nt][0] = acc[mt][nt][1] = 0.0;
#pragma unroll 4
                for (int ks = 0; ks < KS; ++ks) {
                    double bf[NT];
#pragma unroll
                    for (int nt = 0; nt < NT; ++nt) bf[nt] = myLr[(nt * 8 + fr) * LDL + ks * 4 + fc];
#pragma unroll
                    for (int mt = 0; mt < MT; ++mt) {
                        const double af = Pd[(mt * 8 + fr) * LDP + ks * 4 + fc];
#pragma unroll
                        for (int nt = 0; nt < NT; ++nt) dmma(acc[mt][nt], af, bf[nt]);
                    }
                }
            };

            // a tip operand's accumulators straight from its table: acc[mt][nt][q] = T[code(pattern nt*8 + 2fc + q)][mt*8 + fr]
            auto gather_direct = [&](int k, int c, double (&acc)[MT][NT][2]) {
                const double* T = p.tiptab + ((size_t)pidx_of(c) * K + k) * p.nc * A + fr;
#pragma unroll
                for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                    for (int q = 0; q < 2; ++q) {
                        const double* trow = T + (size_t)s_codes[c * TS + warp * WR + nt * 8 + 2 * fc + q] * A;
#pragma unroll
                        for (int mt = 0; mt < MT; ++mt)
                            acc[mt][nt][q] = (mt * 8 + 8 <= A || mt * 8 + fr < A) ? __ldg(trow + mt * 8) : 0.0;
                    }
            };
            // output j-1 of category k = acc0 * acc1.  It leaves through this warp's rows of L buffer `sbuf`: the one the
            // last phase read (the next copy into it is issued behind a barrier), or any when the row has no phases
            auto emit = [&](int j, int k, int sbuf) {
                {
                    double* out = p.clv + (size_t)(j == 1 ? row.dst[0] : row.dst[1]) * S * K * A;
                    double* myL = Lbuf + ((size_t)sbuf * TS + (size_t)warp * WR) * LDL;
                    __syncwarp();      // all lanes have read their operand rows; they now become the output rows
                    // fragment element (mt, nt, q) = state mt*8 + fr of pattern nt*8 + 2fc + q.  Padding states (>= A) are
                    // neither stored nor allowed into the maximum: a tip-table gather reads past its row for them.
                    const int sh = fast ? (k & 1) : 0;        // the output rows have the parity of their category too
                    double* corner = myL + (2 * fc) * LDL + fr + sh;
#pragma unroll
                    for (int mt = 0; mt < MT; ++mt) {
                        if (mt * 8 + 8 > A && mt * 8 + fr >= A) continue;
#pragma unroll
                        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                            for (int q = 0; q < 2; ++q) {
                                const double o = acc0[mt][nt][q] * acc1[mt][nt][q];
                                corner[(nt * 8 + q) * LDL + mt * 8] = o;
                                if (j == 1) mx[0][nt][q] = max(mx[0][nt][q], __double2hiint(o));
                                else mx[1][nt][q] = max(mx[1][nt][q], __double2hiint(o));
                            }
                    }
                    __syncwarp();
                    if (fast) {
                        // lane = piece, as on the way in: 16-byte loads from the shifted rows, 16-byte stores
                        const int n_valid = (int)min((int64_t)WR, p.S - wsite0);
                        double* g = out + ((size_t)wsite0 * K + k) * A;
                        if (lane < HALF) {
                            const int j0 = 2 * lane + sh;
                            double* gq = g + j0;
                            const double* sq = myL + j0 + sh;
                            asm volatile("" : "+l"(gq));
#pragma unroll
                            for (int n = 0; n < WR; ++n)
                                if (n < n_valid)
                                    *reinterpret_cast<double2*>(gq + (size_t)n * (K * A)) = *reinterpret_cast<const double2*>(sq + n * LDL);
                        }
                        if (lane < n_valid) {                 // the 8-byte piece of row `lane`
                            const int j8 = sh == 0 ? A - 1 : 0;
                            g[(size_t)lane * (K * A) + j8] = myL[lane * LDL + j8 + sh];
                        }
                    } else {
                        // coalesced: A contiguous doubles per pattern, in 16-byte pieces when A is even.  Columns >= A of
                        // the rows are never written (they stay zero for their next life as an operand row).
                        constexpr int PB = (A % 2 == 0) ? 16 : 8, PIECES = A * 8 / PB;
                        const int n_valid = (int)min((int64_t)WR, p.S - wsite0);
                        unsigned char* g = reinterpret_cast<unsigned char*>(out + ((size_t)wsite0 * K + k) * A);
                        const unsigned char* src = reinterpret_cast<const unsigned char*>(myL);
                        if (PB == 16 && PIECES <= 16) {
                            constexpr int RPT = 32 / PIECES > 0 ? 32 / PIECES : 1, TRIPS = (WR + RPT - 1) / RPT;   // (PIECES > 32: this branch is not taken)
                            const int r0 = lane / PIECES, piece = lane - r0 * PIECES;
                            if (r0 < RPT) {
                                unsigned char* g0 = g + (size_t)r0 * (K * A * 8) + piece * 16;
                                const unsigned char* s0 = src + r0 * (LDL * 8) + piece * 16;
                                asm volatile("" : "+l"(g0));
#pragma unroll
                                for (int i = 0; i < TRIPS; ++i)
                                    if (r0 + RPT * i < n_valid)
                                        *reinterpret_cast<int4*>(g0 + (size_t)i * (RPT * K * A * 8)) =
                                            *reinterpret_cast<const int4*>(s0 + i * (RPT * LDL * 8));
                            }
                        } else {
                            for (int e = lane; e < n_valid * PIECES; e += 32) {
                                const int n = e / PIECES, piece = e - n * PIECES;
                                if (PB == 16)
                                    *reinterpret_cast<int4*>(g + (size_t)n * (K * A * 8) + piece * 16) =
                                        *reinterpret_cast<const int4*>(src + n * (LDL * 8) + piece * 16);
                                else
                                    *reinterpret_cast<double*>(g + (size_t)n * (K * A * 8) + piece * 8) =
                                        *reinterpret_cast<const double*>(src + n * (LDL * 8) + piece * 8);
                            }
                        }
                    }
                    __syncwarp();
                }
            };

            __syncthreads();          // the previous row is completely done with both buffer pairs
            if (p.tiptab == nullptr) park_codes();   // without tip tables the first copy gathers look-up rows by code
            if (n_phases > 0) prefetch(0, 0, d0 ? 1 : 0);
            if (p.tiptab != nullptr) park_codes();
            int ph = 0, sbuf = 0;
            for (int k = 0; k < K; ++k) {
                if (FUSED) {
                    run_phase(ph, k, 0, acc0);
                    ++ph;
                    for (int j = 1; j < n_ops; ++j, ++ph) {
                        run_phase(ph, k, j, acc1);
                        emit(j, k, ph & 1);
                    }
                } else {
                    if (d0) gather_direct(k, 0, acc0);       // in flight while the other operand's products run
                    if (d1) gather_direct(k, 1, acc1);
                    if (!d0) {
                        run_phase(ph, k, 0, acc0);
                        sbuf = ph & 1;
                        ++ph;
                    }
                    if (!d1) {
                        run_phase(ph, k, 1, acc1);
                        sbuf = ph & 1;
                        ++ph;
                    }
                    emit(1, k, sbuf);
                }
            }
            // per-pattern maximum over states (lanes sharing fc) and categories (already folded into mx)
            for (int j = 0; j + 1 < n_ops; ++j) {
#pragma unroll
                for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                    for (int q = 0; q < 2; ++q) {
                        int m = j == 0 ? mx[0][nt][q] : mx[1][nt][q];
                        m = max(m, __shfl_xor_sync(0xffffffffu, m, 4));
                        m = max(m, __shfl_xor_sync(0xffffffffu, m, 8));
                        m = max(m, __shfl_xor_sync(0xffffffffu, m, 16));
                        if (j == 0) mx[0][nt][q] = m;
                        else mx[1][nt][q] = m;
                    }
                // lanes 0..3 (fr == 0) hold the maxima of the patterns 2*fc + q of every n-tile: the shift of each pattern
                if (fr == 0) {
#pragma unroll
                    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                        for (int q = 0; q < 2; ++q) {
                            const int hi = j == 0 ? mx[0][nt][q] : mx[1][nt][q];
                            s_shift[warp * WR + nt * 8 + 2 * fc + q] = (hi < kScaleThresholdHi && hi >= 0x00100000) ? 1023 - (hi >> 20) : 0;
                        }
                }
                __syncwarp();
                // one pattern per lane: exponent of the output, and - rarely - the rescaling of a block row that is
                // already in global memory, all 32 lanes on one pattern's K A doubles at a time (coalesced)
                const int my_shift = lane_ok ? s_shift[warp * WR + lane] : 0;
                const int dst_blk = j == 0 ? row.dst[0] : row.dst[1];
                if (lane_ok)
                    p.scale[(size_t)dst_blk * S + wsite0 + lane] =
                        s_exp[warp * WR + lane] + s_exp[(j + 1) * TS + warp * WR + lane] - my_shift;
                unsigned todo = __ballot_sync(0xffffffffu, my_shift != 0);
                if (K * A <= 96) {
                    // up to four patterns at a time, all of their loads (three per lane and pattern) before the first
                    // store: the latency of the read-modify-write is paid once per four patterns
                    while (todo) {
                        double* mine[4];
                        double f[4], v[4][3];
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const int pi = todo ? __ffs(todo) - 1 : -1;
                            if (todo) todo &= todo - 1;
                            mine[u] = pi >= 0 ? p.clv + ((size_t)dst_blk * S + wsite0 + pi) * K * A : nullptr;
                            f[u] = pi >= 0 ? pow2i(__shfl_sync(0xffffffffu, my_shift, pi)) : 1.0;
                        }
#pragma unroll
                        for (int u = 0; u < 4; ++u)
#pragma unroll
                            for (int i = 0; i < 3; ++i)
                                v[u][i] = (mine[u] != nullptr && lane + 32 * i < K * A) ? mine[u][lane + 32 * i] : 0.0;
#pragma unroll
                        for (int u = 0; u < 4; ++u)
#pragma unroll
                            for (int i = 0; i < 3; ++i)
                                if (mine[u] != nullptr && lane + 32 * i < K * A) mine[u][lane + 32 * i] = v[u][i] * f[u];
                    }
                } else {
                    while (todo) {
                        const int pi = __ffs(todo) - 1;
                        todo &= todo - 1;
                        const double f = pow2i(__shfl_sync(0xffffffffu, my_shift, pi));
                        double* mine = p.clv + ((size_t)dst_blk * S + wsite0 + pi) * K * A;
                        for (int z0 = lane; z0 < K * A; z0 += 128) {   // four loads in flight per lane, then the stores
                            double v[4];
#pragma unroll
                            for (int u = 0; u < 4; ++u) v[u] = z0 + 32 * u < K * A ? mine[z0 + 32 * u] : 0.0;
#pragma unroll
                            for (int u = 0; u < 4; ++u)
                                if (z0 + 32 * u < K * A) mine[z0 + 32 * u] = v[u] * f;
                        }
                    }
                }
                __syncwarp();
            }
            __threadfence_block();   // this row's block is visible to the cp.async reads of the next row
            __syncwarp();
        }
    }
}

template <int AA, int MT, int KS, int NT, int WARPS, bool LEVEL>
int launch_mma(Ctx* c, const OpRow* d_rows, int row_begin, int row_end, const MmaRow* d_frows = nullptr) {
    constexpr int MROWS = MT * 8, KCOLS = KS * 4;
    constexpr int LDP = pad_pitch(KCOLS);
    constexpr int LDL = pad_pitch(KCOLS);
    constexpr int TS = WARPS * NT * 8;
    MmaArgs a;
    a.rows = d_rows;
    a.frows = d_frows;
    a.row_begin = row_begin;
    a.row_end = row_end;
    a.pmats = c->d_pmats;
    // tables are staged in the P buffer (MROWS rows): usable when every code has a row there
    a.nc = tip_table_rows(c);
    a.tiptab = (tip_tables_usable(c) && a.nc <= MROWS && !tuning().disable_tiptab) ? c->d_tiptab : nullptr;
    a.pimg = (AA % 2 == 1 && c->pimg_rows == MROWS && c->pimg_pitch == LDP) ? c->d_pimg : nullptr;
    a.codes = c->d_codes;
    a.pitch = c->code_pitch;
    a.lut = c->d_lut;
    a.clv = c->d_clv;
    a.scale = c->d_scale;
    a.S = c->S;
    a.n_tiles = (c->S + TS - 1) / TS;
    a.A = c->A;
    a.K = c->K;
    const size_t smem = (2 * (size_t)MROWS * LDP + 2 * (size_t)TS * LDL) * sizeof(double) + 4 * TS + 4 * TS * sizeof(int);
    auto kern = d_frows != nullptr ? (c->K == 4 ? mma_prune_kernel<AA, MT, KS, NT, WARPS, LEVEL, true, 4> : mma_prune_kernel<AA, MT, KS, NT, WARPS, LEVEL, true>)
                                   : (c->K == 4 ? mma_prune_kernel<AA, MT, KS, NT, WARPS, LEVEL, false, 4> : mma_prune_kernel<AA, MT, KS, NT, WARPS, LEVEL, false>);
    PHB_CUDA(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    PHB_CUDA(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, WARPS * 32, smem));
    if (per_sm < 1) per_sm = 1;
    const int64_t items = LEVEL ? (int64_t)(row_end - row_begin) * a.n_tiles : a.n_tiles;
    const int64_t cap = (int64_t)c->sm_count * per_sm;
    const int grid = (int)(items < cap ? items : cap);
    if (grid <= 0) return PHB_OK;
    kern<<<grid, WARPS * 32, smem, c->stream>>>(a);
    c->launches++;
    PHB_CUDA(c, cudaGetLastError());
    return PHB_OK;
}

template <int AA, int MT, int KS, int NT, int WARPS>
int run_rows_mma(Ctx* c, const RowSet& rs, int mode, const MmaRow* d_frows = nullptr) {
    if (rs.n_rows == 0) return PHB_OK;
    if (mode == PHB_MODE_LEVEL) {
        const std::vector<int32_t>& lv = *rs.levels;
        for (int l = 0; l + 1 < (int)lv.size(); ++l) {
            if (lv[l + 1] <= lv[l]) continue;
            int st = launch_mma<AA, MT, KS, NT, WARPS, true>(c, rs.d_rows, lv[l], lv[l + 1], d_frows);
            if (st) return st;
        }
        return PHB_OK;
    }
    return launch_mma<AA, MT, KS, NT, WARPS, false>(c, rs.d_rows, 0, rs.n_rows, d_frows);
}

}  // namespace

bool mma_supported(const Ctx* c) { return c->A == 20 || c->A == 61; }

// Padded staging images of the P blocks and tip tables [first_mat, first_mat + n_mats) (see MmaArgs::pimg)
__global__ void mma_image_kernel(const double* __restrict__ pmats, const double* __restrict__ tiptab, int A, int K, int nc,
                                 int rows, int pitch, int first_mat, double* __restrict__ out) {
    const int m = first_mat + blockIdx.x, k = blockIdx.y, which = blockIdx.z;
    const double* src = which == 0 ? pmats + ((size_t)m * K + k) * A * A : (tiptab ? tiptab + ((size_t)m * K + k) * nc * A : nullptr);
    const int n_src_rows = which == 0 ? A : (tiptab ? nc : 0);
    double* dst = out + (((size_t)m * K + k) * 2 + which) * rows * pitch;
    for (int e = threadIdx.x; e < rows * pitch; e += blockDim.x) {
        const int i = e / pitch, j = e - i * pitch;
        dst[e] = (i < n_src_rows && j < A) ? src[(size_t)i * A + j] : 0.0;
    }
}

int launch_mma_images(Ctx* c, int first_mat, int n_mats) {
    if (c->d_pimg == nullptr || n_mats <= 0) return PHB_OK;
    const int nc = tip_table_rows(c);
    const bool tables = tip_tables_usable(c) && nc <= c->pimg_rows;
    dim3 grid(n_mats, c->K, 2);
    mma_image_kernel<<<grid, 256, 0, c->stream>>>(c->d_pmats, tables ? c->d_tiptab : nullptr, c->A, c->K, nc, c->pimg_rows,
                                                  c->pimg_pitch, first_mat, c->d_pimg);
    c->launches++;
    PHB_CUDA(c, cudaGetLastError());
    return PHB_OK;
}

static int mma_dispatch(Ctx* c, const RowSet& rs, int mode, const MmaRow* d_frows) {
    const int variant = tuning().mma_variant;   // tuning knob: alternative tile shapes
    if (c->A == 20) {
        if (variant == 1) return run_rows_mma<20, 3, 5, 4, 8>(c, rs, mode, d_frows);   // 256 patterns per CTA, 8 warps
        return run_rows_mma<20, 3, 5, 4, 4>(c, rs, mode, d_frows);                     // 128 patterns per CTA, 4 warps
    }
    if (c->A == 61) {
        if (variant == 1) return run_rows_mma<61, 8, 16, 1, 8>(c, rs, mode, d_frows);  // 64 patterns per CTA
        return run_rows_mma<61, 8, 16, 2, 8>(c, rs, mode, d_frows);                    // 128 patterns per CTA
    }
    return c->fail(PHB_ERR_UNSUPPORTED, "DMMA kernels cover 20 and 61 states");
}

int mma_run_rows(Ctx* c, const RowSet& rs, int mode) { return mma_dispatch(c, rs, mode, nullptr); }

// Pre-order pass as one three-operand row per parent.  parents[i] = {X src, X kind, X pidx, child0 .., child1 .., up block
// of child0, up block of child1} in execution order (levels: offsets of independent groups, or empty for one launch that
// walks the rows in order); the row table is staged in the context's up-row buffer.
int mma_run_parent_rows(Ctx* c, const std::vector<int32_t>& parents, const std::vector<int32_t>& levels) {
    const int n = (int)(parents.size() / 11);
    if (n == 0) return PHB_OK;
    if ((size_t)n * sizeof(MmaRow) > 2 * (size_t)c->max_rows() * sizeof(OpRow))
        return c->fail(PHB_ERR_STATE, "pre-order pass: row table overflow");
    std::vector<MmaRow> rows(n);
    for (int i = 0; i < n; ++i) {
        const int32_t* q = &parents[(size_t)i * 11];
        MmaRow r;
        // op0 = X, op1 = child 1 (so that dst[0] = up[child 0]), op2 = child 0 (dst[1] = up[child 1])
        r.src[0] = q[0]; r.kind[0] = q[1]; r.pidx[0] = q[2];
        r.src[1] = q[6]; r.kind[1] = q[7]; r.pidx[1] = q[8];
        r.src[2] = q[3]; r.kind[2] = q[4]; r.pidx[2] = q[5];
        r.dst[0] = q[9];
        r.dst[1] = q[10];
        r.n_ops = 3;
        rows[i] = r;
    }
    MmaRow* d_frows = reinterpret_cast<MmaRow*>(c->d_up_rows);
    PHB_CUDA(c, cudaMemcpyAsync(d_frows, rows.data(), rows.size() * sizeof(MmaRow), cudaMemcpyHostToDevice, c->stream));
    PHB_CUDA(c, cudaStreamSynchronize(c->stream));   // `rows` is a stack object
    const RowSet rs{nullptr, n, &levels};
    return mma_dispatch(c, rs, levels.empty() ? PHB_MODE_TILE : PHB_MODE_LEVEL, d_frows);
}

}  // namespace phb
