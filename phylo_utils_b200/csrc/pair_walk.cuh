// lnL-only operand-resident pruning for 4-state models, two patterns per lane.
//
// Same arithmetic as clv_dna.cu (reference `clv`, numba_likelihood_engine.py:10-46; root
// step = tree_model.py:178-217 with lnl_node, numba_likelihood_engine.py:82-87), same parking plan
// (resident_plan.cuh), different mapping:
//
//   * a WARP owns a tile of 64 patterns and walks ALL rows (+ the virtual-root pseudo-row) for it; every lane
//     carries TWO adjacent patterns, so every broadcast read of a P row, every descriptor decode, every
//     prefetch address and every branch is paid once per 64 pattern-node updates instead of once per 32;
//   * the result of a row stays in registers when the next row consumes it (2/3 of the internal operands);
//   * a result needed later is parked in a per-warp scratch stripe that lives in L2.  The stripe layout is
//     private to the warp, so it is chosen for the hardware: 16-byte chunk c of lane l sits at
//     (c * 32 + l) * 16 - the park is 4K coalesced 128-bit stores straight from registers (no staging tile,
//     no __syncwarp), the fetch is 4K coalesced cp.async one row ahead, the consumer's LDS.128 is conflict
//     free.  Every lane reads back exactly the chunks it wrote itself;
//   * everything else a row needs - the two operands' P block (or the P.lut tip table of a tip operand), the
//     tile's tip codes (64 bytes per tip operand, 32 when packed two per byte) and the 16-byte row descriptor
//     (two rows ahead) - arrives by cp.async one row ahead in per-warp double buffers, so nothing a row reads
//     has a load latency on its critical path and no register carries a pending load across the row loop
//     (a first version that passed descriptors and codes through registers spent half its time stalled on
//     them: profiles/r01k_pair_v1_1000x1M.txt);
//   * there is ONE operand tile: the parked operand of row r+1 is fetched during row r, or right after row
//     r's arithmetic when row r reads the tile itself (1 row in 9);
//   * the last pseudo-row does the root combine, pi-dot, Gamma mixture, log and the weighted tile sum;
//   * SYM (reversible models - every model the reference's TreeModel can drive): the kernel is bound by the shared-memory
//     data pipe, and a third of its wavefronts are the warp-wide broadcast reads of the operands' 4 x 4 P blocks
//     (profiles/r02a_pair_lnl.txt).  Detailed balance makes R = diag(pi) P symmetric, so
//         (P L)_i = (1 / pi_i) sum_j r_ij L_j
//     needs the 10 numbers of R's upper triangle instead of the 16 of P: 5 broadcast LDS.128 per category and
//     operand instead of 8, the same 16 FMAs, and one multiplication by the constant pi_i^-n (n = internal
//     operands of the row) folded into the product.  No basis change, no cancellation: every term stays >= 0.
//
// This header holds the templates (device code, the two kernels, their launch); they are instantiated in four translation
// units - clv_dna_pair_k4n8.cu, _k4n16.cu, _k12.cu, _k8.cu - that compile side by side (one unit took 4.5 minutes), and
// clv_dna_pair.cu keeps the store walk, the parking plan and the host logic.
#pragma once

#include <algorithm>
#include <cstdlib>

#include "pair_common.cuh"

namespace phb {

// the lnL-only walk reads the symmetric form of the P blocks (10 numbers instead of 16) when the model is reversible
inline bool pair_sym(const Ctx* c) { return c->reversible && c->d_rmats != nullptr && !tuning().pair_full_p; }

// one function per translation unit of instantiations: (K, look-up table rows) groups, patterns per lane as an argument
int launch_pair_k4n8(Ctx* c, int ppt, int mode, int n_steps, int n_slots, int64_t b, int64_t e, double* ps, int max_grid, int* grid_out, int chunk_shift);
int launch_pair_k4n16(Ctx* c, int ppt, int mode, int n_steps, int n_slots, int64_t b, int64_t e, double* ps, int max_grid, int* grid_out, int chunk_shift);
int launch_pair_k12(Ctx* c, int nc, int mode, int n_steps, int n_slots, int64_t b, int64_t e, double* ps, int max_grid, int* grid_out, int chunk_shift);
int launch_pair_k8(Ctx* c, int nc, int mode, int n_steps, int n_slots, int64_t b, int64_t e, double* ps, int max_grid, int* grid_out, int chunk_shift);

namespace {


// 16-byte row descriptor of this kernel
struct __align__(16) PairRow {
    uint32_t off_a;   // 16-byte units from PairArgs::opbase: operand a's tip table (tip) or P block (otherwise)
    uint32_t off_b;
    int32_t src_a;    // tip row of operand a (operand a is never a parked block: canonical order TIP <= PREV <= SLOT)
    uint32_t packed;  // src_b [0:24) (tip row | scratch slot) | kind_a [24:26) | kind_b [26:28) | dst slot [28:32), 15 = none
};

struct PairArgs {
    const PairRow* rows;
    int n_steps;                  // rows walked per tile, including the root pseudo-row
    const unsigned char* opbase;  // base the descriptors' operand offsets refer to
    const uint8_t* codes;
    size_t pitch;                 // bytes between tip rows of `codes`
    const uint8_t* codes_hi;      // CODES_SPLIT3: the plane of high bits and its row pitch
    size_t pitch_hi;
    unsigned char* scratch;
    int n_slots;
    const double* freqs;
    const double* catw;
    const double* weights;
    double* pattern_lnl;
    double* partial_sums;         // one per CTA
    int64_t S;
    int64_t tile_begin, tile_end; // 64-pattern tiles covered by this launch
    // host->device pipelining (dna_pair_from_host): the codes of tile t are valid once flags[t / tiles_per_chunk] ==
    // epoch - written by the copy engine right behind the chunk's bytes.  flags == nullptr: codes are resident.
    const int* flags;
    int epoch, chunk_shift;       // tiles per chunk = 1 << chunk_shift
    int* error;                   // set to 1 if a chunk never arrived (bounded wait)
    double ipi[4];                // SYM: 1 / pi_i
    int stagger_ns;               // multi-warp CTA form: worker w starts (w mod 32) * stagger_ns late (0 = together)
};


// prev[p][k] <- (Pa[k] . a[p][k]) * (Pb[k] . b[p][k]) for the lane's PPT patterns; pe <- cumulative exponents
// y[p][i] <- sum_j r_ij v[p][j] from the packed upper triangle [r00 r01 | r02 r03 | r11 r12 | r13 r22 | r23 r33] of a
// symmetric 4 x 4 block: five warp-wide broadcast reads, sixteen FMAs per pattern
template <int PPT>
__device__ __forceinline__ void sym_matvec(const unsigned char* blk, const double (&v)[PPT][4], double (&y)[PPT][4]) {
    const double2* q = reinterpret_cast<const double2*>(blk);
    const double2 q0 = q[0], q1 = q[1];
#pragma unroll
    for (int p = 0; p < PPT; ++p) {
        y[p][0] = fma(q1.y, v[p][3], fma(q1.x, v[p][2], fma(q0.y, v[p][1], q0.x * v[p][0])));
        y[p][1] = q0.y * v[p][0];
        y[p][2] = q1.x * v[p][0];
        y[p][3] = q1.y * v[p][0];
    }
    const double2 q2 = q[2], q3 = q[3];
#pragma unroll
    for (int p = 0; p < PPT; ++p) {
        y[p][1] = fma(q3.x, v[p][3], fma(q2.y, v[p][2], fma(q2.x, v[p][1], y[p][1])));
        y[p][2] = fma(q3.y, v[p][2], fma(q2.y, v[p][1], y[p][2]));
        y[p][3] = fma(q3.x, v[p][1], y[p][3]);
    }
    const double2 q4 = q[4];
#pragma unroll
    for (int p = 0; p < PPT; ++p) {
        y[p][2] = fma(q4.x, v[p][3], y[p][2]);
        y[p][3] = fma(q4.y, v[p][3], fma(q4.x, v[p][2], y[p][3]));
    }
}

// pair_product: the products, the operands' exponents summed into pe, mh <- high word of the largest entry per pattern;
// pair_rescale: the threshold test on mh and the rescaling.  The lnL-only walk runs the second step once behind its
// four row shapes (one copy of the rare path, and the shapes end where their last product is written).
template <int K, int NC, int PPT, int CM, int KA, int KB, int LAYOUT = LAYOUT_PRIVATE, bool SYM = false>
__device__ __forceinline__ void pair_product(const unsigned char* st, const unsigned char* opin, int lane,
                                             double (&prev)[PPT][K][4], int (&pe)[PPT], int (&mh)[PPT], const double* ipi = nullptr) {
    constexpr int PB = SYM ? 80 : 128;   // bytes of one category's P block (SYM: the upper triangle of diag(pi) P)
    using L = PairLayout<K, NC, PPT>;
    constexpr int ROWB = K * 32 + 16;   // LAYOUT_ARRAY: one pattern's row of the operand tile
    static_assert(KA != KIND_SLOT, "operand a is a tip or the previous row");
    int e[PPT];
#pragma unroll
    for (int p = 0; p < PPT; ++p) e[p] = (KA == KIND_PREV || KB == KIND_PREV) ? pe[p] : 0;
    if (KB == KIND_SLOT && LAYOUT == LAYOUT_ARRAY) {
#pragma unroll
        for (int p = 0; p < PPT; ++p) e[p] += *reinterpret_cast<const int*>(opin + (lane + 32 * p) * ROWB + K * 32);   // in the row's padding
    } else if (KB == KIND_SLOT) {
        const int* x = reinterpret_cast<const int*>(opin + L::BLOCK_BYTES + lane * L::EXP_STRIDE);
        if (PPT == 2) {
            const int2 v = *reinterpret_cast<const int2*>(x);
            e[0] += v.x;
            e[1] += v.y;
        } else {
            const int4 v = *reinterpret_cast<const int4*>(x);
            e[0] += v.x;
            e[1] += v.y;
            e[2] += v.z;
            if (PPT == 4) e[PPT - 1] += v.w;
        }
    }
    int ra[PPT], rb[PPT];
#pragma unroll
    for (int p = 0; p < PPT; ++p) ra[p] = rb[p] = 0;
    if (KA == KIND_TIP) table_rows<NC, PPT, CM, LAYOUT>(st + L::CODES_OFF, lane, ra);
    if (KB == KIND_TIP) table_rows<NC, PPT, CM, LAYOUT>(st + L::CODES_OFF + L::TILE, lane, rb);
#pragma unroll
    for (int p = 0; p < PPT; ++p) mh[p] = 0;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        double x[PPT][4];
        if (KA == KIND_TIP) {
            // a tip operand contributes the row `code` of its staged table T[k] = P[k] . lut - no arithmetic
#pragma unroll
            for (int p = 0; p < PPT; ++p) lds32(st + k * NC * 32 + ra[p], x[p]);
        } else if (SYM) {
            double a[PPT][4];
#pragma unroll
            for (int p = 0; p < PPT; ++p)
#pragma unroll
                for (int i = 0; i < 4; ++i) a[p][i] = prev[p][k][i];
            sym_matvec<PPT>(st + k * PB, a, x);
        } else {
            const double2* q = reinterpret_cast<const double2*>(st + k * 128);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const double2 r0 = q[2 * i], r1 = q[2 * i + 1];   // P row i: warp-wide broadcast
#pragma unroll
                for (int p = 0; p < PPT; ++p)
                    x[p][i] = fma(r1.y, prev[p][k][3], fma(r1.x, prev[p][k][2], fma(r0.y, prev[p][k][1], r0.x * prev[p][k][0])));
            }
        }
        // the second operand's contribution is folded into x as it is produced
        if (KB == KIND_TIP) {
#pragma unroll
            for (int p = 0; p < PPT; ++p) {
                double y[4];
                lds32(st + L::OPER_BYTES + k * NC * 32 + rb[p], y);
#pragma unroll
                for (int i = 0; i < 4; ++i) x[p][i] *= y[i];
            }
        } else {
            double b[PPT][4];
#pragma unroll
            for (int p = 0; p < PPT; ++p) {
                if (KB == KIND_PREV) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) b[p][i] = prev[p][k][i];
                } else if (LAYOUT == LAYOUT_ARRAY) {
                    lds32(opin + (lane + 32 * p) * ROWB + k * 32, b[p]);
                } else {
                    const unsigned char* src = opin + ((p * K + k) * 2) * 512 + lane * 16;
                    const double2 lo = *reinterpret_cast<const double2*>(src);
                    const double2 hi = *reinterpret_cast<const double2*>(src + 512);
                    b[p][0] = lo.x; b[p][1] = lo.y; b[p][2] = hi.x; b[p][3] = hi.y;
                }
            }
            if (SYM) {
                double y[PPT][4];
                sym_matvec<PPT>(st + L::OPER_BYTES + k * PB, b, y);
#pragma unroll
                for (int p = 0; p < PPT; ++p)
#pragma unroll
                    for (int i = 0; i < 4; ++i) x[p][i] *= y[p][i];
            } else {
                const double2* q = reinterpret_cast<const double2*>(st + L::OPER_BYTES + k * 128);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const double2 r0 = q[2 * i], r1 = q[2 * i + 1];
#pragma unroll
                    for (int p = 0; p < PPT; ++p)
                        x[p][i] *= fma(r1.y, b[p][3], fma(r1.x, b[p][2], fma(r0.y, b[p][1], r0.x * b[p][0])));
                }
            }
        }
        if (SYM && (KA != KIND_TIP || KB != KIND_TIP)) {
            // (P L)_i = (R L)_i / pi_i for every internal operand of the row: one or two factors 1 / pi_i
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const double f = (KA != KIND_TIP && KB != KIND_TIP) ? ipi[i] * ipi[i] : ipi[i];
#pragma unroll
                for (int p = 0; p < PPT; ++p) x[p][i] *= f;
            }
        }
#pragma unroll
        for (int p = 0; p < PPT; ++p)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                prev[p][k][i] = x[p][i];
                mh[p] = max(mh[p], __double2hiint(x[p][i]));   // partials are >= 0: the high word orders them
            }
    }
#pragma unroll
    for (int p = 0; p < PPT; ++p) pe[p] = e[p];
}

// 0 < max < 2^-128: multiply by the exact power of two that brings the maximum into [1, 2)
template <int K, int PPT>
__device__ __forceinline__ void pair_rescale(double (&prev)[PPT][K][4], int (&pe)[PPT], const int (&mh)[PPT]) {
    bool small[PPT], any = false;
#pragma unroll
    for (int p = 0; p < PPT; ++p) {
        small[p] = mh[p] < kScaleThresholdHi && mh[p] >= 0x00100000;
        any = any || small[p];
    }
    if (__any_sync(0xffffffffu, any)) {
#pragma unroll
        for (int p = 0; p < PPT; ++p) {
            if (small[p]) {
                const int shift = 1023 - (mh[p] >> 20);
                const double f = pow2i(shift);
#pragma unroll
                for (int k = 0; k < K; ++k)
#pragma unroll
                    for (int i = 0; i < 4; ++i) prev[p][k][i] *= f;
                pe[p] -= shift;
            }
        }
    }
}


template <int K, int NC, int PPT, int CM, int KA, int KB, int LAYOUT = LAYOUT_PRIVATE, bool SYM = false>
__device__ __forceinline__ void pair_update(const unsigned char* st, const unsigned char* opin, int lane,
                                            double (&prev)[PPT][K][4], int (&pe)[PPT], const double* ipi = nullptr) {
    int mh[PPT];
    pair_product<K, NC, PPT, CM, KA, KB, LAYOUT, SYM>(st, opin, lane, prev, pe, mh, ipi);
    pair_rescale<K, PPT>(prev, pe, mh);
}

// Block (politely, and not forever) until the copy engine has delivered the chunk that holds tile t.
// Deliberately NOT inlined: it runs once per tile, and as a call its registers stay out of the row loop's allocation.
__device__ __noinline__ void wait_for_chunk(const int* flags, int chunk_shift, int epoch, int* error, int t) {
    PHB_DCHECK((t >> chunk_shift) < kMaxFlagChunks);
    const int* f = flags + (t >> chunk_shift);
#pragma unroll 1
    for (int spin = 0; spin < (1 << 23); ++spin) {   // ~10 s of 1 us naps: the copies were never issued - report, do not hang
        int v;
        asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
        if (v == epoch) return;
        __nanosleep(1000);
    }
    *error = 1;
}

// The walk of ONE warp: `worker` of `n_workers` takes the tiles tile_begin + worker, + n_workers, ... below tile_end (tiles
// of 32 PPT patterns), parks in its own stripe `my_scratch`, works in its own `smem` (WARP_BYTES + 256) and leaves the
// weighted sum of its tiles' lnL in *sum_out.
template <int K, int NC, int PPT, int CM, bool PIPE, bool SYM>
__device__ __forceinline__ void pair_walk(const PairArgs& p, unsigned char* const smem, const int lane, const int worker,
                                          const int n_workers, const int tile_begin, const int tile_end,
                                          unsigned char* const my_scratch, double* const sum_out) {
    using L = PairLayout<K, NC, PPT>;
    PairRow* const s_desc = reinterpret_cast<PairRow*>(smem);
    unsigned char* const s_stage = smem + L::DESC_BYTES;
    unsigned char* const s_opin = s_stage + 2 * L::STAGE_BYTES;
    const int wstride = n_workers, n_steps = p.n_steps;
    double* const s_acc = reinterpret_cast<double*>(smem + L::WARP_BYTES);   // per-lane running sum of weight * lnL
    s_acc[lane] = 0.0;

    // parked block `slot` -> the operand tile (the lane's own chunks, in the layout it wrote them)
    auto fetch_slot = [&](int slot) {
        PHB_DCHECK(slot >= 0 && slot < p.n_slots);
        const unsigned char* src = my_scratch + (size_t)slot * L::SLOT_BYTES;
#pragma unroll
        for (int j = 0; j < L::CHUNKS; ++j) cp_async16(s_opin + j * 512 + lane * 16, src + j * 512 + lane * 16);
        if (PPT == 2) cp_async8(s_opin + L::BLOCK_BYTES + lane * 8, src + L::BLOCK_BYTES + lane * 8);
        else cp_async16(s_opin + L::BLOCK_BYTES + lane * 16, src + L::BLOCK_BYTES + lane * 16);
    };
    // read-only inputs of row `d` at tile t -> stage buffer q: per operand its P block or tip table, and its codes
    auto stage_row = [&](const PairRow d, int t, int q) {
        unsigned char* st = s_stage + q * L::STAGE_BYTES;
        const int kind_a = (d.packed >> 24) & 3, kind_b = (d.packed >> 26) & 3;
        const unsigned char* ga = p.opbase + (size_t)d.off_a * 16 + lane * 16;
        const unsigned char* gb = p.opbase + (size_t)d.off_b * 16 + lane * 16;
#pragma unroll
        for (int j = 0; j < L::ROUNDS; ++j)
            if (j < L::P_ROUNDS || kind_a == KIND_TIP) cp_async16(st + j * 512 + lane * 16, ga + j * 512);
#pragma unroll
        for (int j = 0; j < L::ROUNDS; ++j)
            if (j < L::P_ROUNDS || kind_b == KIND_TIP) cp_async16(st + L::OPER_BYTES + j * 512 + lane * 16, gb + j * 512);
        // codes of the tile: TILE bytes per tip operand (TILE / 2 as nibbles; TILE / 4 + TILE / 8 as split 3-bit codes);
        // lanes 0..7 serve operand a, 8..15 operand b
        constexpr int CL = (CM == CODES_BYTE ? L::TILE : (CM == CODES_NIBBLE ? L::TILE / 2 : L::TILE / 4)) / 16;
        const int which = lane >> 3, piece = lane & 7;
        const bool tip = which == 0 ? kind_a == KIND_TIP : kind_b == KIND_TIP;
        PHB_DCHECK(kind_a != KIND_SLOT && t >= tile_begin && t < tile_end);
        if (which < 2 && tip) {
            const int tip_row = which == 0 ? d.src_a : (int)(d.packed & 0xffffff);
            PHB_DCHECK(tip_row >= 0 && (piece >= CL || (size_t)t * (CL * 16) + piece * 16 + 16 <= p.pitch));   // inside the tip's code row
            if (piece < CL)
                cp_async16(st + L::CODES_OFF + which * L::TILE + piece * 16,
                           p.codes + (size_t)tip_row * p.pitch + (size_t)t * (CL * 16) + piece * 16);
            if (CM == CODES_SPLIT3 && piece == CL) {   // the plane of high bits: TILE / 8 bytes, right behind the low plane
                unsigned char* dst = st + L::CODES_OFF + which * L::TILE + L::TILE / 4;
                const unsigned char* src = p.codes_hi + (size_t)tip_row * p.pitch_hi + (size_t)t * (L::TILE / 8);
                if (PPT == 2) cp_async8(dst, src);
                else cp_async16(dst, src);
            }
        }
    };

    int tile = tile_begin + worker;
    if (tile < tile_end) {
        // prologue: descriptors of rows 0 and 1, then the inputs of row 0
        if (lane < 2) cp_async16(&s_desc[lane], &p.rows[lane < n_steps ? lane : 0]);
        cp_async_commit();
        cp_async_wait_all();
        __syncwarp();
        {
            const PairRow d0 = s_desc[0];
            if (PIPE) wait_for_chunk(p.flags, p.chunk_shift, p.epoch, p.error, tile);
            stage_row(d0, tile, 0);
            if (((d0.packed >> 26) & 3) == KIND_SLOT) fetch_slot(d0.packed & 0xffffff);   // never: row 0 has no parked operand
        }
        cp_async_commit();

        double prev[PPT][K][4];
        int pe[PPT];
#pragma unroll
        for (int q = 0; q < PPT; ++q) {
            pe[q] = 0;
#pragma unroll
            for (int k = 0; k < K; ++k)
#pragma unroll
                for (int i = 0; i < 4; ++i) prev[q][k][i] = 0.0;
        }

        int row = 0, q = 0;
        int row2 = n_steps > 2 ? 2 : 0;   // row index two steps ahead (descriptors do not depend on the tile)
        while (true) {
            int row_n = row + 1, tile_n = tile;
            if (row_n == n_steps) {
                row_n = 0;
                tile_n += wstride;
            }
            const bool has_next = tile_n < tile_end;
            cp_async_wait_all();   // everything issued one row ago has had a whole row to land
            __syncwarp();
            const uint32_t pk = s_desc[q & 3].packed;
            const int kinds = (pk >> 24) & 15, dst_slot = pk >> 28;   // kind_a | kind_b << 2
            const bool opin_busy = (kinds >> 2) == KIND_SLOT;         // this row still has to read the operand tile
            bool fetch_late = false;
            int slot_n = 0;
            if (lane == 0) cp_async16(&s_desc[(q + 2) & 3], &p.rows[row2]);
            if (has_next) {
                const PairRow dn = s_desc[(q + 1) & 3];
                if (PIPE && row_n == 0) wait_for_chunk(p.flags, p.chunk_shift, p.epoch, p.error, tile_n);
                stage_row(dn, tile_n, (q + 1) & 1);
                if (((dn.packed >> 26) & 3) == KIND_SLOT) {
                    slot_n = dn.packed & 0xffffff;
                    if (opin_busy) fetch_late = true;
                    else fetch_slot(slot_n);
                }
            }
            cp_async_commit();

            const unsigned char* st = s_stage + (q & 1) * L::STAGE_BYTES;
            // four row shapes (canonical operand order); a short if-chain instead of a jump table: no table load and
            // indirect branch on the row's critical path
            const int kind_a = kinds & 3, kind_b = kinds >> 2;
            int mh[PPT];
            if (kind_b == KIND_SLOT) {
                if (kind_a == KIND_PREV) pair_product<K, NC, PPT, CM, KIND_PREV, KIND_SLOT, LAYOUT_PRIVATE, SYM>(st, s_opin, lane, prev, pe, mh, p.ipi);
                else pair_product<K, NC, PPT, CM, KIND_TIP, KIND_SLOT, LAYOUT_PRIVATE, SYM>(st, s_opin, lane, prev, pe, mh, p.ipi);
            } else if (kind_b == KIND_PREV) {
                pair_product<K, NC, PPT, CM, KIND_TIP, KIND_PREV, LAYOUT_PRIVATE, SYM>(st, s_opin, lane, prev, pe, mh, p.ipi);
            } else {
                pair_product<K, NC, PPT, CM, KIND_TIP, KIND_TIP, LAYOUT_PRIVATE, SYM>(st, s_opin, lane, prev, pe, mh, p.ipi);
            }
            pair_rescale<K, PPT>(prev, pe, mh);
            if (fetch_late) {   // the operand tile is free now (a lane only ever touches its own chunks of it)
                fetch_slot(slot_n);
                cp_async_commit();
            }
            if (row != n_steps - 1) {
                if (dst_slot != 15) {
                    PHB_DCHECK(dst_slot < p.n_slots);
                    // park: coalesced 128-bit stores straight from registers into the warp's own stripe
                    unsigned char* dst = my_scratch + (size_t)dst_slot * L::SLOT_BYTES + lane * 16;
#pragma unroll
                    for (int h = 0; h < PPT; ++h)
#pragma unroll
                        for (int k = 0; k < K; ++k) {
                            *reinterpret_cast<double2*>(dst + ((h * K + k) * 2) * 512) = make_double2(prev[h][k][0], prev[h][k][1]);
                            *reinterpret_cast<double2*>(dst + ((h * K + k) * 2 + 1) * 512) = make_double2(prev[h][k][2], prev[h][k][3]);
                        }
                    int* ex = reinterpret_cast<int*>(my_scratch + (size_t)dst_slot * L::SLOT_BYTES + L::BLOCK_BYTES + lane * L::EXP_STRIDE);
                    if (PPT == 2) *reinterpret_cast<int2*>(ex) = make_int2(pe[0], pe[1]);
                    else *reinterpret_cast<int4*>(ex) = make_int4(pe[0], pe[1], pe[2], PPT == 4 ? pe[PPT - 1] : 0);
                }
            } else {
                // root pseudo-row: pi-dot, Gamma mixture, log, weighted sum (tree_model.py:200-217)
                const int64_t s0 = (int64_t)tile * L::TILE + PPT * lane;
                PHB_DCHECK((int64_t)tile * L::TILE < p.S);
                double lnl[PPT];
#pragma unroll
                for (int h = 0; h < PPT; ++h) {
                    double mix = 0.0;
#pragma unroll
                    for (int k = 0; k < K; ++k) {
                        double f = p.freqs[0] * prev[h][k][0];
                        f = fma(p.freqs[1], prev[h][k][1], f);
                        f = fma(p.freqs[2], prev[h][k][2], f);
                        f = fma(p.freqs[3], prev[h][k][3], f);
                        if (f > 0) mix = fma(p.catw[k], f, mix);
                    }
                    lnl[h] = mix > 0 ? log(mix) + (double)pe[h] * kLn2 : -INFINITY;
                }
                double acc = s_acc[lane];
                if (PPT % 2 == 0 && s0 + PPT <= p.S) {
#pragma unroll
                    for (int h = 0; h + 1 < PPT; h += 2) {
                        *reinterpret_cast<double2*>(p.pattern_lnl + s0 + h) = make_double2(lnl[h], lnl[h + 1]);
                        if (p.weights) {
                            const double2 w = *reinterpret_cast<const double2*>(p.weights + s0 + h);
                            acc += w.x * lnl[h];
                            acc += w.y * lnl[h + 1];
                        } else {
                            acc += lnl[h];
                            acc += lnl[h + 1];
                        }
                    }
                } else {
#pragma unroll
                    for (int h = 0; h < PPT; ++h)
                        if (s0 + h < p.S) {
                            p.pattern_lnl[s0 + h] = lnl[h];
                            acc += (p.weights ? p.weights[s0 + h] : 1.0) * lnl[h];
                        }
                }
                s_acc[lane] = acc;
            }
            if (!has_next) break;
            row = row_n;
            tile = tile_n;
            if (++row2 == n_steps) row2 = 0;
            ++q;
        }
        cp_async_wait_all();
    }
    const double total = warp_sum(s_acc[lane]);
    if (lane == 0) *sum_out = total;
}

// one warp per CTA: the CTAs of an SM spread over its four sub-partitions
template <int K, int NC, int PPT, int CM, bool PIPE, bool SYM>
__global__ void __launch_bounds__(32, PairLayout<K, NC, PPT>::MIN_CTAS) dna_pair_kernel(const __grid_constant__ PairArgs p) {
    extern __shared__ __align__(128) unsigned char smem[];
    pair_walk<K, NC, PPT, CM, PIPE, SYM>(p, smem, threadIdx.x, blockIdx.x, gridDim.x, (int)p.tile_begin, (int)p.tile_end,
                                         p.scratch + (size_t)blockIdx.x * p.n_slots * PairLayout<K, NC, PPT>::SLOT_BYTES,
                                         p.partial_sums + blockIdx.x);
}

// All the warps an SM can hold in ONE CTA.  Warp w of a CTA runs on sub-partition w mod 4, so every sub-partition gets
// the same number of walks (twelve 1-warp CTAs leave that to the block scheduler), and the warps of a CTA start together
// and stay close to each other in the row loop: they fetch the same instructions and the same P blocks at about the
// same time (1M patterns: 13.9 ms against 14.5 ms for the 1-warp CTAs on the same box, profiles/r02g_*).  Worker ids
// are interleaved over the CTAs - worker = warp * CTAs + CTA - so that a last, partial round of tiles spreads over all
// SMs instead of filling the first few.
template <int K, int NC, int PPT, int CM, bool PIPE, bool SYM>
__global__ void __launch_bounds__(32 * PairLayout<K, NC, PPT>::MIN_CTAS, 1) dna_pair_cta_kernel(const __grid_constant__ PairArgs p) {
    using L = PairLayout<K, NC, PPT>;
    extern __shared__ __align__(128) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int worker = warp * gridDim.x + blockIdx.x, n_workers = (int)(blockDim.x >> 5) * gridDim.x;
    if (p.stagger_ns > 0) __nanosleep((unsigned)((worker * 7) & 31) * (unsigned)p.stagger_ns);
    pair_walk<K, NC, PPT, CM, PIPE, SYM>(p, smem + (size_t)warp * (L::WARP_BYTES + 256), lane, worker, n_workers, (int)p.tile_begin,
                                         (int)p.tile_end, p.scratch + (size_t)worker * p.n_slots * L::SLOT_BYTES, p.partial_sums + worker);
}

template <int CM, bool SYM>
PairArgs pair_args(Ctx* c, int n_steps, int n_slots, int64_t tile_begin, int64_t tile_end, double* partial_sums, int chunk_shift) {
    PairArgs a;
    a.rows = static_cast<const PairRow*>(c->d_res_rows);
    a.n_steps = n_steps;
    a.opbase = reinterpret_cast<const unsigned char*>(c->d_pmats);
    a.codes = c->d_codes;
    a.pitch = CM == CODES_BYTE ? c->code_pitch : (CM == CODES_NIBBLE ? c->code_pitch / 2 : c->code_pitch / 4);
    a.codes_hi = c->d_codes + (size_t)c->n_tips * (c->code_pitch / 4);
    a.pitch_hi = c->code_pitch / 8;
    a.scratch = c->d_scratch;
    a.n_slots = n_slots;
    a.freqs = c->model_freqs();
    a.catw = c->model_catw();
    a.weights = c->d_weights;
    a.pattern_lnl = c->d_pattern_lnl;
    a.partial_sums = partial_sums;
    a.S = c->S;
    a.tile_begin = tile_begin;
    a.tile_end = tile_end;
    int* const flags = c->d_flags_cur != nullptr ? c->d_flags_cur : c->d_flags;
    a.flags = flags;
    a.epoch = c->flag_epoch;
    a.chunk_shift = chunk_shift;
    a.error = flags + kMaxFlagChunks;
    for (int i = 0; i < 4; ++i) a.ipi[i] = SYM ? 1.0 / c->h_freqs[i] : 1.0;
    a.stagger_ns = tuning().pair_stagger;
    return a;
}

template <int K, int NC, int PPT, int CM, bool PIPE, bool SYM>
int launch_pair(Ctx* c, int n_steps, int n_slots, int64_t tile_begin, int64_t tile_end, double* partial_sums,
                int max_grid, int* grid_out, int chunk_shift) {
    using L = PairLayout<K, NC, PPT>;
    PairArgs a = pair_args<CM, SYM>(c, n_steps, n_slots, tile_begin, tile_end, partial_sums, chunk_shift);
    // Which form?  Measured on one box (profiles/r02g_cta_form_ab.jsonl, 1-warp CTAs -> one CTA per SM): 8.8 rounds of
    // tiles 14.56 -> 13.78 ms, 4.4 rounds 7.20 -> 7.25, 2.2 rounds 3.68 -> 3.94, one wave 1.96 -> 2.21: the big CTAs win
    // where every warp walks many tiles, the independent 1-warp CTAs (launched one after the other, out of step from
    // the start) where it walks one or two.
    const int64_t rounds_x10 = 10 * (tile_end - tile_begin) / ((int64_t)c->sm_count * L::MIN_CTAS);
    const int min_rounds_x10 = tuning().pair_cta_rounds != 0 ? 10 * tuning().pair_cta_rounds : 60;
    if (!tuning().pair_one_warp_ctas && rounds_x10 >= min_rounds_x10) {
        // one CTA per SM that carries all the warps the tiles need (whole rounds of the four sub-partitions)
        auto kern = dna_pair_cta_kernel<K, NC, PPT, CM, PIPE, SYM>;
        const size_t per_warp = L::WARP_BYTES + 256;   // + the per-lane running sums
        const int64_t n_tiles = tile_end - tile_begin;
        int warps = (int)std::min<int64_t>(L::MIN_CTAS, (n_tiles + c->sm_count - 1) / c->sm_count);
        warps = warps <= 1 ? 1 : std::min<int>(L::MIN_CTAS, (warps + 3) / 4 * 4);
        if (tuning().pair_ctas > 0) warps = std::min(warps, tuning().pair_ctas);
        warps = (int)std::min<size_t>((size_t)warps, c->smem_optin / per_warp);
        if (warps < 1) return c->fail(PHB_ERR_UNSUPPORTED, "pair kernel: does not fit in shared memory");
        int64_t ctas = std::min<int64_t>(c->sm_count, (n_tiles + warps - 1) / warps);
        // every warp needs its own scratch stripe and its own partial sum
        const int64_t cap = std::min<int64_t>((int64_t)(c->scratch_bytes / ((size_t)n_slots * L::SLOT_BYTES)), max_grid);
        if (cap < 1) return c->fail(PHB_ERR_NOMEM, "pair kernel: scratch area too small");
        if (ctas * warps > cap) {
            warps = (int)std::max<int64_t>(1, std::min<int64_t>(warps, cap / std::max<int64_t>(1, ctas)));
            ctas = std::max<int64_t>(1, std::min<int64_t>(ctas, cap / warps));
        }
        const size_t smem = (size_t)warps * per_warp;
        PHB_CUDA(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        PHB_CUDA(c, cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        kern<<<(int)ctas, 32 * warps, smem, c->stream>>>(a);
        c->launches++;
        PHB_CUDA(c, cudaGetLastError());
        c->resident_warps = warps;
        *grid_out = (int)(ctas * warps);
        return PHB_OK;
    }
    auto kern = dna_pair_kernel<K, NC, PPT, CM, PIPE, SYM>;
    const size_t smem = L::WARP_BYTES + 256;   // + the per-lane running sums
    if (smem > c->smem_optin) return c->fail(PHB_ERR_UNSUPPORTED, "pair kernel: does not fit in shared memory");
    PHB_CUDA(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    PHB_CUDA(c, cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    int per_sm = 0;
    PHB_CUDA(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 32, smem));
    if (per_sm < 1) per_sm = 1;
    if (tuning().pair_ctas > 0) per_sm = std::min(per_sm, tuning().pair_ctas);
    const int64_t n_tiles = tile_end - tile_begin, resident = (int64_t)c->sm_count * per_sm;
    int64_t grid = std::min<int64_t>(n_tiles, resident);
    if (tuning().pair_grid == 2 && n_tiles > resident) {
        // the same number of tiles for every warp: no tail round in which a few warps walk the tree on their own
        const int64_t rounds = (n_tiles + resident - 1) / resident;
        grid = (n_tiles + rounds - 1) / rounds;
    }
    grid = std::min<int64_t>(grid, max_grid);
    // every resident warp needs its own scratch stripe
    const int64_t cap = (int64_t)(c->scratch_bytes / ((size_t)n_slots * L::SLOT_BYTES));
    if (cap < 1) return c->fail(PHB_ERR_NOMEM, "pair kernel: scratch area too small");
    grid = std::max<int64_t>(1, std::min(grid, cap));
    kern<<<(int)grid, 32, smem, c->stream>>>(a);
    c->launches++;
    PHB_CUDA(c, cudaGetLastError());
    c->resident_warps = per_sm;
    *grid_out = (int)grid;
    return PHB_OK;
}

template <int K, int NC, int PPT>
int launch_pair_v(Ctx* c, int mode, int n_steps, int n_slots, int64_t b, int64_t e, double* ps, int max_grid,
                  int* grid_out, int chunk_shift) {
    const int cs = chunk_shift < 0 ? 0 : chunk_shift;
    const int flavour = mode * 4 + (chunk_shift >= 0 ? 2 : 0) + (pair_sym(c) ? 1 : 0);
    switch (flavour) {
#define PHB_PAIR_FLAVOUR(F_, CM_, PIPE_, SYM_) \
    case F_: return launch_pair<K, NC, PPT, CM_, PIPE_, SYM_>(c, n_steps, n_slots, b, e, ps, max_grid, grid_out, cs);
        PHB_PAIR_FLAVOUR(1, CODES_BYTE, false, true)
#ifdef PHB_PAIR_PROBE_ONLY   // developer builds: only the headline instantiation (SASS inspection in seconds)
    }
    return PHB_ERR_UNSUPPORTED;
#else
        PHB_PAIR_FLAVOUR(0, CODES_BYTE, false, false)
        PHB_PAIR_FLAVOUR(2, CODES_BYTE, true, false)
        PHB_PAIR_FLAVOUR(3, CODES_BYTE, true, true)
        PHB_PAIR_FLAVOUR(4, CODES_NIBBLE, false, false)
        PHB_PAIR_FLAVOUR(5, CODES_NIBBLE, false, true)
        PHB_PAIR_FLAVOUR(6, CODES_NIBBLE, true, false)
        PHB_PAIR_FLAVOUR(7, CODES_NIBBLE, true, true)
#undef PHB_PAIR_FLAVOUR
    }
    // split 3-bit codes: look-up tables of at most 8 rows, reversible models (the symmetric-block walk)
    if constexpr (NC == 8) {
        if (flavour == 9) return launch_pair<K, NC, PPT, CODES_SPLIT3, false, true>(c, n_steps, n_slots, b, e, ps, max_grid, grid_out, cs);
        if (flavour == 11) return launch_pair<K, NC, PPT, CODES_SPLIT3, true, true>(c, n_steps, n_slots, b, e, ps, max_grid, grid_out, cs);
    }
    return c->fail(PHB_ERR_UNSUPPORTED, "split 3-bit tip codes need a look-up table of at most 8 rows and a reversible model");
#endif
}

}  // namespace
}  // namespace phb
