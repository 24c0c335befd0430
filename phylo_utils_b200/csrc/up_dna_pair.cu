// Pre-order ("up") pass for 4-state models as ONE operand-resident walk.
//
// up[c] = partial of everything OUTSIDE the subtree of c, seen at the top end of c's branch.  For a parent `par` with
// children o and k and X = what sits above par (up[par], or the partial of the other end of the root edge):
//
//     U     = P(par's branch) . X
//     up[o] = U * (P_k . down[k])          up[k] = U * (P_o . down[o])
//
// (the composition of the reference's `clv`, numba_likelihood_engine.py:10-46, along the re-rooting rows
// [PAR, SIB, GPA] of utils.py:137-188).  The streaming tile walk runs this as two independent
// rows per parent - four matrix-vector products, X read twice.  Here, like the post-order walks of clv_dna_pair.cu:
//
//   * a WARP owns a tile of 32 * PPT patterns and walks the whole tree for it in pre-order, the smaller internal
//     child first; a lane carries PPT patterns with all K categories;
//   * one step per parent: U once, both children's outputs from it (three products instead of four);
//   * up[k] of the child the walk descends into next stays in REGISTERS and is the next step's X; the other child's
//     block is picked up again from the array it has to be written to anyway (a LOAD pseudo-step, at most one per
//     parent with two internal children; smaller-child-first keeps that re-read close behind the write: L2);
//   * the children's down blocks arrive by cp.async one step ahead into double-buffered pattern-major tiles (rows
//     padded by 16 bytes - conflict-free 128-bit reads, and the pad holds the row's exponent); tip children contribute
//     rows of the staged P.lut tables, no arithmetic;
//   * both outputs leave as coalesced streaming 128-bit stores, staged through the (by then dead) operand tile of the
//     child they belong to.
//
// HBM traffic per parent: two block writes, one block read per internal child, one re-read per two-internal-children
// parent - against two writes and up to four reads for the two-row form.
//
// Sum-table form (ST, the default).  What the edge-derivative kernels need from up[c] is only, per category k,
//     s_km = (V^-1 down[c]_k)_m * (V^T (pi * up[c]_k))_m          f_k^(d)(t) = sum_m g_d(lambda_m r_k) e^(lambda_m r_k t) s_km
// (derivs.cu).  At the parent's step both up[c] (registers) and down[c] (the operand tile) are on chip for BOTH
// children, so the walk writes s - one block per edge, the size of a partial - into the up block instead of up[c], and
// every Newton iteration reads ONE block per edge instead of two.  An up[other] that is needed again is parked in the
// warp's private scratch stripe (clv_dna_pair.cu's layout: coalesced 128-bit stores straight from registers; at most
// log2(internal nodes) deep because the lighter child goes first).
#include <algorithm>
#include <cstdlib>

#include "pair_common.cuh"

namespace phb {

namespace {

constexpr int UP_X_PREV = 0, UP_X_TIP = 1, UP_LOAD = 2, UP_LOAD_SLOT = 3;

// 32-byte step descriptor
struct __align__(16) UpStep {
    uint32_t off_x;   // 16-byte units from UpArgs::opbase: P block of the branch above par, or its tip table (UP_X_TIP)
    uint32_t off_o;   // the other child's P block / tip table
    uint32_t off_k;   // the kept child's
    int32_t src_x;    // tip row of X (UP_X_TIP) | block (UP_LOAD) or scratch slot (UP_LOAD_SLOT) to load into the registers
    int32_t src_o;    // tip row | down block of the other child
    int32_t src_k;
    int32_t dst_o;    // block that receives up[other]
    uint32_t packed;  // block of up[kept] [0:24) | mode [24:26) | other child is a tip [26] | kept child is a tip [27]
                      // | scratch slot up[other] is parked in [28:32), 15 = none (ST only)
};
static_assert(sizeof(UpStep) == 32, "UpStep must stay 32 bytes");

struct UpArgs {
    const UpStep* steps;
    int n_steps;
    const unsigned char* opbase;
    const uint8_t* codes;
    size_t pitch;
    double* clv;       // shared block array: down blocks, then up blocks  [block][S][K][4]
    int32_t* scale;    // [block][S]
    int64_t S, n_tiles;
    // sum-table form
    double m1[16];     // V^-1, row-major
    double m2[16];     // m2[m][i] = V[i][m] pi[i]
    const double* lut; // [256][4]
    unsigned char* scratch;
    int n_slots;
};

template <int K, int NC, int PPT>
struct UpLayout {
    using L = PairLayout<K, NC, PPT>;
    static constexpr int ROWB = K * 32 + 16;                            // one pattern's row of a tile (+ exponent)
    static constexpr int TILE_BYTES = L::TILE * ROWB;
    static constexpr int STAGE_BYTES = 3 * L::OPER_BYTES + 3 * L::TILE; // x, o, k operand blocks + their codes
    static constexpr int CODES_OFF = 3 * L::OPER_BYTES;
    static constexpr int DESC_BYTES = 4 * 32;
    static constexpr int XTAB_BYTES = NC * 32;                          // ST: V^-1 . lut[code]
    // (no staging tile of its own: an output leaves through the operand tile of the child it belongs to, which is
    // dead by then)
    static constexpr int WARP_BYTES = DESC_BYTES + 2 * STAGE_BYTES + 4 * TILE_BYTES + XTAB_BYTES;
    // a parked block in the warp's scratch stripe (the private chunk layout of clv_dna_pair.cu)
    static constexpr int CHUNKS = PPT * K * 2;
    static constexpr int BLOCK_BYTES = CHUNKS * 512;
    static constexpr int SLOT_BYTES = BLOCK_BYTES + 128 * PPT;
    static_assert(SLOT_BYTES <= TILE_BYTES, "a parked block is fetched into an operand tile");
};

// y[p] <- P[k] . v[p]: the four rows of P[k] are warp-wide broadcast reads
template <int PPT>
__device__ __forceinline__ void matvec(const unsigned char* pk, const double (&v)[PPT][4], double (&y)[PPT][4]) {
    const double2* q = reinterpret_cast<const double2*>(pk);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const double2 r0 = q[2 * i], r1 = q[2 * i + 1];
#pragma unroll
        for (int p = 0; p < PPT; ++p) y[p][i] = fma(r1.y, v[p][3], fma(r1.x, v[p][2], fma(r0.y, v[p][1], r0.x * v[p][0])));
    }
}

// 0 < max < 2^-128: multiply by the exact power of two that brings the maximum into [1, 2)
template <int K, int PPT>
__device__ __forceinline__ void rescale(double (&v)[PPT][K][4], const int (&mh)[PPT], int (&e)[PPT]) {
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
                    for (int i = 0; i < 4; ++i) v[p][k][i] *= f;
                e[p] -= shift;
            }
        }
    }
}

// d[p] <- what a child contributes in category k: a row of its staged P.lut table (tip), or P[k] . (its down partial).
// `tip` is warp-uniform: a branch instead of a template parameter keeps the step loop's code small (the eight-way
// instantiation cost 10 % of the issue slots in instruction-cache misses: profiles/r01t_up_walk_cfg5.txt)
template <int K, int NC, int PPT>
__device__ __forceinline__ void child_term(bool tip, const unsigned char* oper, const unsigned char* tile,
                                           const int (&trow)[PPT], int lane, int k, double (&d)[PPT][4]) {
    constexpr int ROWB = K * 32 + 16;
    if (tip) {
#pragma unroll
        for (int p = 0; p < PPT; ++p) lds32(oper + k * NC * 32 + trow[p], d[p]);
    } else {
        double v[PPT][4];
#pragma unroll
        for (int p = 0; p < PPT; ++p) lds32(tile + (lane + 32 * p) * ROWB + k * 32, v[p]);
        matvec<PPT>(oper + k * 128, v, d);
    }
}

// prev <- up[kept], oth <- up[other]; on entry prev / pe hold X and its exponents (unless X is a tip)
template <int K, int NC, int PPT>
__device__ __forceinline__ void up_update(bool x_tip, bool o_tip, bool k_tip, const unsigned char* st,
                                          const unsigned char* tile_o, const unsigned char* tile_k, int lane,
                                          double (&prev)[PPT][K][4], int (&pe)[PPT], double (&oth)[PPT][K][4],
                                          int (&oe)[PPT], int (&ed_o)[PPT], int (&ed_k)[PPT]) {
    using U = UpLayout<K, NC, PPT>;
    using L = PairLayout<K, NC, PPT>;
    constexpr int ROWB = U::ROWB;
    int rx[PPT], ro[PPT], rk[PPT];
#pragma unroll
    for (int p = 0; p < PPT; ++p) rx[p] = ro[p] = rk[p] = 0;
    if (x_tip) table_rows<NC, PPT, false, LAYOUT_ARRAY>(st + U::CODES_OFF, lane, rx);
    if (o_tip) table_rows<NC, PPT, false, LAYOUT_ARRAY>(st + U::CODES_OFF + L::TILE, lane, ro);
    if (k_tip) table_rows<NC, PPT, false, LAYOUT_ARRAY>(st + U::CODES_OFF + 2 * L::TILE, lane, rk);
    int mo[PPT], mk[PPT];
#pragma unroll
    for (int p = 0; p < PPT; ++p) {
        const int ex = x_tip ? 0 : pe[p];
        ed_k[p] = k_tip ? 0 : *reinterpret_cast<const int*>(tile_k + (lane + 32 * p) * ROWB + K * 32);
        ed_o[p] = o_tip ? 0 : *reinterpret_cast<const int*>(tile_o + (lane + 32 * p) * ROWB + K * 32);
        oe[p] = ex + ed_k[p];
        pe[p] = ex + ed_o[p];
        mo[p] = mk[p] = 0;
    }
#pragma unroll
    for (int k = 0; k < K; ++k) {
        double u[PPT][4], d[PPT][4];
        if (x_tip) {
#pragma unroll
            for (int p = 0; p < PPT; ++p) lds32(st + k * NC * 32 + rx[p], u[p]);
        } else {
            double x[PPT][4];
#pragma unroll
            for (int p = 0; p < PPT; ++p)
#pragma unroll
                for (int i = 0; i < 4; ++i) x[p][i] = prev[p][k][i];
            matvec<PPT>(st + k * 128, x, u);
        }
        child_term<K, NC, PPT>(k_tip, st + 2 * L::OPER_BYTES, tile_k, rk, lane, k, d);
#pragma unroll
        for (int p = 0; p < PPT; ++p)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const double r = u[p][i] * d[p][i];
                oth[p][k][i] = r;
                mo[p] = max(mo[p], __double2hiint(r));   // partials are >= 0: the high word orders them
            }
        child_term<K, NC, PPT>(o_tip, st + L::OPER_BYTES, tile_o, ro, lane, k, d);
#pragma unroll
        for (int p = 0; p < PPT; ++p)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const double r = u[p][i] * d[p][i];
                prev[p][k][i] = r;
                mk[p] = max(mk[p], __double2hiint(r));
            }
    }
    rescale<K, PPT>(oth, mo, oe);
    rescale<K, PPT>(prev, mk, pe);
}

template <int K, int NC, int PPT, bool ST>
__global__ void __launch_bounds__(32, PPT == 1 ? 8 : 4) dna_up_kernel(const __grid_constant__ UpArgs p) {
    using U = UpLayout<K, NC, PPT>;
    using L = PairLayout<K, NC, PPT>;
    constexpr int ROWB = U::ROWB;
    constexpr int PIECES = K * 2;                        // 16-byte pieces per pattern
    constexpr int ROUNDS = L::TILE * PIECES / 32;        // warp-wide copy rounds per block tile
    extern __shared__ __align__(128) unsigned char smem[];
    const int lane = threadIdx.x;
    UpStep* const s_desc = reinterpret_cast<UpStep*>(smem);
    unsigned char* const s_stage = smem + U::DESC_BYTES;
    unsigned char* const s_tiles = s_stage + 2 * U::STAGE_BYTES;   // [buffer][o | k]
    unsigned char* const s_xtab = s_tiles + 4 * U::TILE_BYTES;
    const int wstride = gridDim.x, n_steps = p.n_steps;
    const int n_tiles = (int)p.n_tiles;
    const size_t S = (size_t)p.S;
    unsigned char* const my_scratch = ST ? p.scratch + (size_t)blockIdx.x * p.n_slots * U::SLOT_BYTES : nullptr;
    if (ST) {   // what a tip contributes to the first factor of a sum table: V^-1 . lut[code]
        if (lane < NC) {
            double v[4], x[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) v[i] = p.lut[lane * 4 + i];
#pragma unroll
            for (int m = 0; m < 4; ++m) x[m] = fma(p.m1[4 * m + 3], v[3], fma(p.m1[4 * m + 2], v[2], fma(p.m1[4 * m + 1], v[1], p.m1[4 * m] * v[0])));
            *reinterpret_cast<double2*>(s_xtab + lane * 32) = make_double2(x[0], x[1]);
            *reinterpret_cast<double2*>(s_xtab + lane * 32 + 16) = make_double2(x[2], x[3]);
        }
        __syncwarp();
    }

    // A warp-wide copy round moves 32 pieces of 16 bytes = PPR patterns; the lane's piece of round j of a block tile sits
    // at lane_off + j * PPR * ROWB in the padded tile and at lane * 16 + j * 512 in the block.  Every tile but the
    // alignment's last one is whole: its copies carry no per-piece bounds and no per-piece address arithmetic (the
    // per-piece form was 90 - 115 instructions per block moved, a quarter of the walk's instructions:
    // profiles/r02i_up_walk_sass_profile.txt); the ragged last tile goes through compact loops.
    constexpr int PPR = 32 / PIECES;
    static_assert(32 % PIECES == 0, "a copy round covers whole patterns");
    const int lane_off = (lane / PIECES) * ROWB + (lane % PIECES) * 16;
    const size_t block_bytes = S * (size_t)(K * 32);
    unsigned char* const clv_bytes = reinterpret_cast<unsigned char*>(p.clv);

    // block `blk` of tile t -> a pattern-major operand tile (exponents into the row padding)
    auto fetch_block = [&](int blk, int t, unsigned char* dst) {
        PHB_DCHECK(blk >= 0 && t >= 0 && t < n_tiles);
        const size_t site0 = (size_t)t * L::TILE;
        const unsigned char* src = clv_bytes + (size_t)blk * block_bytes + site0 * (K * 32);
        const int32_t* ex = p.scale + (size_t)blk * S + site0;
        const unsigned d0 = (unsigned)__cvta_generic_to_shared(dst);
        if (site0 + L::TILE <= S) {
            const unsigned d = d0 + lane_off;
            const unsigned char* g = src + lane * 16;
#pragma unroll
            for (int j = 0; j < ROUNDS; ++j)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d + j * (PPR * ROWB)), "l"(g + j * 512) : "memory");
#pragma unroll
            for (int q = 0; q < PPT; ++q)
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d0 + (lane + 32 * q) * ROWB + K * 32), "l"(ex + lane + 32 * q) : "memory");
            return;
        }
        const int valid = (int)(S - site0);
#pragma unroll 1
        for (int c = lane; c < valid * PIECES; c += 32)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d0 + (c / PIECES) * ROWB + (c % PIECES) * 16), "l"(src + (size_t)c * 16) : "memory");
#pragma unroll 1
        for (int r = lane; r < valid; r += 32)
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d0 + r * ROWB + K * 32), "l"(ex + r) : "memory");
    };
    // everything step `d` reads at tile t -> buffer b: operand P blocks / tip tables, tip codes, down-block tiles
    auto stage_step = [&](const UpStep d, int t, int b) {
        unsigned char* st = s_stage + b * U::STAGE_BYTES;
        unsigned char* tl = s_tiles + b * 2 * U::TILE_BYTES;
        const int mode = (d.packed >> 24) & 3;
        if (mode == UP_LOAD) {
            fetch_block(d.src_x, t, tl);
            return;
        }
        if (mode == UP_LOAD_SLOT) {   // the lane's own chunks of a parked block, in the layout it wrote them
            PHB_DCHECK(ST && d.src_x >= 0 && d.src_x < p.n_slots);
            const unsigned char* src = my_scratch + (size_t)d.src_x * U::SLOT_BYTES;
#pragma unroll
            for (int j = 0; j < U::CHUNKS; ++j) cp_async16(tl + j * 512 + lane * 16, src + j * 512 + lane * 16);
            if (PPT == 2) cp_async8(tl + U::BLOCK_BYTES + lane * 8, src + U::BLOCK_BYTES + lane * 8);
            else asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(tl + U::BLOCK_BYTES + lane * 4)),
                              "l"(src + U::BLOCK_BYTES + lane * 4) : "memory");
            return;
        }
        const bool x_tip = mode == UP_X_TIP, o_tip = (d.packed >> 26) & 1, k_tip = (d.packed >> 27) & 1;
        const unsigned char* gx = p.opbase + (size_t)d.off_x * 16 + lane * 16;
        const unsigned char* go = p.opbase + (size_t)d.off_o * 16 + lane * 16;
        const unsigned char* gk = p.opbase + (size_t)d.off_k * 16 + lane * 16;
#pragma unroll
        for (int j = 0; j < L::ROUNDS; ++j) {
            if (j < L::P_ROUNDS || x_tip) cp_async16(st + j * 512 + lane * 16, gx + j * 512);
            if (j < L::P_ROUNDS || o_tip) cp_async16(st + L::OPER_BYTES + j * 512 + lane * 16, go + j * 512);
            if (j < L::P_ROUNDS || k_tip) cp_async16(st + 2 * L::OPER_BYTES + j * 512 + lane * 16, gk + j * 512);
        }
        // tip codes of the tile: lanes 0..7 serve X, 8..15 the other child, 16..23 the kept child
        constexpr int CL = L::TILE / 16;   // code rows are pitched and zero-padded: a whole tile can always be read
        const int which = lane >> 3, piece = lane & 7;
        const bool tip = which == 0 ? x_tip : (which == 1 ? o_tip : k_tip);
        if (which < 3 && piece < CL && tip) {
            const int tip_row = which == 0 ? d.src_x : (which == 1 ? d.src_o : d.src_k);
            cp_async16(st + U::CODES_OFF + which * L::TILE + piece * 16,
                       p.codes + (size_t)tip_row * p.pitch + (size_t)t * L::TILE + piece * 16);
        }
        if (!o_tip) fetch_block(d.src_o, t, tl);
        if (!k_tip) fetch_block(d.src_k, t, tl + U::TILE_BYTES);
    };
    // registers -> padded staging tile (a dead operand tile) -> coalesced streaming stores into block `blk`
    auto store_block = [&](const double (&v)[PPT][K][4], const int (&e)[PPT], int blk, int t, unsigned char* s_out) {
        PHB_DCHECK(blk >= 0 && t >= 0 && t < n_tiles);
        const size_t site0 = (size_t)t * L::TILE;
        const bool whole = site0 + L::TILE <= S;
        const int valid = whole ? L::TILE : (int)(S - site0);
        unsigned char* dst = clv_bytes + (size_t)blk * block_bytes + site0 * (K * 32);
        int32_t* ex = p.scale + (size_t)blk * S + site0;
#pragma unroll
        for (int h = 0; h < PPT; ++h) {   // 32 patterns at a time
            __syncwarp();                 // the staging tile's previous readers are done
#pragma unroll
            for (int k = 0; k < K; ++k) {
                unsigned char* d = s_out + lane * ROWB + k * 32;
                *reinterpret_cast<double2*>(d) = make_double2(v[h][k][0], v[h][k][1]);
                *reinterpret_cast<double2*>(d + 16) = make_double2(v[h][k][2], v[h][k][3]);
            }
            __syncwarp();
            if (whole) {
                const unsigned char* sl = s_out + lane_off;
                unsigned char* g = dst + (size_t)h * (32 * PIECES * 16) + lane * 16;
#pragma unroll
                for (int j = 0; j < PIECES; ++j)
                    __stcs(reinterpret_cast<int4*>(g + j * 512), *reinterpret_cast<const int4*>(sl + j * (PPR * ROWB)));
                ex[lane + 32 * h] = e[h];
            } else {
#pragma unroll 1
                for (int c = lane; c + h * 32 * PIECES < valid * PIECES && c < 32 * PIECES; c += 32) {
                    const int4 w = *reinterpret_cast<const int4*>(s_out + (c / PIECES) * ROWB + (c % PIECES) * 16);
                    __stcs(reinterpret_cast<int4*>(dst + (size_t)(c + h * 32 * PIECES) * 16), w);
                }
                if (lane + 32 * h < valid) ex[lane + 32 * h] = e[h];
            }
        }
    };

    // v <- (V^-1 down[c]) * (V^T (pi * v)) per category, v = up[c]: the edge's sum table (derivs.cu)
    // (`tip` is warp-uniform: a tip child's first factor is one table row for all categories, and its K products
    // V^-1 . down are skipped by a real branch - half of all children are tips; with the staging code above at a third of
    // its former size the two bodies fit the instruction cache.)
    auto sum_table = [&](double (&v)[PPT][K][4], bool tip, const unsigned char* tl, const unsigned char* codes) {
#pragma unroll
        for (int h = 0; h < PPT; ++h) {
            double y[K][4];
#pragma unroll
            for (int k = 0; k < K; ++k)
#pragma unroll
                for (int m = 0; m < 4; ++m)
                    y[k][m] = fma(p.m2[4 * m + 3], v[h][k][3], fma(p.m2[4 * m + 2], v[h][k][2], fma(p.m2[4 * m + 1], v[h][k][1], p.m2[4 * m] * v[h][k][0])));
            if (tip) {
                double xt[4];
                lds32(s_xtab + (int)(codes[lane + 32 * h] & (NC - 1)) * 32, xt);
#pragma unroll
                for (int k = 0; k < K; ++k)
#pragma unroll
                    for (int m = 0; m < 4; ++m) v[h][k][m] = xt[m] * y[k][m];
            } else {
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    double a[4];
                    lds32(tl + (lane + 32 * h) * ROWB + k * 32, a);
#pragma unroll
                    for (int m = 0; m < 4; ++m) {
                        const double x = fma(p.m1[4 * m + 3], a[3], fma(p.m1[4 * m + 2], a[2], fma(p.m1[4 * m + 1], a[1], p.m1[4 * m] * a[0])));
                        v[h][k][m] = x * y[k][m];
                    }
                }
            }
        }
    };
    // park: coalesced 128-bit stores straight from registers into the warp's own stripe
    auto park_block = [&](const double (&v)[PPT][K][4], const int (&e)[PPT], int slot) {
        PHB_DCHECK(slot >= 0 && slot < p.n_slots);
        unsigned char* dst = my_scratch + (size_t)slot * U::SLOT_BYTES + lane * 16;
#pragma unroll
        for (int h = 0; h < PPT; ++h)
#pragma unroll
            for (int k = 0; k < K; ++k) {
                *reinterpret_cast<double2*>(dst + ((h * K + k) * 2) * 512) = make_double2(v[h][k][0], v[h][k][1]);
                *reinterpret_cast<double2*>(dst + ((h * K + k) * 2 + 1) * 512) = make_double2(v[h][k][2], v[h][k][3]);
            }
        int* ex = reinterpret_cast<int*>(my_scratch + (size_t)slot * U::SLOT_BYTES + U::BLOCK_BYTES + lane * (4 * PPT));
#pragma unroll
        for (int h = 0; h < PPT; ++h) ex[h] = e[h];
    };

    int tile = blockIdx.x;
    if (tile >= n_tiles) return;
    // prologue: descriptors of steps 0 and 1 (two 16-byte halves each), then the inputs of step 0
    if (lane < 4) {
        const int s = (lane >> 1) < n_steps ? (lane >> 1) : 0;
        cp_async16(reinterpret_cast<unsigned char*>(s_desc) + lane * 16,
                   reinterpret_cast<const unsigned char*>(p.steps + s) + (lane & 1) * 16);
    }
    cp_async_commit();
    cp_async_wait_all();
    __syncwarp();
    stage_step(s_desc[0], tile, 0);
    cp_async_commit();

    double prev[PPT][K][4], oth[PPT][K][4];
    int pe[PPT], oe[PPT], ed_o[PPT], ed_k[PPT];
#pragma unroll
    for (int q = 0; q < PPT; ++q) {
        pe[q] = oe[q] = ed_o[q] = ed_k[q] = 0;
#pragma unroll
        for (int k = 0; k < K; ++k)
#pragma unroll
            for (int i = 0; i < 4; ++i) prev[q][k][i] = oth[q][k][i] = 0.0;
    }
    int step = 0, q = 0;
    int step2 = n_steps > 2 ? 2 : 0;   // two steps ahead (descriptors do not depend on the tile)
    while (true) {
        int step_n = step + 1, tile_n = tile;
        if (step_n == n_steps) {
            step_n = 0;
            tile_n += wstride;
        }
        const bool has_next = tile_n < n_tiles;
        cp_async_wait_all();   // everything issued one step ago has had a whole step to land
        __syncwarp();
        const UpStep* const dcur = &s_desc[q & 3];
        const uint32_t pk = dcur->packed;
        const int dst_o = dcur->dst_o;
        if (lane < 2)
            cp_async16(reinterpret_cast<unsigned char*>(&s_desc[(q + 2) & 3]) + lane * 16,
                       reinterpret_cast<const unsigned char*>(p.steps + step2) + lane * 16);
        if (has_next) stage_step(s_desc[(q + 1) & 3], tile_n, (q + 1) & 1);
        cp_async_commit();

        const unsigned char* st = s_stage + (q & 1) * U::STAGE_BYTES;
        unsigned char* const tile_o = s_tiles + (q & 1) * 2 * U::TILE_BYTES;
        unsigned char* const tile_k = tile_o + U::TILE_BYTES;
        const int mode = (pk >> 24) & 3;
        if (mode == UP_LOAD) {
            // X of the next step comes back from the block array
#pragma unroll
            for (int h = 0; h < PPT; ++h) {
                const unsigned char* r = tile_o + (lane + 32 * h) * ROWB;
#pragma unroll
                for (int k = 0; k < K; ++k) lds32(r + k * 32, prev[h][k]);
                pe[h] = *reinterpret_cast<const int*>(r + K * 32);
            }
        } else if (mode == UP_LOAD_SLOT) {
            // ... or from the warp's scratch stripe
#pragma unroll
            for (int h = 0; h < PPT; ++h) {
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    const unsigned char* src = tile_o + ((h * K + k) * 2) * 512 + lane * 16;
                    const double2 lo = *reinterpret_cast<const double2*>(src);
                    const double2 hi = *reinterpret_cast<const double2*>(src + 512);
                    prev[h][k][0] = lo.x; prev[h][k][1] = lo.y; prev[h][k][2] = hi.x; prev[h][k][3] = hi.y;
                }
                pe[h] = *reinterpret_cast<const int*>(tile_o + U::BLOCK_BYTES + lane * (4 * PPT) + 4 * h);
            }
        } else {
            const int shape = (pk >> 26) & 3;   // other is a tip | kept is a tip << 1
            up_update<K, NC, PPT>(mode == UP_X_TIP, (shape & 1) != 0, (shape & 2) != 0, st, tile_o, tile_k, lane, prev, pe, oth, oe,
                                  ed_o, ed_k);
            if (!ST) {
                store_block(oth, oe, dst_o, tile, tile_o);
                store_block(prev, pe, (int)(pk & 0xffffff), tile, tile_k);
            } else {
                const int park = pk >> 28;
                if (park != 15) park_block(oth, oe, park);
                sum_table(oth, (shape & 1) != 0, tile_o, st + U::CODES_OFF + L::TILE);
#pragma unroll
                for (int h = 0; h < PPT; ++h) oe[h] += ed_o[h];
                store_block(oth, oe, dst_o, tile, tile_o);
#pragma unroll
                for (int h = 0; h < PPT; ++h) {
                    oe[h] = pe[h] + ed_k[h];
#pragma unroll
                    for (int k = 0; k < K; ++k)
#pragma unroll
                        for (int i = 0; i < 4; ++i) oth[h][k][i] = prev[h][k][i];
                }
                sum_table(oth, (shape & 2) != 0, tile_k, st + U::CODES_OFF + 2 * L::TILE);
                store_block(oth, oe, (int)(pk & 0xffffff), tile, tile_k);
            }
        }
        if (!has_next) break;
        step = step_n;
        tile = tile_n;
        if (++step2 == n_steps) step2 = 0;
        ++q;
    }
    cp_async_wait_all();
}

template <int K, int NC, int PPT, bool ST>
int launch_up(Ctx* c, int n_steps, int n_slots) {
    using U = UpLayout<K, NC, PPT>;
    using L = PairLayout<K, NC, PPT>;
    UpArgs a{};
    a.steps = reinterpret_cast<const UpStep*>(c->d_up_rows);
    a.n_steps = n_steps;
    a.opbase = reinterpret_cast<const unsigned char*>(c->d_pmats);
    a.codes = c->d_codes;
    a.pitch = c->code_pitch;
    a.clv = c->d_clv;
    a.scale = c->d_scale;
    a.S = c->S;
    a.n_tiles = (c->S + L::TILE - 1) / L::TILE;
    a.lut = c->d_lut;
    a.scratch = c->d_scratch;
    a.n_slots = n_slots;
    if (ST) {
        for (int m = 0; m < 4; ++m)
            for (int i = 0; i < 4; ++i) {
                a.m1[4 * m + i] = c->h_ivecs[4 * m + i];
                a.m2[4 * m + i] = c->h_evecs[4 * i + m] * c->h_freqs[i];
            }
    }
    auto kern = dna_up_kernel<K, NC, PPT, ST>;
    const size_t smem = U::WARP_BYTES;
    if (smem > c->smem_optin) return c->fail(PHB_ERR_UNSUPPORTED, "up kernel: does not fit in shared memory");
    PHB_CUDA(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    PHB_CUDA(c, cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    int per_sm = 0;
    PHB_CUDA(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 32, smem));
    if (per_sm < 1) per_sm = 1;
    int64_t grid = std::max<int64_t>(1, std::min<int64_t>(a.n_tiles, (int64_t)c->sm_count * per_sm));
    if (ST) {   // every resident warp needs its own scratch stripe
        const int64_t cap = (int64_t)(c->scratch_bytes / ((size_t)n_slots * U::SLOT_BYTES));
        if (c->d_scratch == nullptr || cap < 1) return c->fail(PHB_ERR_NOMEM, "up kernel: scratch area too small");
        grid = std::min(grid, cap);
    }
    // a tile is one long job (the whole tree): every warp gets the same number of them - the launch takes
    // ceil(tiles / warps) rounds either way, and fewer co-resident warps finish a round sooner
    const int64_t rounds = (a.n_tiles + grid - 1) / grid;
    grid = (a.n_tiles + rounds - 1) / rounds;
    kern<<<(int)grid, 32, smem, c->stream>>>(a);
    c->launches++;
    PHB_CUDA(c, cudaGetLastError());
    c->resident_warps = per_sm;
    c->resident_slots = n_slots;
    return PHB_OK;
}

}  // namespace

// Pre-order pass as one walk.  PHB_ERR_UNSUPPORTED (no message) when the shape is not covered: the caller falls back
// to the two-rows-per-parent form.
int dna_up_walk(Ctx* c, int node_a, int node_b) {
    if (!dna_supported(c) || c->K > 4 || !tip_tables_usable(c) || c->d_up_rows == nullptr) return PHB_ERR_UNSUPPORTED;
    const int n_rows = c->n_rows();
    if (n_rows == 0) return PHB_OK;
    const int K = c->K;
    const size_t tab_bytes = (size_t)K * tip_table_rows(c) * 32, p_bytes = (size_t)K * 128;
    const size_t tab_base = reinterpret_cast<const unsigned char*>(c->d_tiptab) - reinterpret_cast<const unsigned char*>(c->d_pmats);
    const int root_p = 2 * c->max_rows() + 1;   // P(root edge length), built by prepare_root
    auto is_tip = [&](int node) { return c->node_tip[node] >= 0; };
    auto oper_off = [&](int node, int pidx) {
        return (uint32_t)((is_tip(node) ? tab_base + (size_t)pidx * tab_bytes : (size_t)pidx * p_bytes) / 16);
    };
    // internal nodes below each node (subtree weight): the walk descends into the lighter internal child first
    std::vector<int> weight(c->n_nodes, 0);
    for (int r = 0; r < n_rows; ++r)
        weight[c->rows_raw[3 * r]] = 1 + weight[c->rows_raw[3 * r + 1]] + weight[c->rows_raw[3 * r + 2]];
    // sum-table form needs the host copy of the eigen-system and the scratch area; PHB_UP_PLAIN keeps plain up partials
    const bool st_form = !tuning().up_plain && c->d_scratch != nullptr && (int)c->h_evecs.size() == 16 &&
                         (int)c->h_ivecs.size() == 16 && (int)c->h_freqs.size() == 4;
    std::vector<UpStep> steps;
    steps.reserve(2 * (size_t)n_rows);
    struct Pending {
        int par;
        bool x_in_regs;
        int slot;   // scratch slot X is parked in (sum-table form), -1: X comes from the block array
    };
    std::vector<Pending> stack;
    int depth = 0, n_slots = 1;
    // the two ends of the root edge, each seeing the other end's down partial through P(root length)
    for (int pass = 1; pass >= 0; --pass) {
        const int par = pass == 0 ? node_a : node_b;
        if (!is_tip(par)) stack.push_back({par, false, -1});
    }
    while (!stack.empty()) {
        const Pending cur = stack.back();
        stack.pop_back();
        const int par = cur.par, r = c->node_row[par];
        if (r < 0) return c->fail(PHB_ERR_STATE, "up partials: node is not computed by the schedule");
        const bool at_root = par == node_a || par == node_b;
        const int above = at_root ? (par == node_a ? node_b : node_a) : -1;
        int mode = UP_X_PREV, x_pidx, src_x = 0;
        uint32_t off_x;
        if (at_root) {
            x_pidx = root_p;
            if (is_tip(above)) {
                mode = UP_X_TIP;
                src_x = c->node_tip[above];
            }
            off_x = oper_off(above, x_pidx);
        } else {
            const int gp = c->node_parent[par];
            if (gp < 0) return c->fail(PHB_ERR_STATE, "up partials: the given root edge does not match the schedule");
            const int rq = c->node_row[gp];
            x_pidx = 2 * rq + (c->rows_raw[3 * rq + 1] == par ? 0 : 1);
            off_x = (uint32_t)((size_t)x_pidx * p_bytes / 16);
        }
        if (mode == UP_X_PREV && !cur.x_in_regs) {
            UpStep ld{};
            if (cur.slot >= 0) {
                ld.src_x = cur.slot;
                ld.packed = (uint32_t)UP_LOAD_SLOT << 24;
                depth = cur.slot;   // LIFO: everything parked after it has been consumed
            } else {
                ld.src_x = at_root ? c->node_slot[above] : c->n_internal + par;
                ld.packed = (uint32_t)UP_LOAD << 24;
            }
            steps.push_back(ld);
        }
        const int ch[2] = {c->rows_raw[3 * r + 1], c->rows_raw[3 * r + 2]};
        // kept child: the internal one; the lighter one if both are
        int ki = 1;
        if (!is_tip(ch[0]) && (is_tip(ch[1]) || weight[ch[0]] < weight[ch[1]])) ki = 0;
        const int kept = ch[ki], other = ch[1 - ki];
        UpStep s{};
        s.off_x = off_x;
        s.off_o = oper_off(other, 2 * r + (1 - ki));
        s.off_k = oper_off(kept, 2 * r + ki);
        s.src_x = src_x;
        s.src_o = is_tip(other) ? c->node_tip[other] : c->node_slot[other];
        s.src_k = is_tip(kept) ? c->node_tip[kept] : c->node_slot[kept];
        s.dst_o = c->n_internal + other;
        const int dst_k = c->n_internal + kept;
        if (dst_k >= (1 << 24)) return c->fail(PHB_ERR_UNSUPPORTED, "up kernel: too many nodes");
        int park = 15;
        if (st_form && !is_tip(other)) {
            park = depth++;
            n_slots = std::max(n_slots, depth);
            if (park >= 15) return PHB_ERR_UNSUPPORTED;   // deeper than 2^15 internal nodes allow; the two-row form takes over
        }
        s.packed = (uint32_t)dst_k | ((uint32_t)mode << 24) | ((uint32_t)is_tip(other) << 26) | ((uint32_t)is_tip(kept) << 27) |
                   ((uint32_t)park << 28);
        steps.push_back(s);
        // LIFO: the kept child is popped first and finds its X in the registers
        if (!is_tip(other)) stack.push_back({other, false, st_form ? park : -1});
        if (!is_tip(kept)) stack.push_back({kept, true, -1});
    }
    if (steps.size() > 2 * (size_t)c->max_rows()) return c->fail(PHB_ERR_STATE, "up kernel: step table overflow");
    PHB_CUDA(c, cudaMemcpyAsync(c->d_up_rows, steps.data(), steps.size() * sizeof(UpStep), cudaMemcpyHostToDevice, c->stream));
    PHB_CUDA(c, cudaStreamSynchronize(c->stream));   // `steps` is a stack object
    const int n_steps = (int)steps.size();
    // one pattern per lane: smaller tiles, 6-7 warps per SM instead of 4 (measured faster at every size tried;
    // PHB_UP_PPT=2 selects two patterns per lane)
    const int ppt = K == 4 && tuning().up_ppt != 2 ? 1 : 2;
    c->up_sumtable = st_form;
    switch (K * 1000 + tip_table_rows(c) * 10 + ppt) {
#define PHB_UP_CASE(K_, NC_, PPT_) \
    case K_ * 1000 + NC_ * 10 + PPT_: \
        return st_form ? launch_up<K_, NC_, PPT_, true>(c, n_steps, n_slots) : launch_up<K_, NC_, PPT_, false>(c, n_steps, n_slots);
        PHB_UP_CASE(1, 8, 2)
        PHB_UP_CASE(1, 16, 2)
        PHB_UP_CASE(2, 8, 2)
        PHB_UP_CASE(2, 16, 2)
        PHB_UP_CASE(4, 8, 2)
        PHB_UP_CASE(4, 16, 2)
        PHB_UP_CASE(4, 8, 1)
        PHB_UP_CASE(4, 16, 1)
#undef PHB_UP_CASE
    }
    return PHB_ERR_UNSUPPORTED;
}

}  // namespace phb
