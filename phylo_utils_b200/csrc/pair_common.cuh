// Device helpers shared by the two-patterns-per-lane 4-state walks (clv_dna_pair.cu, up_dna_pair.cu):
// asynchronous-copy wrappers, the per-warp shared-memory layout, tip-code decoding.
#pragma once

#include "common.cuh"
#include "resident_plan.cuh"

namespace phb {
namespace {

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gmem_src) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(s), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

template <int K, int NC, int PPT>
struct PairLayout {
    static constexpr int P_ROUNDS = (K * 128 + 511) / 512;        // warp-wide 512-byte copy rounds of a P block
    static constexpr int T_ROUNDS = (K * NC * 32 + 511) / 512;    // ... of a tip table [k][code][4 doubles]
    static constexpr int ROUNDS = P_ROUNDS > T_ROUNDS ? P_ROUNDS : T_ROUNDS;
    static constexpr int OPER_BYTES = ROUNDS * 512;
    static constexpr int TILE = 32 * PPT;                         // patterns per warp tile
    static constexpr int CODES_OFF = 2 * OPER_BYTES;              // TILE bytes of tip codes per operand (half when packed)
    static constexpr int STAGE_BYTES = 2 * OPER_BYTES + 2 * TILE;
    static constexpr int DESC_BYTES = 4 * 16;                     // descriptor ring: rows r .. r+2 in flight
    static constexpr int CHUNKS = PPT * K * 2;                    // 16-byte chunks per lane in a parked block
    static constexpr int BLOCK_BYTES = CHUNKS * 512;              // [pattern of the lane][k][half][lane]
    static constexpr int EXP_STRIDE = PPT == 1 ? 4 : (PPT == 2 ? 8 : 16);   // bytes of exponents per lane (PPT = 3: one word unused)
    static constexpr int SLOT_BYTES = BLOCK_BYTES + 32 * EXP_STRIDE;        // + PPT exponents per lane
    static constexpr int WARP_BYTES = DESC_BYTES + 2 * STAGE_BYTES + SLOT_BYTES;
    // CTAs (= warps) per SM the register file is budgeted for.  A warp lives in one of the four SM
    // sub-partitions with 16384 registers each: 12 CTAs = 3 warps per sub-partition = 168 registers.
    // Four patterns per lane (K <= 4) double the register footprint: 8 CTAs = 2 warps per sub-partition = 255 registers.
    // Three patterns per lane (96-pattern tiles, K = 4): 9 CTAs = 224 registers.
    static constexpr int MIN_CTAS = PPT == 2 ? (K <= 2 ? 16 : (K <= 4 ? 12 : 4)) : (PPT == 3 ? 9 : (K <= 1 ? 16 : (K <= 2 ? 12 : 8)));
};

// Which patterns of its tile a lane owns, and how a parked operand tile is laid out in shared memory:
//   LAYOUT_PRIVATE  lane l owns patterns PPT*l .. PPT*l + PPT-1; the operand tile is the warp's private chunk layout
//   LAYOUT_ARRAY    lane l owns patterns l, l + 32, ...; the operand tile mirrors the caller-visible partials array
//                   (pattern-major rows, padded to ROWB bytes so that 128-bit accesses are conflict-free)
constexpr int LAYOUT_PRIVATE = 0, LAYOUT_ARRAY = 1;

// How the tip codes of a tile reach the kernel (CM): one byte per code; two 4-bit codes per byte; or - look-up tables of
// at most 8 rows, i.e. any alignment without partial ambiguity codes - 3 bits per code split into a plane of 2-bit
// values (the low bits, TILE / 4 bytes per tile) and a plane of single bits (the high bit, TILE / 8 bytes per tile).
constexpr int CODES_BYTE = 0, CODES_NIBBLE = 1, CODES_SPLIT3 = 2;

// the PPT codes of a lane, as byte offsets of their tip-table rows
template <int NC, int PPT, int CM, int LAYOUT>
__device__ __forceinline__ void table_rows(const unsigned char* codes, int lane, int (&row)[PPT]) {
    if (LAYOUT == LAYOUT_ARRAY) {
#pragma unroll
        for (int p = 0; p < PPT; ++p) row[p] = (int)(codes[lane + 32 * p] & (NC - 1)) * 32;
        return;
    }
    if (CM == CODES_SPLIT3) {
        // low plane at [0, TILE / 4), high plane behind it; the lane's PPT patterns are adjacent
        constexpr int TILE = 32 * PPT;
        const unsigned lo = PPT == 2 ? ((unsigned)codes[lane >> 1] >> (4 * (lane & 1))) & 15u : (unsigned)codes[lane];
        const unsigned hi = PPT == 2 ? ((unsigned)codes[TILE / 4 + (lane >> 2)] >> (2 * (lane & 3))) & 3u
                                     : ((unsigned)codes[TILE / 4 + (lane >> 1)] >> (4 * (lane & 1))) & 15u;
#pragma unroll
        for (int p = 0; p < PPT; ++p) row[p] = (int)((((lo >> (2 * p)) & 3u) | (((hi >> p) & 1u) << 2)) & (NC - 1)) * 32;
        return;
    }
    constexpr bool PACKED = CM == CODES_NIBBLE;
    unsigned raw;
    if (PPT == 3) {
        // patterns 3 lane .. 3 lane + 2: three bytes, or three nibbles starting at nibble 3 lane (two bytes cover them)
        if (PACKED) {
            const int n0 = 3 * lane;
            raw = ((unsigned)codes[n0 >> 1] | ((unsigned)codes[(n0 >> 1) + 1] << 8)) >> (4 * (n0 & 1));
        } else {
            raw = (unsigned)codes[3 * lane] | ((unsigned)codes[3 * lane + 1] << 8) | ((unsigned)codes[3 * lane + 2] << 16);
        }
    } else if (PACKED) raw = PPT == 2 ? (unsigned)codes[lane] : (unsigned)*reinterpret_cast<const unsigned short*>(codes + 2 * lane);
    else raw = PPT == 2 ? (unsigned)*reinterpret_cast<const unsigned short*>(codes + 2 * lane)
                        : *reinterpret_cast<const unsigned*>(codes + 4 * lane);
#pragma unroll
    for (int p = 0; p < PPT; ++p) row[p] = (int)((raw >> ((PACKED ? 4 : 8) * p)) & (NC - 1)) * 32;
}

__device__ __forceinline__ void lds32(const unsigned char* p, double (&v)[4]) {
    const double2 lo = *reinterpret_cast<const double2*>(p);
    const double2 hi = *reinterpret_cast<const double2*>(p + 16);
    v[0] = lo.x; v[1] = lo.y; v[2] = hi.x; v[3] = hi.y;
}

}  // namespace
}  // namespace phb
