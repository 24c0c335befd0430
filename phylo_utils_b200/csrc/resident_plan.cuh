// Parking plan of the operand-resident 4-state walks (clv_dna_pair.cu, up_dna_pair.cu).
#pragma once

#include <vector>

#include "common.cuh"

namespace phb {

constexpr int KIND_TIP = 0, KIND_PREV = 1, KIND_SLOT = 2;
constexpr int kScratchSlots = 15;

// 16-byte row descriptor
struct __align__(16) ResRow {
    int32_t src_a;   // tip row | parked-block id (scratch slot, or producer row in STORE mode)
    int32_t src_b;
    int32_t pidx_a;  // P block of operand a
    uint32_t packed; // pidx_b [0:24) | kind_a [24:26) | kind_b [26:28) | dst slot [28:32) (15 = not parked)
};

struct ResPlan {
    std::vector<ResRow> rows;
    int n_slots = 0;
};

// Walks the schedule like a register allocator (defined in clv_dna_pair.cu).  Operands come out in the
// canonical order TIP <= PREV <= SLOT; rows with two parked operands are rejected (PHB_ERR_UNSUPPORTED).
int plan_rows(Ctx* c, int root_a, int root_b, bool with_root, bool store, ResPlan* out);

}  // namespace phb
