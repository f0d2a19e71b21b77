// qxb200 -- host-side builder of row programs (see qxb_rowprog.h).
#pragma once
#include <string>
#include <vector>

#include "qxb_ir.h"
#include "qxb_rowprog.h"

namespace qxb {

struct RowPlanOptions {
    int min_tt_bits = 6;          // keep at least 2^min_tt_bits thread-tiles per op when choosing the register tile
    int max_tile_bits = 4;        // ma + nb <= this (each side <= 2)
    int tile_reg_budget = 100;    // 32-bit registers for staged operands + accumulators (chooses the K chunk)
    bool alap = true;             // schedule every op as late as its consumers allow
    bool stage_shared = true;     // copy operands shared by all rows into the arena one level before their first use
    int stage_max_bits = 12;      // ... when they span at most 2^n elements
    bool dmma = true;             // ComplexF64 nodes with >= 3 M-only / N-only bits: 8 x 8 tiles on the FP64 tensor pipe
    double chain_min_macs = 2048; // fused chain: complex MACs per row below which an end node is not worth a level
    long long max_arena_bytes = 200 * 1024;
    bool bank_opt = true;             // choose the low lane bits of a unit (and chain-intermediate layouts) by the bank-conflict model
    int chain_side = 0;               // select_chain: depth of side branches taken into the fused program (0 = a pure path)
    bool bank_search_tiles = false;   // also choose WHICH M-only / N-only bits form the register tile by the bank-conflict model
                                      // (a few 10^5 evaluations per op: for the handful of ops of a fused chain or a ring launch)
};

struct RowProgramHost {
    bool ok = false;
    std::string why;              // why not, when !ok
    Phase phase = PH_CHUNK;
    std::vector<RowOp> ops;       // descriptors; g* pointers and fixed-variable offsets are filled in per launch
    std::vector<int> lop;         // index into Lowered::ops (-1: copy pseudo-op of a staged operand, source = ref_a)
    std::vector<int> ref_a, ref_b, ref_c;          // LTensor indices
    std::vector<char> in_arena_a, in_arena_b, in_arena_c;
    std::vector<RowUnit> units;
    std::vector<int> level_start; // n_levels + 1
    int n_levels = 0;
    std::vector<RowLeaf> leaves;
    int arena_elems = 0;          // per row (chunk phase)
    int elem_bytes = 16;
    int root_off = 0, root_span = 0;
    double flops_per_row = 0;     // 8 * complex MACs
    double elems_per_row_amp = 0; // algorithmic elements of per-row tensors (A, B, C of every op)
    double elems_shared = 0;      // algorithmic elements of operands shared by all rows
};

// subset != nullptr (chunk phase only): a program for just these ops (indices into L.ops, ascending) -- a FUSED CHAIN.
// Per-row tensors produced outside the subset are staged from global memory (row stride 2^span_bits), results that are
// consumed outside go to global memory; no output leaves, no root reduction.
RowProgramHost build_row_program(const Lowered& L, Phase phase, int dtype, const RowPlanOptions& o,
                                 const std::vector<int>* subset = nullptr);

// The chain of chunk-phase contractions worth fusing: a maximal path op -> its only consumer -> ... through nodes
// whose rows are small, chosen by total FLOPs (the dominant contractions of the elimination-order-like plans are such
// a chain: a growing intermediate times one small side operand after the other).  Empty when there is none.
std::vector<int> select_chain(const Lowered& L, int dtype, const RowPlanOptions& o);
// Move the ops of `chain` (ascending indices) so that they are contiguous at the position of the last one -- legal
// because nothing outside the chain reads their intermediates -- and recompute dependency and use indices.
// Returns the new indices of the chain ops.  Call BEFORE plan_memory.
std::vector<int> make_contiguous(Lowered& L, const std::vector<int>& chain);

// What the kernel reads: one resolved descriptor per unit and the slot table (levels padded to a multiple of
// kRowWarps slots).  `ops` = rp.ops with oA/oB/oC final (fixed-variable offsets XORed in, 0 for global tensors) and
// gA/gB/gC set for tensors outside the arena.
struct RowDeviceTables {
    std::vector<RowUnitDesc> descs;
    std::vector<uint16_t> slots;
    std::vector<int> level_start;     // in slots, n_levels + 1
    std::vector<int> desc_op;         // op index (into rp.ops) of each descriptor
};
RowDeviceTables build_row_tables(const RowProgramHost& rp, const std::vector<RowOp>& ops);

// Unit descriptors of ONE contraction for the ring kernel (lane bases relative to each operand's own region);
// empty when the op has no tile variant / fewer than 32 thread-tiles.
std::vector<RowUnitDesc> build_ring_descs(const LOp& op, int dtype, const RowPlanOptions& o, std::string& why);

}  // namespace qxb
