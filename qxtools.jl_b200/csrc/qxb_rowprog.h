// qxb200 -- "row programs": a whole phase of the lowered contraction tree executed by ONE persistent kernel.
//
// Why: in the batched lowering (qxb_lower.cpp) every chunk-phase tensor is a small dense 2^n array PER BITSTRING ROW
// (RQC 7x7 d20, 2^12 slices batched: 100 contractions, largest intermediate 2^11 elements = 32 KB, 76 KB live at
// the peak).  Launching one kernel per contraction streams every intermediate through HBM (62 GB per 131072
// bitstrings, 14 ms at 63 % of the HBM peak).  A B200 SM has 227 KB of shared memory: the live set of a row FITS.
// So one CTA takes one bitstring row through the ENTIRE chunk phase with every intermediate in shared memory:
//     HBM traffic per row = its bitstring (n_qubits bytes) + one amplitude; the step becomes FP64-pipe-bound.
// The same interpreter runs the block phase (slice-only nodes: ~250 tiny contractions, pure launch latency as
// separate kernels) as a single-CTA program over global memory.
//
// Replaces, for the rows it covers, the inner loops of QXContexts.execute (call site
// /root/reference/bin/qxrun.jl:83-87): "for bitstring: for slice: for ncon" -> one launch.
//
// Program = ops grouped into dependency LEVELS; the work of a level is cut into warp-sized UNITS (32 thread-tiles
// of one op) that the warps of the CTA take round-robin; one __syncthreads per level.
#pragma once
#include <stddef.h>
#include <stdint.h>

namespace qxb {

constexpr int kRowThreads = 256;
constexpr int kRowWarps = kRowThreads / 32;
constexpr int kRowMaxLevels = 96;
constexpr int kRowMaxSeg = 12;          // (src, dst, len) runs of the thread-tile-index -> address maps
constexpr int kRowMaxKSeg = 8;          // runs of the k-bits >= 4 -> address maps

struct RSeg { unsigned char src, dst, len, pad; };

// kinds of inner loop
constexpr int kRowKindKred = 0;         // <= 16 thread-tiles: lanes also split K, shuffle-reduced
constexpr int kRowKindDmma = 254;       // ComplexF64 on the FP64 tensor pipe: a warp computes 8 x 8 output tiles with
                                        // mma.sync.m8n8k4.f64 (4 real MMAs per complex k4 step); see RowUnitDesc
constexpr int kRowKindCopy = 255;       // staging: 2^ntt elements from global memory (gA) into the arena at lC[0] | lC[1] << 16
// tile kinds: 1 + (((ma * 3 + nb) * 3 + kc) * 2 + gen)
inline int row_tile_kind(int ma, int nb, int kc, int gen) { return 1 + (((ma * 3 + nb) * 3 + kc) * 2 + gen); }

// HOT part (128 bytes): the tables the inner loops index; read from the warp's shared-memory slot (broadcasts).  Offsets are ELEMENT indices relative to the
// tensor (all < 2^16: row programs are only built for tensors of <= 2^16 elements), made of disjoint bits, so
// they simply ADD; the per-lane bases (RowUnitDesc) carry the arena offset, so a tensor may sit anywhere in the arena.
struct alignas(16) RowOpHot {
    uint16_t aT[4], bT[4];      // register-tile offsets into A (M-only bits) / B (N-only bits)
    uint16_t cT[16];            // register-tile offsets into C, index jm * 2^nb + jn
    uint16_t ktA[16], ktB[16];  // offsets of the low min(nK, 4) bits of k
    uint8_t nK, kc, ma, nb;     // K bits; log2 K chunk held in registers; tile bits
    uint8_t ntt, kind, ks, gen; // log2 #thread-tiles; inner-loop kind; K-split bits (kred); 1 = some tensor is global
};
static_assert(sizeof(RowOpHot) == 128, "hot descriptor must be 128 bytes");

struct alignas(16) RowOp {
    RowOpHot hot;
    // COLD part (host side only since the unit descriptors exist): address maps and tensor locations
    unsigned long long gA, gB, gC;   // global base pointer, or 0 when the tensor lives in the row arena
    long long rsA, rsB, rsC;         // elements between consecutive rows of a global per-row tensor (0 = shared)
    int oA, oB, oC;                  // arena element offset when g* == 0
    uint8_t nsA, nsB, nsC, nkA, nkB;
    uint8_t lsA, lsB, lsC;           // log2 row stride of global per-row tensors (kRowShared: none)
    RSeg tA[kRowMaxSeg], tB[kRowMaxSeg], tC[kRowMaxSeg];   // thread-tile (kRowKindDmma: warp-tile) index bits -> address bits
    RSeg kA[kRowMaxKSeg], kB[kRowMaxKSeg];                 // (k >> 4) bits -> address bits
    uint16_t frA[32], frB[32], frC[32];                    // kRowKindDmma: per-lane fragment offsets (without oA / oB / oC)
    uint16_t c_n1;                                         // kRowKindDmma: C offset of n + 1 (the lane's second output)
};

struct RowUnit { uint16_t op, chunk; };      // host side: chunk = which group of 32 thread-tiles of the op

// What the kernel executes: one descriptor PER UNIT (32 thread-tiles of one op), fully resolved on the host -- the
// per-lane base offsets replace the segment evaluation the first version did per lane and unit (3 x ~8 segments x
// ~12 instructions: half of the 80 000 warp instructions per row that version executed, profiles/r2d_summary.md).
// 416 bytes = 26 cp.async of 16 bytes into the warp's shared-memory slot, one unit ahead of the compute.
struct alignas(16) RowUnitDesc {
    RowOpHot hot;
    unsigned long long gA, gB, gC;           // global base pointers (0: the tensor lives in the row arena)
    RSeg kA[kRowMaxKSeg], kB[kRowMaxKSeg];   // (k >> 4) bits -> address bits (only read when nK > 4)
    uint8_t nkA, nkB;
    uint8_t lsA, lsB, lsC;                   // log2 of the row stride (elements) of a global PER-ROW tensor (fused chains:
                                             // inputs produced by other kernels, the chain's result); kRowShared = one
                                             // tensor for all rows
    uint8_t pad[3];
    uint16_t lA[32], lB[32], lC[32];         // per lane: base element offsets of its thread-tile (arena offset folded
                                             // in); lC == 0xFFFF: the lane has no thread-tile in this unit
                                             // kRowKindDmma: the lane's MMA fragment element -- lane = 4 g + t holds
                                             // A[m = g][k = t], B[k = t][n = g], C[m = g][n = 2 t], [2 t + 1]; the unit's
                                             // 2^hot.ma tiles add hot.aT / bT / cT[tile]; hot.cT[4] = offset of n + 1;
                                             // k = 4 s + t, step s adds hot.ktA / ktB[4 s]
};
static_assert(sizeof(RowUnitDesc) == 416, "unit descriptor must be 26 x 16 bytes");
constexpr uint16_t kRowNull = 0xFFFF;        // slot table: no unit for this warp in this round
constexpr uint8_t kRowShared = 0xFF;         // lsA / lsB / lsC: the global tensor does not depend on the row

struct RowLeaf { int off, span_bits, out_idx; };   // output leaf materialised in the arena from the row's bitstring

// kernel argument (by value, __grid_constant__)
struct RowLaunch {
    const RowUnitDesc* descs;       // unit descriptors
    const uint16_t* slots;          // slot table: per level a multiple of kRowWarps entries; warp w takes entries
                                    // level_start[lv] + w, + kRowWarps, ...; value = descriptor index or kRowNull
    const RowLeaf* leaves;
    const unsigned char* bits;      // [n_rows_total][n_outputs] bitstring bytes (0, 1, 2 = '+', 3 = '-')
    double* acc;                    // [n_rows_total] complex double accumulators (closed network), or nullptr
    long long amp0, n_rows;         // rows [amp0, amp0 + n_rows)
    double scale;                   // root_scale
    long long* timing;              // diagnostics: per-level clock cycles of CTA 0, summed over its rows (or nullptr)
    int slots_bytes;                // bytes reserved for the slot table in shared memory (multiple of 128)
    int n_levels, n_leaves, n_outputs;
    int root_off, root_span;        // root tensor in the arena (summed over its 2^root_span elements into acc)
    int level_start[kRowMaxLevels + 1];     // in slots
};

// ---------------------------------------------------------------------------------------------------------------
// Ring kernel: ONE contraction whose operand rows A[u], B[u] and result row C[u] are small dense per-bitstring rows
// (the dominant nodes of the batched RQC plans: 32 KB + 8 KB -> 32 KB per row, K = 8).  Persistent CTA per SM;
// the rows travel by 1-D TMA bulk copies: cp.async.bulk global -> shared into a ring of stages (mbarrier
// complete_tx), the same unit interpreter computes the row from shared memory INTO shared memory, and the result
// row leaves by cp.async.bulk shared -> global (bulk_group).  No thread ever touches global memory, so the
// thread <-> element mapping is free (2 x 2-bit register tiles on any M / N bits) and the bytes in flight cost no
// registers -- what limited contract_kernel to 55-63 % of the HBM peak on these nodes (profiles/r1p_summary.md).
struct RingLaunch {
    const RowUnitDesc* descs;       // n_units descriptors; lane bases relative to the operand's own region
    const void* A;
    const void* B;
    void* C;
    long long sUA, sUB, sUC;        // elements between consecutive rows (0 = operand shared by all rows)
    long long U;                    // rows
    int n_units;
    int nA, nB, nC;                 // elements per row
    int stages;                     // 2..4
};
const void* ring_func(int dtype);
constexpr int kRingMaxStages = 4;
// dynamic shared memory of the ring kernel
inline size_t ring_smem_bytes(int n_units, int nA, int nB, int nC, bool sharedA, bool sharedB, int stages, size_t es) {
    const size_t head = ((size_t)n_units * sizeof(RowUnitDesc) + 8 * kRingMaxStages + 127) / 128 * 128;
    const size_t stage = ((sharedA ? 0 : (size_t)nA) + (sharedB ? 0 : (size_t)nB) + (size_t)nC) * es;
    return head + ((sharedA ? (size_t)nA : 0) + (sharedB ? (size_t)nB : 0)) * es + (size_t)stages * stage;
}

// entry point: (RowLaunch by value); dynamic shared memory = row_smem_bytes(...)
const void* rowprog_func(int dtype);
inline size_t row_slots_bytes(size_t n_slots) { return (n_slots * sizeof(uint16_t) + 127) / 128 * 128; }
inline size_t row_fixed_smem_bytes(size_t n_slots) { return 2 * kRowWarps * sizeof(RowUnitDesc) + row_slots_bytes(n_slots); }

}  // namespace qxb
