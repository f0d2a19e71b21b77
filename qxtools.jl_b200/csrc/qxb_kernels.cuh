// qxb200 -- device kernels (sm_100a).
//
// Every tensor is a 2^n array of interleaved complex numbers; a contraction node
// is   C[u][c] = sum_k A[u*sUA + fA(c) + gA(k)] * B[u*sUB + fB(c) + gB(k)]
// where u is the bitstring (amplitude) row, c runs over the bits of C (tensor
// modes AND batched slice-variable bits), and fA/fB/gA/gB are bit-scatter maps
// given as a handful of (src, dst, len) segments.  C inherits the bit order of
// its larger operand, so consecutive threads write consecutive C elements and
// read (near-)consecutive A elements: the kernel streams at HBM rate for the
// "big x small" nodes that carry the bytes.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace qxb {

constexpr int kMaxSeg = 32;
constexpr int kMaxKSeg = 32;
constexpr int kKTab = 16;
constexpr int kThreads = 256;
constexpr int kMaxEpt = 4;          // elements per thread per tile -> tile of up to 1024 elements

struct DSeg { unsigned char src, dst, len, pad; };

struct OpParams {
    const void* A;
    const void* B;
    void* C;
    long long sUA, sUB, sUC;        // elements between consecutive amplitude rows (0 = shared)
    long long tiles;                // U << hb
    int nC, nK, U;
    int lob;                        // low C bits covered by the threads of a block (<= 8)
    int ma, nb;                     // register tile: 2^ma M-only bits x 2^nb N-only bits per thread
    int kc;                         // log2 of the K chunk staged in registers per step (<= 3, <= nK)
    int hb;                         // remaining ("hi") C bits, enumerated by the tile index
    int aBits, bBits;               // log2 of the stored elements of A and B per bitstring row
    int nsAlo, nsBlo, nsClo, nsAhi, nsBhi, nsChi, nkA, nkB;
    long long ktabA[kKTab], ktabB[kKTab];   // offsets of the low min(nK,4) bits of k
    long long aT[4], bT[4];         // register-tile offsets into A (M bits) and B (N bits)
    long long cT[16];               // register-tile offsets into C, index jm * 2^nb + jn
    DSeg sAlo[8], sBlo[8], sClo[8]; // thread-bit part of the address maps (thread bits are the lowest
                                    // C positions that are not register-tile bits)
    DSeg sAhi[kMaxSeg], sBhi[kMaxSeg], sChi[kMaxSeg];   // hi-index part
    DSeg kA[kMaxKSeg], kB[kMaxKSeg];
};

// Tiled complex GEMM for GEMM-shaped nodes (K >= 2^kcb, >= 5 M-only and >= 5 N-only bits).
// Block tile = 2^tmb M-only bits x 2^tnb N-only bits x 2^kcb K bits; every other C bit
// (batch bits, further M/N bits) and the bitstring row are enumerated by the tile index.
struct GemmParams {
    const void* A;
    const void* B;
    void* C;
    long long sUA, sUB, sUC;
    long long tiles;                 // U << hb
    int U, hb;
    int nK;                          // total K bits (chunk bits are the low kcb of the k index)
    int nsAhi, nsBhi, nsChi, nkA, nkB;
    // operand tile loads: load-index bit j -> global offset / shared-memory index contribution
    // (bits sorted by ascending global offset so consecutive threads read ascending addresses)
    long long aLoadOff[12], bLoadOff[12];
    int aLoadSm[12], bLoadSm[12];
    int aLoadSmT[12], bLoadSmT[12];  // same for the tensor-core kernels' shared-memory layouts (gemm_mma_smem_bit / gemm_tc5_smem_bit)
    long long cM[8], cN[8];          // C offset of each M-tile / N-tile bit
    DSeg sAhi[kMaxSeg], sBhi[kMaxSeg], sChi[kMaxSeg];
    DSeg kA[kMaxKSeg], kB[kMaxKSeg]; // k index (all nK bits) -> offsets; chunk c covers k = c << kcb ...
};

// nullptr when the shape has no instantiation; threads = 2^(tmb + tnb - 4)
const void* gemm_func(int dtype, int tmb, int tnb);
int gemm_kcb(int dtype);             // K chunk bits the kernels are built for (4 for c32, 3 for c64)
// tensor-core variant (mma.sync: DMMA for c64, 3xTF32 for c32) on the same GemmParams; nullptr when
// the tile shape has no instantiation (needs tmb == tnb == 6)
const void* gemm_mma_func(int dtype, int tmb, int tnb);
int gemm_mma_threads(int dtype);
size_t gemm_mma_smem_bytes(int dtype);
int gemm_mma_smem_bit(int dtype, bool is_b, int tb, int bit);
// tcgen05 / TMEM ComplexF32 GEMM (qxb_gemm_tc5.cu): block tile 2^7 M-only bits x 2^6 N-only bits x 2^4 K bits,
// 3xTF32 through kind::tf32 UMMAs with TMEM accumulators; same GemmParams, aLoadSmT / bLoadSmT = BYTE offsets of the
// canonical K-major core-matrix layout (gemm_tc5_smem_bit)
const void* gemm_tc5_func();
size_t gemm_tc5_smem_bytes();
int gemm_tc5_threads();
int gemm_tc5_smem_bit(int tile_bits, int bit);

struct OutLeafDesc { long long offset_per_amp; int span_bits; int out_idx; };

// Kernel entry points as function pointers (for cudaLaunchKernel / cudaGraphAddKernelNode).
// contract: one argument (OpParams by value).
const void* contract_func(int dtype, int kc, int ma, int nb, bool single_chunk, int min_blocks = 2);
// shared-memory-staged variant for broadcast-type nodes (nullptr if the shape has none); dynamic
// shared memory = (2^aBits + 2^bBits) * sizeof(element)
const void* contract_smem_func(int dtype, int kc, int ma, int nb, bool single_chunk);
// EXPERIMENT (QXB_SMEM_TMA=1): the same node shape with A[u], B[u] brought in by 1-D TMA bulk copies into a ring of
// stages (nullptr if the shape has no instantiation); arguments (OpParams, int stages)
const void* contract_tma_func(int dtype, int kc, int ma, int nb, bool single_chunk);
// warp-per-output reduction variant for nC <= 8 and long K (same OpParams argument)
const void* kreduce_func(int dtype);
// block-per-output variant for very long K (>= 2^12) and few outputs
const void* kreduce_block_func(int dtype);
// split-K variant: 2^OpParams::kc blocks per output, partial sums combined with atomicAdd into a zeroed C
const void* kreduce_split_func(int dtype);
// tiled split-K variant (qxb_kred.cu): a CTA stages the operand tiles of a K chunk in shared memory once, a thread per
// output accumulates over the chunk, partial sums meet by atomicAdd in a zeroed C.  Arguments (OpParams, KredTile);
// dynamic shared memory = (n_rows_a + n_rows_b) * (kKredTileK + 1) * sizeof(element)
constexpr int kKredTileKBits = 7, kKredTileK = 1 << kKredTileKBits, kKredMaxRows = 64;
struct KredTile {
    int n_rows_a, n_rows_b;                        // distinct operand rows the C bits select (<= kKredMaxRows each)
    unsigned char row_a[256], row_b[256];          // output c -> its row of A / B
    long long off_a[kKredMaxRows], off_b[kKredMaxRows];   // row -> element offset in the operand
    unsigned char grid_c[256];                     // kreduce_grid_kernel: (row of A) * n_rows_b + (row of B) -> output c
    unsigned char kperm_b[8];                      // staging of B: bit j of (tid % chunk) -> bit of the chunk's k (B-ascending)
    int kperm_set;                                 // 0: identity
};
const void* kreduce_tile_func(int dtype);
// Same staging, for outputs that form the full grid rows(A) x rows(B) (both even): a thread owns a 2 x 2 or 4 x 4 block of outputs
// (a half / a quarter of the shared-memory reads per FMA) and the K chunks arrive through a two-stage cp.async ring, so the loads of the
// next chunk overlap the FMAs of this one.  Dynamic shared memory = 2 * (n_rows_a + n_rows_b) * (kKredTileK + 1) * sizeof(element)
const void* kreduce_grid_func(int dtype, int tile);       // tile = 2 or 4 (rows of A and of B both multiples of it)
// "big x small" streaming nodes (qxb_kred.cu): a thread owns one position of the big operand's free index space and all
// 2^n_bits outputs of it (further N bits are enumerated by the CTA index).  Arguments (BigSmallParams); dynamic shared
// memory = the whole small operand
struct BigSmallParams {
    const void* big;
    const void* small_;
    void* C;
    long long sUbig, sUsmall, sUC;     // elements between bitstring rows (0 = shared)
    long long n_pos;                   // position indices = 2^(nC - n_bits): [8 thread bits][nNhi N bits][block bits]
    int U, nK, ntA, ntC, nNhi;
    DSeg tA[16], tC[16];               // position index bits -> address bits of the big operand / of C
    long long aK[32];                  // k -> offset in the big operand
    long long cN[32];                  // n (register tile) -> offset in C
    int bK[32], bN[32];                // k, n -> offset in the small operand
    int bH[256];                       // N bits beyond the register tile -> offset in the small operand
};
const void* bigsmall_func(int dtype, int n_bits, int k_bits, bool packed);     // packed: ComplexF32 through FFMA2 (n_bits <= 4)
// TMA variant (U == 1, the 8 thread bits of the position index = the 8 lowest address bits of the big operand): second
// kernel argument = number of 32 KB stages; dynamic shared memory = stages * kBigSmallStageBytes + small operand
// (rounded up to 16 B) + 8 B per stage (mbarriers)
constexpr int kBigSmallStageBytes = 32 * 1024;
const void* bigsmall_tma_func(int dtype, int n_bits);
// measured FMA-pipe peak of the current device in TFLOP/s (dtype 0: FFMA, 1: DFMA), see qxb_kred.cu
double fma_peak_tflops(int dtype, int num_sms, cudaStream_t st);
// outleaf:  (R2* base, const OutLeafDesc* d, const unsigned char* bits, int n_outputs, long long amp0, long long n)
const void* outleaf_func(int dtype);
// reduce:   (const R2* root, long long sU, int span_bits, long long n, double scale, double* acc, long long amp0)
const void* reduce_root_func(int dtype);
// open (tensor-valued) root: gather of the saved tensor's modes into Julia (column-major) order + sum over the
// batched slice bits still open in it.  reduce_open: (const R2* root, long long sU, long long n, double scale,
// double* acc, long long amp0, RootDesc d); acc holds n_amp x d.elems complex doubles
struct RootDesc {
    long long elems;                 // values per bitstring = prod ext
    int n_modes, n_vseg, vtotal;     // vtotal = sum of vbits
    int ext[32];                     // true extent of each mode, Julia order
    unsigned char pos[32];           // address bit of each mode in the lowered root
    unsigned char vpos[16], vbits[16];
};
const void* reduce_root_open_func(int dtype);
// finalize: (const double* acc, R2* out, long long n)
const void* finalize_func(int dtype);

}  // namespace qxb
