// qxb200 -- row-program interpreter kernel (sm_100a).  See qxb_rowprog.h for the why and the data layout.
//
// One CTA = one bitstring row at a time (persistent, grid-stride over rows): the row's output leaves are written
// into the shared-memory arena from its bitstring, then the levels of the program run with one __syncthreads each,
// every intermediate staying in shared memory, and the root is summed into the row's accumulator.
// Within a level, each warp takes units (32 thread-tiles of one op) round-robin; a thread-tile is a
// 2^ma x 2^nb register tile of outputs, K walked in register chunks of 2^kc.
#include <cuda_runtime.h>

#include "qxb_rowprog.h"

namespace qxb {

namespace {

template <typename R2>
__device__ __forceinline__ void rcmac(R2& acc, const R2 a, const R2 b) {
    acc.x = fma(a.x, b.x, acc.x);
    acc.x = fma(-a.y, b.y, acc.x);
    acc.y = fma(a.x, b.y, acc.y);
    acc.y = fma(a.y, b.x, acc.y);
}

__device__ __forceinline__ int rseg(const RSeg* __restrict__ s, int n, unsigned x) {
    int r = 0;
    for (int i = 0; i < n; ++i) {
        const RSeg g = s[i];
        r |= (int)(((x >> g.src) & ((1u << g.len) - 1u)) << g.dst);
    }
    return r;
}

// registers of the staged operands + accumulators of a tile variant (32-bit registers)
template <typename R2, int MA, int NB, int KC>
struct TileRegs {
    static constexpr int value = (int)(sizeof(R2) / 4) * (((1 << MA) + (1 << NB)) * (1 << KC) + (1 << (MA + NB)));
};

// One thread-tile.  pA / pB / pC: the arena (shared memory, known to the compiler when !GEN) or generic pointers.
template <typename R2, int MA, int NB, int KC, bool GEN>
__device__ __forceinline__ void row_tile(const RowOpHot* __restrict__ h, const RowOp* __restrict__ op,
                                         const R2* pA, const R2* pB, R2* pC, int bA, int bB, int bC) {
    constexpr int TM = 1 << MA, TN = 1 << NB, KK = 1 << KC;
    const int nK = h->nK;
    int xa[TM], xb[TN], kta[KK], ktb[KK];
#pragma unroll
    for (int j = 0; j < TM; ++j) xa[j] = bA ^ h->aT[j];
#pragma unroll
    for (int j = 0; j < TN; ++j) xb[j] = bB ^ h->bT[j];
#pragma unroll
    for (int k = 0; k < KK; ++k) { kta[k] = h->ktA[k]; ktb[k] = h->ktB[k]; }
    R2 acc[TM][TN];
#pragma unroll
    for (int jm = 0; jm < TM; ++jm)
#pragma unroll
        for (int jn = 0; jn < TN; ++jn) { acc[jm][jn].x = 0; acc[jm][jn].y = 0; }
    const int nch = 1 << (nK - KC);
    for (int ch = 0; ch < nch; ++ch) {
        const int kb = ch << KC;
        int ka = h->ktA[kb & 15], kbo = h->ktB[kb & 15];
        if (nK > 4) { ka ^= rseg(op->kA, op->nkA, (unsigned)kb >> 4); kbo ^= rseg(op->kB, op->nkB, (unsigned)kb >> 4); }
        R2 av[TM][KK], bv[TN][KK];
#pragma unroll
        for (int j = 0; j < TM; ++j)
#pragma unroll
            for (int k = 0; k < KK; ++k) av[j][k] = pA[xa[j] ^ ka ^ kta[k]];
#pragma unroll
        for (int j = 0; j < TN; ++j)
#pragma unroll
            for (int k = 0; k < KK; ++k) bv[j][k] = pB[xb[j] ^ kbo ^ ktb[k]];
#pragma unroll
        for (int jm = 0; jm < TM; ++jm)
#pragma unroll
            for (int jn = 0; jn < TN; ++jn)
#pragma unroll
                for (int k = 0; k < KK; ++k) rcmac(acc[jm][jn], av[jm][k], bv[jn][k]);
    }
#pragma unroll
    for (int jm = 0; jm < TM; ++jm)
#pragma unroll
        for (int jn = 0; jn < TN; ++jn) pC[bC ^ h->cT[jm * TN + jn]] = acc[jm][jn];
}

// Ops with < 32 thread-tiles: the spare lane bits split K (interleaved), partial sums meet by xor-shuffles.
template <typename R2>
__device__ __forceinline__ void row_kred(const RowOpHot* __restrict__ h, const RowOp* __restrict__ op, const R2* pA,
                                         const R2* pB, R2* pC, int oA, int oB, int oC, int lane) {
    const int ntt = h->ntt, ks = h->ks, nK = h->nK;
    const int tt = lane & ((1 << ntt) - 1), ksub = lane >> ntt;
    const bool active = ksub < (1 << ks);
    R2 acc; acc.x = 0; acc.y = 0;
    int bC = 0;
    if (active) {
        const int bA = oA ^ rseg(op->tA, op->nsA, (unsigned)tt) ^ h->aT[0];
        const int bB = oB ^ rseg(op->tB, op->nsB, (unsigned)tt) ^ h->bT[0];
        bC = oC ^ rseg(op->tC, op->nsC, (unsigned)tt) ^ h->cT[0];
        const int nl = 1 << (nK - ks);
        for (int kl = 0; kl < nl; ++kl) {
            const int k = ksub | (kl << ks);
            int ka = h->ktA[k & 15], kbo = h->ktB[k & 15];
            if (nK > 4) { ka ^= rseg(op->kA, op->nkA, (unsigned)k >> 4); kbo ^= rseg(op->kB, op->nkB, (unsigned)k >> 4); }
            rcmac(acc, pA[bA ^ ka], pB[bB ^ kbo]);
        }
    }
    for (int i = 0; i < ks; ++i) {
        acc.x += __shfl_xor_sync(0xffffffffu, acc.x, 1 << (ntt + i));
        acc.y += __shfl_xor_sync(0xffffffffu, acc.y, 1 << (ntt + i));
    }
    if (active && ksub == 0) pC[bC] = acc;
}

template <typename R2, int MA, int NB, int KC>
__device__ __forceinline__ void row_tile_pick(const RowOpHot* __restrict__ h, const RowOp* __restrict__ op, R2* arena,
                                              const R2* pA, const R2* pB, R2* pC, int bA, int bB, int bC) {
    if constexpr (TileRegs<R2, MA, NB, KC>::value <= 100) {
        if (h->gen) row_tile<R2, MA, NB, KC, true>(h, op, pA, pB, pC, bA, bB, bC);
        else row_tile<R2, MA, NB, KC, false>(h, op, arena, arena, arena, bA, bB, bC);
    }
}

template <typename R2, int MA, int NB>
__device__ __forceinline__ void row_tile_kc(const RowOpHot* __restrict__ h, const RowOp* __restrict__ op, R2* arena,
                                            const R2* pA, const R2* pB, R2* pC, int bA, int bB, int bC) {
    switch (h->kc) {
    case 0: row_tile_pick<R2, MA, NB, 0>(h, op, arena, pA, pB, pC, bA, bB, bC); break;
    case 1: row_tile_pick<R2, MA, NB, 1>(h, op, arena, pA, pB, pC, bA, bB, bC); break;
    default: row_tile_pick<R2, MA, NB, 2>(h, op, arena, pA, pB, pC, bA, bB, bC); break;
    }
}

}  // namespace

template <typename R2>
__global__ void __launch_bounds__(kRowThreads, 2) rowprog_kernel(const __grid_constant__ RowLaunch P) {
    extern __shared__ __align__(16) unsigned char row_smem[];
    RowOpHot* slots = reinterpret_cast<RowOpHot*>(row_smem);
    R2* arena = reinterpret_cast<R2*>(row_smem + kRowWarps * sizeof(RowOpHot));
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    RowOpHot* h = slots + warp;
    for (long long row = blockIdx.x; row < P.n_rows; row += gridDim.x) {
        // output leaves of this row: one-hot (or +/-) vectors from the bitstring bytes
        // (docs/src/users_guide.md:149-158, docs/src/basics.md:55-63 of the reference)
        if (P.n_leaves) {
            const unsigned char* rb = P.bits + (P.amp0 + row) * (long long)P.n_outputs;
            for (int l = tid; l < P.n_leaves; l += kRowThreads) {
                const RowLeaf lf = P.leaves[l];
                const unsigned char val = rb[lf.out_idx - 1];
                const int len = 1 << lf.span_bits;
                for (int j = 0; j < len; ++j) {
                    R2 v; v.x = 0; v.y = 0;
                    if (val == 0) v.x = (j == 0);
                    else if (val == 1) v.x = (j == 1);
                    else if (val == 2) v.x = (j < 2);
                    else v.x = j == 0 ? 1 : (j == 1 ? -1 : 0);
                    arena[lf.off + j] = v;
                }
            }
        }
        __syncthreads();
        for (int lv = 0; lv < P.n_levels; ++lv) {
            const int u1 = P.level_start[lv + 1];
            for (int u = P.level_start[lv] + warp; u < u1; u += kRowWarps) {
                const RowUnit un = P.units[u];
                const RowOp* __restrict__ op = P.ops + un.op;
                // stage the hot descriptor (128 bytes) in this warp's slot
                __syncwarp();
                reinterpret_cast<unsigned*>(h)[lane] = reinterpret_cast<const unsigned*>(&op->hot)[lane];
                __syncwarp();
                const int oA = op->oA, oB = op->oB, oC = op->oC;
                const R2 *pA = arena, *pB = arena;
                R2* pC = arena;
                if (h->gen) {
                    if (op->gA) pA = reinterpret_cast<const R2*>(op->gA) + (P.amp0 + row) * op->rsA;
                    if (op->gB) pB = reinterpret_cast<const R2*>(op->gB) + (P.amp0 + row) * op->rsB;
                    if (op->gC) pC = reinterpret_cast<R2*>(op->gC) + (P.amp0 + row) * op->rsC;
                }
                if (h->kind == kRowKindKred) {
                    row_kred<R2>(h, op, pA, pB, pC, oA, oB, oC, lane);
                    continue;
                }
                const int tt = un.chunk * 32 + lane;
                if (tt >= (1 << h->ntt)) continue;
                const int bA = oA ^ rseg(op->tA, op->nsA, (unsigned)tt);
                const int bB = oB ^ rseg(op->tB, op->nsB, (unsigned)tt);
                const int bC = oC ^ rseg(op->tC, op->nsC, (unsigned)tt);
                switch (h->ma * 3 + h->nb) {
                case 0: row_tile_kc<R2, 0, 0>(h, op, arena, pA, pB, pC, bA, bB, bC); break;
                case 1: row_tile_kc<R2, 0, 1>(h, op, arena, pA, pB, pC, bA, bB, bC); break;
                case 2: row_tile_kc<R2, 0, 2>(h, op, arena, pA, pB, pC, bA, bB, bC); break;
                case 3: row_tile_kc<R2, 1, 0>(h, op, arena, pA, pB, pC, bA, bB, bC); break;
                case 4: row_tile_kc<R2, 1, 1>(h, op, arena, pA, pB, pC, bA, bB, bC); break;
                case 5: row_tile_kc<R2, 1, 2>(h, op, arena, pA, pB, pC, bA, bB, bC); break;
                case 6: row_tile_kc<R2, 2, 0>(h, op, arena, pA, pB, pC, bA, bB, bC); break;
                case 7: row_tile_kc<R2, 2, 1>(h, op, arena, pA, pB, pC, bA, bB, bC); break;
                default: row_tile_kc<R2, 2, 2>(h, op, arena, pA, pB, pC, bA, bB, bC); break;
                }
            }
            __syncthreads();
        }
        // root: sum over the batched slice bits still open in it, in double (what reduce_root_kernel does)
        if (P.acc && warp == 0) {
            double sx = 0, sy = 0;
            const int len = 1 << P.root_span;
            for (int i = lane; i < len; i += 32) {
                const R2 v = arena[P.root_off + i];
                sx += (double)v.x; sy += (double)v.y;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                sx += __shfl_xor_sync(0xffffffffu, sx, o);
                sy += __shfl_xor_sync(0xffffffffu, sy, o);
            }
            if (lane == 0) {
                P.acc[2 * (P.amp0 + row)] += P.scale * sx;
                P.acc[2 * (P.amp0 + row) + 1] += P.scale * sy;
            }
        }
        __syncthreads();                       // the arena is rewritten by the next row
    }
}

const void* rowprog_func(int dtype) {
    return dtype == 0 ? (const void*)&rowprog_kernel<float2> : (const void*)&rowprog_kernel<double2>;
}

}  // namespace qxb
