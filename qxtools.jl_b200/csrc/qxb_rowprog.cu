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
__device__ __forceinline__ void row_tile(const RowOpHot* __restrict__ h, const RowUnitDesc* __restrict__ op,
                                         const R2* pA, const R2* pB, R2* pC, int bA, int bB, int bC) {
    constexpr int TM = 1 << MA, TN = 1 << NB, KK = 1 << KC;
    const int nK = h->nK;
    int xa[TM], xb[TN], kta[KK], ktb[KK];
#pragma unroll
    for (int j = 0; j < TM; ++j) xa[j] = bA + h->aT[j];
#pragma unroll
    for (int j = 0; j < TN; ++j) xb[j] = bB + h->bT[j];
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
        if (nK > 4) { ka += rseg(op->kA, op->nkA, (unsigned)kb >> 4); kbo += rseg(op->kB, op->nkB, (unsigned)kb >> 4); }
        R2 av[TM][KK], bv[TN][KK];
#pragma unroll
        for (int j = 0; j < TM; ++j)
#pragma unroll
            for (int k = 0; k < KK; ++k) av[j][k] = pA[xa[j] + ka + kta[k]];
#pragma unroll
        for (int j = 0; j < TN; ++j)
#pragma unroll
            for (int k = 0; k < KK; ++k) bv[j][k] = pB[xb[j] + kbo + ktb[k]];
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
        for (int jn = 0; jn < TN; ++jn) pC[bC + h->cT[jm * TN + jn]] = acc[jm][jn];
}

// Ops with < 32 thread-tiles: the spare lane bits split K (interleaved), partial sums meet by xor-shuffles.
template <typename R2>
__device__ __forceinline__ void row_kred(const RowOpHot* __restrict__ h, const RowUnitDesc* __restrict__ op, const R2* pA,
                                         const R2* pB, R2* pC, int bA, int bB, int bC, int lane) {
    const int ntt = h->ntt, ks = h->ks, nK = h->nK;
    const int ksub = lane >> ntt;
    const bool active = bC != kRowNull;
    R2 acc; acc.x = 0; acc.y = 0;
    if (active) {
        const int nl = 1 << (nK - ks);
        auto koffs = [&](int k, int& ka, int& kbo) {
            ka = h->ktA[k & 15]; kbo = h->ktB[k & 15];
            if (nK > 4) { ka += rseg(op->kA, op->nkA, (unsigned)k >> 4); kbo += rseg(op->kB, op->nkB, (unsigned)k >> 4); }
        };
        int kl = 0;
        for (; kl + 4 <= nl; kl += 4) {                      // four k per round: the loads of a round are independent
            R2 a[4], b[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                int ka, kbo;
                koffs(ksub | ((kl + q) << ks), ka, kbo);
                a[q] = pA[bA + ka]; b[q] = pB[bB + kbo];
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) rcmac(acc, a[q], b[q]);
        }
        for (; kl < nl; ++kl) {
            int ka, kbo;
            koffs(ksub | (kl << ks), ka, kbo);
            rcmac(acc, pA[bA + ka], pB[bB + kbo]);
        }
    }
    for (int i = 0; i < ks; ++i) {
        acc.x += __shfl_xor_sync(0xffffffffu, acc.x, 1 << (ntt + i));
        acc.y += __shfl_xor_sync(0xffffffffu, acc.y, 1 << (ntt + i));
    }
    if (active && ksub == 0) pC[bC] = acc;
}

// ComplexF64 nodes with >= 3 M-only and >= 3 N-only bits on the FP64 tensor pipe (DMMA): the warp computes up to four
// 8 x 8 output tiles at once, K in steps of 4.  Per complex k4 step and tile: 2 LDS.128 + 4 MMAs for 256 complex MACs
// (the SIMT 4 x 4 register tile needs 8 LDS.128 + 64 DFMA per lane for 16: twice the shared-memory wavefronts and
// sixteen times the issue slots) -- the fused launches are bound by exactly those two (profiles/r2_summary.md).
__device__ __forceinline__ void dmma(double (&d)[2], const double a, const double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d[0]), "+d"(d[1]) : "d"(a), "d"(b));
}
template <int NT>
__device__ __forceinline__ void row_dmma_tiles(const RowOpHot* __restrict__ h, const RowUnitDesc* __restrict__ op,
                                               const double2* pA, const double2* pB, double2* pC, int bA, int bB, int bC) {
    const int nK = h->nK;
    double cre[NT][2], cim[NT][2];
    int ta[NT], tb[NT];
#pragma unroll
    for (int i = 0; i < NT; ++i) { cre[i][0] = cre[i][1] = cim[i][0] = cim[i][1] = 0.0; ta[i] = bA + h->aT[i]; tb[i] = bB + h->bT[i]; }
    const int nsteps = 1 << (nK - 2);
    for (int s = 0; s < nsteps; ++s) {
        const int k = s << 2;
        int ka = h->ktA[k & 15], kb = h->ktB[k & 15];
        if (nK > 4) { ka += rseg(op->kA, op->nkA, (unsigned)k >> 4); kb += rseg(op->kB, op->nkB, (unsigned)k >> 4); }
        double2 a[NT], b[NT];
#pragma unroll
        for (int i = 0; i < NT; ++i) { a[i] = pA[ta[i] + ka]; b[i] = pB[tb[i] + kb]; }
#pragma unroll
        for (int i = 0; i < NT; ++i) {
            dmma(cre[i], a[i].x, b[i].x);
            dmma(cim[i], a[i].x, b[i].y);
            dmma(cre[i], -a[i].y, b[i].y);
            dmma(cim[i], a[i].y, b[i].x);
        }
    }
    const int n1 = h->cT[4];
#pragma unroll
    for (int i = 0; i < NT; ++i) {
        const int c = bC + h->cT[i];
        pC[c] = make_double2(cre[i][0], cim[i][0]);
        pC[c + n1] = make_double2(cre[i][1], cim[i][1]);
    }
}
template <typename R2>
__device__ __forceinline__ void row_dmma(const RowOpHot* __restrict__ h, const RowUnitDesc* __restrict__ op, const R2* pA,
                                         const R2* pB, R2* pC, int bA, int bB, int bC) {
    if constexpr (sizeof(R2) == 16) {
        switch (h->ma) {
        case 0: row_dmma_tiles<1>(h, op, pA, pB, pC, bA, bB, bC); break;
        case 1: row_dmma_tiles<2>(h, op, pA, pB, pC, bA, bB, bC); break;
        default: row_dmma_tiles<4>(h, op, pA, pB, pC, bA, bB, bC); break;
        }
    }
}

template <typename R2, int MA, int NB, int KC>
__device__ __forceinline__ void row_tile_pick(const RowOpHot* __restrict__ h, const RowUnitDesc* __restrict__ op, R2* arena,
                                              const R2* pA, const R2* pB, R2* pC, int bA, int bB, int bC) {
    if constexpr (TileRegs<R2, MA, NB, KC>::value <= 100) {
        if (h->gen) row_tile<R2, MA, NB, KC, true>(h, op, pA, pB, pC, bA, bB, bC);
        else row_tile<R2, MA, NB, KC, false>(h, op, arena, arena, arena, bA, bB, bC);
    }
}

template <typename R2, int MA, int NB>
__device__ __forceinline__ void row_tile_kc(const RowOpHot* __restrict__ h, const RowUnitDesc* __restrict__ op, R2* arena,
                                            const R2* pA, const R2* pB, R2* pC, int bA, int bB, int bC) {
    switch (h->kc) {
    case 0: row_tile_pick<R2, MA, NB, 0>(h, op, arena, pA, pB, pC, bA, bB, bC); break;
    case 1: row_tile_pick<R2, MA, NB, 1>(h, op, arena, pA, pB, pC, bA, bB, bC); break;
    default: row_tile_pick<R2, MA, NB, 2>(h, op, arena, pA, pB, pC, bA, bB, bC); break;
    }
}

}  // namespace

// 16-byte asynchronous global -> shared copy (LDGSTS): descriptor prefetch without touching registers
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
template <typename R2>
__device__ __forceinline__ void cp_async_elem(R2* smem_dst, const R2* gsrc) {      // one element: 8 or 16 bytes
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    if (sizeof(R2) == 16) asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
    else asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_but_newest() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// Shared memory: [2 descriptor buffers per warp][slot table][arena].  The descriptor of a warp's next unit is in
// flight (cp.async) while it computes the current one; nothing on the unit path reads global memory synchronously.
template <typename R2>
__global__ void __launch_bounds__(kRowThreads, 2) rowprog_kernel(const __grid_constant__ RowLaunch P) {
    extern __shared__ __align__(16) unsigned char row_smem[];
    RowUnitDesc* bufs = reinterpret_cast<RowUnitDesc*>(row_smem);
    uint16_t* slots = reinterpret_cast<uint16_t*>(row_smem + 2 * kRowWarps * sizeof(RowUnitDesc));
    R2* arena = reinterpret_cast<R2*>(row_smem + 2 * kRowWarps * sizeof(RowUnitDesc) + P.slots_bytes);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n_slots = P.level_start[P.n_levels];
    for (int i = tid; i < n_slots; i += kRowThreads) slots[i] = P.slots[i];
    __syncthreads();
    auto prefetch = [&](int s, int par) {           // descriptor of slot s -> buffer par of this warp (26 x 16 bytes)
        if (s >= 0) {
            const unsigned di = slots[s];
            if (di != kRowNull && lane < (int)(sizeof(RowUnitDesc) / 16))
                cp_async16(reinterpret_cast<char*>(bufs + warp * 2 + par) + lane * 16,
                           reinterpret_cast<const char*>(P.descs + di) + lane * 16);
        }
        cp_async_commit();
    };
    long long t_prev = 0;
    int par = 0;
    prefetch(P.level_start[0] + warp, par);
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
        if (P.timing && blockIdx.x == 0 && tid == 0) t_prev = clock64();
        for (int lv = 0; lv < P.n_levels; ++lv) {
            const int s1 = P.level_start[lv + 1];
            for (int s = P.level_start[lv] + warp; s < s1; s += kRowWarps) {
                cp_async_wait_all();                 // the descriptor of slot s has landed in buffer `par`
                __syncwarp();
                const RowUnitDesc* __restrict__ op = bufs + warp * 2 + par;
                const unsigned di = slots[s];
                // next slot of this warp: same level, next level, or the first slot of the next row
                int ns = s + kRowWarps;
                if (ns >= s1) ns = lv + 1 < P.n_levels ? s1 + warp : (row + gridDim.x < P.n_rows ? P.level_start[0] + warp : -1);
                const RowOpHot* __restrict__ h = &op->hot;
                if (di != kRowNull && h->kind == kRowKindCopy) {
                    // staged operand: 2^ntt elements global -> arena; committed BEFORE the descriptor prefetch so that
                    // "all groups but the newest" at the level barrier means "every copy has landed"
                    const int n = 1 << h->ntt;
                    const unsigned dst = (unsigned)op->lC[0] | ((unsigned)op->lC[1] << 16);
                    const R2* src = reinterpret_cast<const R2*>(op->gA) + (op->lsA == kRowShared ? 0ll : (P.amp0 + row) << op->lsA);
                    for (int i = lane; i < n; i += 32) cp_async_elem<R2>(arena + dst + i, src + i);
                    cp_async_commit();
                }
                par ^= 1;
                prefetch(ns, par);
                if (di == kRowNull || h->kind == kRowKindCopy) continue;
                const int bA = op->lA[lane], bB = op->lB[lane], bC = op->lC[lane];
                const R2 *pA = arena, *pB = arena;
                R2* pC = arena;
                if (h->gen) {
                    const long long r = P.amp0 + row;
                    if (op->gA) pA = reinterpret_cast<const R2*>(op->gA) + (op->lsA == kRowShared ? 0ll : r << op->lsA);
                    if (op->gB) pB = reinterpret_cast<const R2*>(op->gB) + (op->lsB == kRowShared ? 0ll : r << op->lsB);
                    if (op->gC) pC = reinterpret_cast<R2*>(op->gC) + (op->lsC == kRowShared ? 0ll : r << op->lsC);
                }
                if (h->kind == kRowKindKred) {
                    row_kred<R2>(h, op, pA, pB, pC, bA, bB, bC, lane);
                    continue;
                }
                if (h->kind == kRowKindDmma) {           // all 32 lanes take part in the MMAs
                    if (h->gen) row_dmma<R2>(h, op, pA, pB, pC, bA, bB, bC);
                    else row_dmma<R2>(h, op, arena, arena, arena, bA, bB, bC);
                    continue;
                }
                if (bC == kRowNull) continue;
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
            cp_async_wait_but_newest();              // staged copies of this level (the descriptor prefetch may stay in flight)
            __syncthreads();
            if (P.timing && blockIdx.x == 0 && tid == 0) {     // per-level cycles of CTA 0 (diagnostics)
                const long long t = clock64();
                P.timing[lv] += t - t_prev;
                t_prev = t;
            }
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
    cp_async_wait_all();
}

// ---------------------------------------------------------------------------------------------------------------
// Ring kernel (see RingLaunch in qxb_rowprog.h).
namespace {
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// bounded wait: a copy that never completes must fail the launch, not hang the GPU
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
    for (unsigned spin = 0; spin < (1u << 28); ++spin) {
        unsigned done;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (done) return;
    }
    __trap();
}
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void* src, unsigned bytes, unsigned bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* dst, unsigned src, unsigned bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
}  // namespace

template <typename R2>
__global__ void __launch_bounds__(kRowThreads, 1) ring_kernel(const __grid_constant__ RingLaunch P) {
    extern __shared__ __align__(128) unsigned char ring_smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool sharedA = P.sUA == 0, sharedB = P.sUB == 0;
    const int S = P.stages;
    RowUnitDesc* descs = reinterpret_cast<RowUnitDesc*>(ring_smem);
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(ring_smem + (size_t)P.n_units * sizeof(RowUnitDesc));
    const size_t head = ((size_t)P.n_units * sizeof(RowUnitDesc) + 8 * kRingMaxStages + 127) / 128 * 128;
    R2* data = reinterpret_cast<R2*>(ring_smem + head);
    // element offsets inside `data`: [A if shared][B if shared][stage 0: A? B? C][stage 1] ...
    const int shA = 0, shB = sharedA ? P.nA : 0;
    const int ring0 = (sharedA ? P.nA : 0) + (sharedB ? P.nB : 0);
    const int stA = 0, stB = sharedA ? 0 : P.nA, stC = stB + (sharedB ? 0 : P.nB);
    const int stage_elems = stC + P.nC;
    const R2* __restrict__ A = reinterpret_cast<const R2*>(P.A);
    const R2* __restrict__ B = reinterpret_cast<const R2*>(P.B);
    R2* __restrict__ C = reinterpret_cast<R2*>(P.C);

    // descriptors (once per CTA) and operands shared by every row
    for (int i = tid; i < P.n_units * (int)(sizeof(RowUnitDesc) / 16); i += kRowThreads)
        reinterpret_cast<uint4*>(descs)[i] = reinterpret_cast<const uint4*>(P.descs)[i];
    if (sharedA) for (int i = tid; i < P.nA; i += kRowThreads) data[shA + i] = A[i];
    if (sharedB) for (int i = tid; i < P.nB; i += kRowThreads) data[shB + i] = B[i];
    if (tid == 0) {
        for (int s = 0; s < S; ++s) mbar_init(smem_u32(bars + s), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const long long n_rows = P.U > (long long)blockIdx.x ? (P.U - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const unsigned row_bytes = (unsigned)(((sharedA ? 0 : P.nA) + (sharedB ? 0 : P.nB)) * (int)sizeof(R2));
    auto issue = [&](long long j) {       // thread 0: operand rows of this CTA's j-th row -> stage j % S
        const int s = (int)(j % S);
        const long long u = blockIdx.x + j * (long long)gridDim.x;
        const unsigned bar = smem_u32(bars + s);
        R2* st = data + ring0 + (long long)s * stage_elems;
        mbar_expect_tx(bar, row_bytes);
        if (!sharedA) bulk_g2s(smem_u32(st + stA), A + u * P.sUA, (unsigned)(P.nA * sizeof(R2)), bar);
        if (!sharedB) bulk_g2s(smem_u32(st + stB), B + u * P.sUB, (unsigned)(P.nB * sizeof(R2)), bar);
    };
    if (tid == 0) for (long long j = 0; j < n_rows && j < S - 1; ++j) issue(j);

    for (long long j = 0; j < n_rows; ++j) {
        const int s = (int)(j % S);
        if (tid == 0 && j + S - 1 < n_rows) {
            // stage (j + S - 1) % S was read by the compute of row j - 1 (all threads passed the barrier that ended it)
            fence_async_smem();
            issue(j + S - 1);
        }
        mbar_wait(smem_u32(bars + s), (unsigned)((j / S) & 1));
        const int base = ring0 + s * stage_elems;
        const int offA = sharedA ? shA : base + stA, offB = sharedB ? shB : base + stB, offC = base + stC;
        for (int u = warp; u < P.n_units; u += kRowWarps) {
            const RowUnitDesc* __restrict__ op = descs + u;
            const RowOpHot* __restrict__ h = &op->hot;
            const int lc = op->lC[lane];
            if (h->kind == kRowKindDmma) {
                row_dmma<R2>(h, op, data, data, data, op->lA[lane] + offA, op->lB[lane] + offB, lc + offC);
                continue;
            }
            if (lc == kRowNull) continue;
            const int bA = op->lA[lane] + offA, bB = op->lB[lane] + offB, bC = lc + offC;
            switch (h->ma * 3 + h->nb) {
            case 0: row_tile_kc<R2, 0, 0>(h, op, data, data, data, data, bA, bB, bC); break;
            case 1: row_tile_kc<R2, 0, 1>(h, op, data, data, data, data, bA, bB, bC); break;
            case 2: row_tile_kc<R2, 0, 2>(h, op, data, data, data, data, bA, bB, bC); break;
            case 3: row_tile_kc<R2, 1, 0>(h, op, data, data, data, data, bA, bB, bC); break;
            case 4: row_tile_kc<R2, 1, 1>(h, op, data, data, data, data, bA, bB, bC); break;
            case 5: row_tile_kc<R2, 1, 2>(h, op, data, data, data, data, bA, bB, bC); break;
            case 6: row_tile_kc<R2, 2, 0>(h, op, data, data, data, data, bA, bB, bC); break;
            case 7: row_tile_kc<R2, 2, 1>(h, op, data, data, data, data, bA, bB, bC); break;
            default: row_tile_kc<R2, 2, 2>(h, op, data, data, data, data, bA, bB, bC); break;
            }
        }
        fence_async_smem();                    // this thread's writes to the C stage -> visible to the bulk store
        if (tid == 0) {
            // the C region of the NEXT row's stage was last read by the bulk store of row j + 1 - S: it must be done
            // before anyone passes the barrier below and starts writing that region
            if (S == 2) bulk_wait_read<0>(); else if (S == 3) bulk_wait_read<1>(); else bulk_wait_read<2>();
        }
        __syncthreads();
        if (tid == 0) {
            const long long u = blockIdx.x + j * (long long)gridDim.x;
            bulk_s2g(C + u * P.sUC, smem_u32(data + offC), (unsigned)(P.nC * sizeof(R2)));
            bulk_commit();
        }
    }
    if (tid == 0) bulk_wait_all();             // shared memory must outlive the last stores
}

const void* ring_func(int dtype) {
    return dtype == 0 ? (const void*)&ring_kernel<float2> : (const void*)&ring_kernel<double2>;
}

const void* rowprog_func(int dtype) {
    return dtype == 0 ? (const void*)&rowprog_kernel<float2> : (const void*)&rowprog_kernel<double2>;
}

}  // namespace qxb
