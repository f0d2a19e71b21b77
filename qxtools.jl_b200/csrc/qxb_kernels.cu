// qxb200 -- device kernels (sm_100a).  See qxb_kernels.cuh for the data layout.
#include "qxb_kernels.cuh"

namespace qxb {

template <typename R2> struct Real;
template <> struct Real<float2> { using T = float; };
template <> struct Real<double2> { using T = double; };

template <typename R2>
__device__ __forceinline__ void cmac(R2& acc, const R2 a, const R2 b) {
    acc.x = fma(a.x, b.x, acc.x);
    acc.x = fma(-a.y, b.y, acc.x);
    acc.y = fma(a.x, b.y, acc.y);
    acc.y = fma(a.y, b.x, acc.y);
}

__device__ __forceinline__ long long segeval(const DSeg* s, int n, unsigned long long x) {
    long long r = 0;
    for (int i = 0; i < n; ++i) {
        const DSeg g = s[i];
        r |= (long long)(((x >> g.src) & ((1ull << g.len) - 1ull)) << g.dst);
    }
    return r;
}

// Generic batched bit-segment contraction with a register tile.
//
// C's bits are split three ways: the low `lob` bits are covered by the threads of
// a block (coalesced stores), `ma` M-only bits and `nb` N-only bits form a
// 2^ma x 2^nb register tile per thread (so 2^ma*K loads of A and 2^nb*K loads of
// B feed 2^(ma+nb) outputs -- the nodes that broadcast small operands into a big
// C are otherwise L1/L2-read-bound), and the remaining "hi" bits are enumerated
// by the tile index through segment maps.  K is walked in chunks of 2^KC: all
// (2^ma + 2^nb) * 2^KC loads of a chunk are issued before its FMAs (the kernel
// lives on memory-level parallelism), accumulators stay in registers.
// One register tile: Ap/Bp/Cp already include every address part except the tile / k offsets.
// SM = operands live in shared memory (plain loads) instead of global memory (__ldg).
template <typename R2, int KC, int MA, int NB, bool ONE, bool SM>
__device__ __forceinline__ void tile_compute(const R2* __restrict__ Ap, const R2* __restrict__ Bp,
                                             R2* __restrict__ Cp, const OpParams& p) {
    constexpr int TM = 1 << MA, TN = 1 << NB, KK = 1 << KC;
    auto ld = [](const R2* q) -> R2 { return SM ? *q : __ldg(q); };
    if (ONE) {
        // K fits one register chunk: load everything, then each output is a short
        // dot product that is stored at once (no accumulator array kept live)
        R2 av[TM][KK], bv[TN][KK];
#pragma unroll
        for (int j = 0; j < TM; ++j)
#pragma unroll
            for (int k = 0; k < KK; ++k) av[j][k] = ld(Ap + p.aT[j] + p.ktabA[k]);
#pragma unroll
        for (int j = 0; j < TN; ++j)
#pragma unroll
            for (int k = 0; k < KK; ++k) bv[j][k] = ld(Bp + p.bT[j] + p.ktabB[k]);
#pragma unroll
        for (int jm = 0; jm < TM; ++jm) {
#pragma unroll
            for (int jn = 0; jn < TN; ++jn) {
                R2 acc; acc.x = 0; acc.y = 0;
#pragma unroll
                for (int k = 0; k < KK; ++k) cmac(acc, av[jm][k], bv[jn][k]);
                Cp[p.cT[jm * TN + jn]] = acc;
            }
        }
    } else {
        const int nchunks = 1 << (p.nK - KC);
        R2 acc[TM][TN];
#pragma unroll
        for (int jm = 0; jm < TM; ++jm)
#pragma unroll
            for (int jn = 0; jn < TN; ++jn) { acc[jm][jn].x = 0; acc[jm][jn].y = 0; }
        for (int ch = 0; ch < nchunks; ++ch) {
            const unsigned long long kb = (unsigned long long)ch << KC;
            const R2* Ac = Ap + segeval(p.kA, p.nkA, kb);
            const R2* Bc = Bp + segeval(p.kB, p.nkB, kb);
            R2 av[TM][KK], bv[TN][KK];
#pragma unroll
            for (int j = 0; j < TM; ++j)
#pragma unroll
                for (int k = 0; k < KK; ++k) av[j][k] = ld(Ac + p.aT[j] + p.ktabA[k]);
#pragma unroll
            for (int j = 0; j < TN; ++j)
#pragma unroll
                for (int k = 0; k < KK; ++k) bv[j][k] = ld(Bc + p.bT[j] + p.ktabB[k]);
#pragma unroll
            for (int jm = 0; jm < TM; ++jm)
#pragma unroll
                for (int jn = 0; jn < TN; ++jn)
#pragma unroll
                    for (int k = 0; k < KK; ++k) cmac(acc[jm][jn], av[jm][k], bv[jn][k]);
        }
#pragma unroll
        for (int jm = 0; jm < TM; ++jm)
#pragma unroll
            for (int jn = 0; jn < TN; ++jn) Cp[p.cT[jm * TN + jn]] = acc[jm][jn];
    }
}

// MINB = resident CTAs per SM the register allocation is bounded for (2: <= 128 registers, 3: <= 85)
template <typename R2, int KC, int MA, int NB, bool ONE, int MINB>
__global__ void __launch_bounds__(kThreads, MINB)
contract_kernel(const __grid_constant__ OpParams p) {
    const R2* __restrict__ A = reinterpret_cast<const R2*>(p.A);
    const R2* __restrict__ B = reinterpret_cast<const R2*>(p.B);
    R2* __restrict__ C = reinterpret_cast<R2*>(p.C);
    const int tid = threadIdx.x;
    const int lob = p.lob;
    const int sub_bits = 8 - lob;
    const int sub = tid >> lob;
    const unsigned lo = tid & ((1u << lob) - 1u);
    const long long aLo = segeval(p.sAlo, p.nsAlo, lo);
    const long long bLo = segeval(p.sBlo, p.nsBlo, lo);
    const long long cLo = segeval(p.sClo, p.nsClo, lo);
    const int hb = p.hb;
    const long long hmask = (1ll << hb) - 1ll;
    for (long long t0 = ((long long)blockIdx.x << sub_bits); t0 < p.tiles;
         t0 += ((long long)gridDim.x << sub_bits)) {
        const long long tile = t0 + sub;
        if (tile >= p.tiles) continue;
        const long long u = tile >> hb;
        const unsigned long long hh = (unsigned long long)(tile & hmask);
        const R2* Ap = A + u * p.sUA + segeval(p.sAhi, p.nsAhi, hh) + aLo;
        const R2* Bp = B + u * p.sUB + segeval(p.sBhi, p.nsBhi, hh) + bLo;
        R2* Cp = C + u * p.sUC + segeval(p.sChi, p.nsChi, hh) + cLo;
        tile_compute<R2, KC, MA, NB, ONE, false>(Ap, Bp, Cp, p);
    }
}

// Broadcast-type nodes (two small operands, large C -- e.g. the outer-product-like steps of
// a re-planned tree): one CTA owns one bitstring row at a time, stages A[u] and B[u]
// (2^aBits + 2^bBits elements) in shared memory once, then produces all 2^nC outputs of that
// row from shared memory.  Needs lob == 8 (nC - ma - nb >= 8).
template <typename R2, int KC, int MA, int NB, bool ONE>
__global__ void __launch_bounds__(kThreads, 2)
contract_smem_kernel(const __grid_constant__ OpParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    R2* sA = reinterpret_cast<R2*>(smem_raw);
    R2* sB = sA + (1ll << p.aBits);
    const R2* __restrict__ A = reinterpret_cast<const R2*>(p.A);
    const R2* __restrict__ B = reinterpret_cast<const R2*>(p.B);
    R2* __restrict__ C = reinterpret_cast<R2*>(p.C);
    const int tid = threadIdx.x;
    const unsigned lo = tid;
    const long long aLo = segeval(p.sAlo, p.nsAlo, lo);
    const long long bLo = segeval(p.sBlo, p.nsBlo, lo);
    const long long cLo = segeval(p.sClo, p.nsClo, lo);
    const long long nA = 1ll << p.aBits, nB = 1ll << p.bBits;
    const long long tiles_u = 1ll << p.hb;
    // an operand shared by every bitstring row (stride 0) is staged once per CTA, not once per row
    const bool sharedA = p.sUA == 0, sharedB = p.sUB == 0;
    if (sharedA) for (long long i = tid; i < nA; i += kThreads) sA[i] = __ldg(A + i);
    if (sharedB) for (long long i = tid; i < nB; i += kThreads) sB[i] = __ldg(B + i);
    for (long long u = blockIdx.x; u < p.U; u += gridDim.x) {
        __syncthreads();                       // previous row fully consumed
        const R2* Au = A + u * p.sUA;
        const R2* Bu = B + u * p.sUB;
        if (!sharedA) for (long long i = tid; i < nA; i += kThreads) sA[i] = __ldg(Au + i);
        if (!sharedB) for (long long i = tid; i < nB; i += kThreads) sB[i] = __ldg(Bu + i);
        __syncthreads();
        R2* Cu = C + u * p.sUC + cLo;
        for (long long hh = 0; hh < tiles_u; ++hh) {
            const R2* Ap = sA + segeval(p.sAhi, p.nsAhi, (unsigned long long)hh) + aLo;
            const R2* Bp = sB + segeval(p.sBhi, p.nsBhi, (unsigned long long)hh) + bLo;
            R2* Cp = Cu + segeval(p.sChi, p.nsChi, (unsigned long long)hh);
            tile_compute<R2, KC, MA, NB, ONE, true>(Ap, Bp, Cp, p);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Same node shape as contract_smem_kernel, but the operand rows arrive through the TMA engine:
// one thread issues 1-D bulk copies (cp.async.bulk ... mbarrier::complete_tx) of A[u] and B[u]
// into a ring of shared-memory stages, several bitstring rows ahead of the row being computed.
// Why: the K = 8 - 32 nodes of the tree-searched plans run at 25 % occupancy with ~20 loads in
// flight per thread -- about 5 - 10 KB in flight per SM where HBM needs ~35 KB (6.5 TB/s / 148 SMs
// x ~800 ns); ncu: "no eligible warp" half of the cycles (profiles/r1p_summary.md).  Bulk copies
// keep (stages - 1) x (|A[u]| + |B[u]|) bytes in flight per CTA without holding a register.
// EXPERIMENT (QXB_SMEM_TMA=1): written without GPU access, never run; off by default.
__device__ __forceinline__ unsigned tma_smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tma_mbar_wait(unsigned bar, unsigned parity) {
    for (unsigned spin = 0; spin < (1u << 27); ++spin) {
        unsigned done;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (done) return;
    }
    __trap();                                  // a copy that never completes must fail the launch, not hang the GPU
}

constexpr int kTmaMaxStages = 4;

template <typename R2, int KC, int MA, int NB, bool ONE>
__global__ void __launch_bounds__(kThreads, 1)
contract_tma_kernel(const __grid_constant__ OpParams p, const int stages) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const long long nA = 1ll << p.aBits, nB = 1ll << p.bBits;
    const bool sharedA = p.sUA == 0, sharedB = p.sUB == 0;
    // layout: [stage 0: A row | B row][stage 1] ... [mbarriers]; a shared operand lives once, in stage 0's slot
    const long long stage_elems = nA + nB;
    R2* ring = reinterpret_cast<R2*>(smem_raw);
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(ring + stage_elems * stages);
    const R2* __restrict__ A = reinterpret_cast<const R2*>(p.A);
    const R2* __restrict__ B = reinterpret_cast<const R2*>(p.B);
    R2* __restrict__ C = reinterpret_cast<R2*>(p.C);
    const int tid = threadIdx.x;
    const unsigned lo = tid;
    const long long aLo = segeval(p.sAlo, p.nsAlo, lo);
    const long long bLo = segeval(p.sBlo, p.nsBlo, lo);
    const long long cLo = segeval(p.sClo, p.nsClo, lo);
    const long long tiles_u = 1ll << p.hb;
    const unsigned row_bytes = (unsigned)(((sharedA ? 0 : nA) + (sharedB ? 0 : nB)) * (long long)sizeof(R2));

    if (tid == 0) {
        for (int s = 0; s < stages; ++s)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(tma_smem_u32(bars + s)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // operands shared by every bitstring row: plain loads, once per CTA
    R2* shA = ring;                       // stage 0's A slot when A is shared
    R2* shB = ring + nA;
    if (sharedA) for (long long i = tid; i < nA; i += kThreads) shA[i] = __ldg(A + i);
    if (sharedB) for (long long i = tid; i < nB; i += kThreads) shB[i] = __ldg(B + i);
    __syncthreads();

    // rows of this CTA: u_j = blockIdx.x + j * gridDim.x, j = 0 .. n_rows - 1; row j uses stage j % stages
    const long long n_rows = p.U > (long long)blockIdx.x ? (p.U - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    auto issue = [&](long long j) {       // thread 0 only
        const int s = (int)(j % stages);
        const long long u = blockIdx.x + j * (long long)gridDim.x;
        const unsigned bar = tma_smem_u32(bars + s);
        R2* dstA = ring + (long long)s * stage_elems;
        R2* dstB = dstA + nA;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%1], %0;" ::"r"(row_bytes), "r"(bar) : "memory");
        if (!sharedA)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(tma_smem_u32(dstA)), "l"(A + u * p.sUA), "r"((unsigned)(nA * sizeof(R2))), "r"(bar) : "memory");
        if (!sharedB)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(tma_smem_u32(dstB)), "l"(B + u * p.sUB), "r"((unsigned)(nB * sizeof(R2))), "r"(bar) : "memory");
    };
    if (tid == 0)
        for (long long j = 0; j < n_rows && j < stages; ++j) issue(j);

    for (long long j = 0; j < n_rows; ++j) {
        const int s = (int)(j % stages);
        const long long u = blockIdx.x + j * (long long)gridDim.x;
        tma_mbar_wait(tma_smem_u32(bars + s), (unsigned)((j / stages) & 1));
        const R2* rowA = sharedA ? shA : ring + (long long)s * stage_elems;
        const R2* rowB = sharedB ? shB : ring + (long long)s * stage_elems + nA;
        R2* Cu = C + u * p.sUC + cLo;
        for (long long hh = 0; hh < tiles_u; ++hh) {
            const R2* Ap = rowA + segeval(p.sAhi, p.nsAhi, (unsigned long long)hh) + aLo;
            const R2* Bp = rowB + segeval(p.sBhi, p.nsBhi, (unsigned long long)hh) + bLo;
            R2* Cp = Cu + segeval(p.sChi, p.nsChi, (unsigned long long)hh);
            tile_compute<R2, KC, MA, NB, ONE, true>(Ap, Bp, Cp, p);
        }
        __syncthreads();                   // every thread is done reading stage s
        if (tid == 0 && j + stages < n_rows) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy reads before the async-proxy overwrite
            issue(j + stages);
        }
    }
}

template <typename R2, int KC, int MA, int NB>
static const void* pick_tma_one(bool one) {
    return one ? (const void*)&contract_tma_kernel<R2, KC, MA, NB, true>
               : (const void*)&contract_tma_kernel<R2, KC, MA, NB, false>;
}
template <typename R2, int KC>
static const void* pick_tma_tile(int ma, int nb, bool one) {
    if (ma == 1 && nb == 1) return pick_tma_one<R2, KC, 1, 1>(one);
    if (ma == 1 && nb == 2) return pick_tma_one<R2, KC, 1, 2>(one);
    if (ma == 2 && nb == 1) return pick_tma_one<R2, KC, 2, 1>(one);
    if (ma == 2 && nb == 2) return pick_tma_one<R2, KC, 2, 2>(one);
    if (ma == 0 && nb == 1) return pick_tma_one<R2, KC, 0, 1>(one);
    if (ma == 0 && nb == 2) return pick_tma_one<R2, KC, 0, 2>(one);
    if (ma == 1 && nb == 0) return pick_tma_one<R2, KC, 1, 0>(one);
    if (ma == 2 && nb == 0) return pick_tma_one<R2, KC, 2, 0>(one);
    return nullptr;
}
template <typename R2>
static const void* pick_tma_kc(int kc, int ma, int nb, bool one) {
    switch (kc) {
    case 1: return pick_tma_tile<R2, 1>(ma, nb, one);
    case 2: return pick_tma_tile<R2, 2>(ma, nb, one);
    case 3: return pick_tma_tile<R2, 3>(ma, nb, one);
    default: return nullptr;
    }
}
// nullptr when there is no TMA-staged variant for this shape; second kernel argument = number of stages (2..4);
// dynamic shared memory = stages * (2^aBits + 2^bBits) * sizeof(element) + 64
const void* contract_tma_func(int dtype, int kc, int ma, int nb, bool single_chunk) {
    return dtype == 0 ? pick_tma_kc<float2>(kc, ma, nb, single_chunk) : pick_tma_kc<double2>(kc, ma, nb, single_chunk);
}

template <typename R2, int KC, int MA, int NB>
static const void* pick_smem_one(bool one) {
    return one ? (const void*)&contract_smem_kernel<R2, KC, MA, NB, true>
               : (const void*)&contract_smem_kernel<R2, KC, MA, NB, false>;
}
template <typename R2, int KC>
static const void* pick_smem_tile(int ma, int nb, bool one) {
    if (ma == 1 && nb == 1) return pick_smem_one<R2, KC, 1, 1>(one);
    if (ma == 1 && nb == 2) return pick_smem_one<R2, KC, 1, 2>(one);
    if (ma == 2 && nb == 1) return pick_smem_one<R2, KC, 2, 1>(one);
    if (ma == 2 && nb == 2) return pick_smem_one<R2, KC, 2, 2>(one);
    if (ma == 0 && nb == 1) return pick_smem_one<R2, KC, 0, 1>(one);
    if (ma == 0 && nb == 2) return pick_smem_one<R2, KC, 0, 2>(one);
    if (ma == 1 && nb == 0) return pick_smem_one<R2, KC, 1, 0>(one);
    if (ma == 2 && nb == 0) return pick_smem_one<R2, KC, 2, 0>(one);
    return nullptr;
}
template <typename R2>
static const void* pick_smem_kc(int kc, int ma, int nb, bool one) {
    switch (kc) {
    case 0: return pick_smem_tile<R2, 0>(ma, nb, one);
    case 1: return pick_smem_tile<R2, 1>(ma, nb, one);
    case 2: return pick_smem_tile<R2, 2>(ma, nb, one);
    case 3: return pick_smem_tile<R2, 3>(ma, nb, one);
    default: return nullptr;
    }
}

// nullptr when there is no shared-memory variant for this shape
const void* contract_smem_func(int dtype, int kc, int ma, int nb, bool single_chunk) {
    return dtype == 0 ? pick_smem_kc<float2>(kc, ma, nb, single_chunk) : pick_smem_kc<double2>(kc, ma, nb, single_chunk);
}

template <typename R2, int KC, int MA, int NB>
static const void* pick_one(bool one, int minb) {
    if (minb >= 3)
        return one ? (const void*)&contract_kernel<R2, KC, MA, NB, true, 3> : (const void*)&contract_kernel<R2, KC, MA, NB, false, 3>;
    return one ? (const void*)&contract_kernel<R2, KC, MA, NB, true, 2> : (const void*)&contract_kernel<R2, KC, MA, NB, false, 2>;
}
template <typename R2, int KC, int MA>
static const void* pick_nb(int nb, bool one, int minb) {
    switch (nb) {
    case 0: return pick_one<R2, KC, MA, 0>(one, minb);
    case 1: return pick_one<R2, KC, MA, 1>(one, minb);
    default: return pick_one<R2, KC, MA, 2>(one, minb);
    }
}
template <typename R2, int KC>
static const void* pick_ma(int ma, int nb, bool one, int minb) {
    switch (ma) {
    case 0: return pick_nb<R2, KC, 0>(nb, one, minb);
    case 1: return pick_nb<R2, KC, 1>(nb, one, minb);
    default: return pick_nb<R2, KC, 2>(nb, one, minb);
    }
}
template <typename R2>
static const void* pick_kc(int kc, int ma, int nb, bool one, int minb) {
    switch (kc) {
    case 0: return pick_ma<R2, 0>(ma, nb, one, minb);
    case 1: return pick_ma<R2, 1>(ma, nb, one, minb);
    case 2: return pick_ma<R2, 2>(ma, nb, one, minb);
    default: return pick_ma<R2, 3>(ma, nb, one, minb);
    }
}

// min_blocks = 3 returns the variant compiled for three resident CTAs per SM, or the two-CTA one when that
// variant spills (local memory > 0): more warps only pay when the registers really fit
const void* contract_func(int dtype, int kc, int ma, int nb, bool single_chunk, int min_blocks) {
    const void* f2 = dtype == 0 ? pick_kc<float2>(kc, ma, nb, single_chunk, 2) : pick_kc<double2>(kc, ma, nb, single_chunk, 2);
    if (min_blocks < 3) return f2;
    const void* f3 = dtype == 0 ? pick_kc<float2>(kc, ma, nb, single_chunk, 3) : pick_kc<double2>(kc, ma, nb, single_chunk, 3);
    cudaFuncAttributes at{};
    if (cudaFuncGetAttributes(&at, f3) != cudaSuccess || at.localSizeBytes > 0) { cudaGetLastError(); return f2; }
    return f3;
}

// ---------------------------------------------------------------------------------------------
// Tiled complex GEMM (SIMT FMA).  The compute-bound regime: large K, many M-only and N-only
// bits (Sycamore-like networks: K = 2^9, M x N = 2^9 x 2^11 per bitstring, ~60 flop/B).
// B200 has no tcgen05 kind for FP64, and a 3xTF32 split needs ~12 real MMAs per complex FP32
// product to keep fp32 accuracy, which caps it near the FFMA peak: so this is a classic
// shared-memory-tiled FMA GEMM -- 2^TMB x 2^TNB x 2^KCB block tile, 4x4 complex micro-tile per
// thread, register-prefetched operand tiles gathered through the bit maps (consecutive threads
// read ascending global addresses), strided micro-tile so shared-memory reads are
// conflict-free.
template <typename R2, int TMB, int TNB, int KCB>
__global__ void __launch_bounds__(1 << (TMB + TNB - 4))
gemm_kernel(const __grid_constant__ GemmParams p) {
    constexpr int BM = 1 << TMB, BN = 1 << TNB, BK = 1 << KCB;
    constexpr int NT = 1 << (TMB + TNB - 4);
    constexpr int LA = (BM * BK) / NT, LB = (BN * BK) / NT;
    constexpr int TX = BN / 4, TY = BM / 4;          // thread grid; micro-tile rows ty + i*TY, cols tx + j*TX
    __shared__ R2 sA[BK * BM];
    __shared__ R2 sB[BK * BN];
    const R2* __restrict__ A = reinterpret_cast<const R2*>(p.A);
    const R2* __restrict__ B = reinterpret_cast<const R2*>(p.B);
    R2* __restrict__ C = reinterpret_cast<R2*>(p.C);
    const int tid = threadIdx.x;
    const int tx = tid % TX, ty = tid / TX;
    // per-thread load slots
    int aOff[LA], bOff[LB];          // 32-bit: the host only picks this kernel for spans <= 2^30
    int aSm[LA], bSm[LB];
#pragma unroll
    for (int i = 0; i < LA; ++i) {
        const int e = tid + i * NT;
        int o = 0, sm = 0;
#pragma unroll
        for (int j = 0; j < TMB + KCB; ++j) if ((e >> j) & 1) { o += (int)p.aLoadOff[j]; sm += p.aLoadSm[j]; }
        aOff[i] = o; aSm[i] = sm;
    }
#pragma unroll
    for (int i = 0; i < LB; ++i) {
        const int e = tid + i * NT;
        int o = 0, sm = 0;
#pragma unroll
        for (int j = 0; j < TNB + KCB; ++j) if ((e >> j) & 1) { o += (int)p.bLoadOff[j]; sm += p.bLoadSm[j]; }
        bOff[i] = o; bSm[i] = sm;
    }
    int cMo[4], cNo[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = ty + i * TY, n = tx + i * TX;
        int om = 0, on = 0;
#pragma unroll
        for (int t = 0; t < TMB; ++t) if ((m >> t) & 1) om += (int)p.cM[t];
#pragma unroll
        for (int t = 0; t < TNB; ++t) if ((n >> t) & 1) on += (int)p.cN[t];
        cMo[i] = om; cNo[i] = on;
    }
    const long long hmask = (1ll << p.hb) - 1ll;
    const int nchunks = 1 << (p.nK - KCB);
    for (long long tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
        const long long u = tile >> p.hb;
        const unsigned long long hh = (unsigned long long)(tile & hmask);
        const R2* Ap = A + u * p.sUA + segeval(p.sAhi, p.nsAhi, hh);
        const R2* Bp = B + u * p.sUB + segeval(p.sBhi, p.nsBhi, hh);
        R2* Cp = C + u * p.sUC + segeval(p.sChi, p.nsChi, hh);
        R2 acc[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) { acc[i][j].x = 0; acc[i][j].y = 0; }
        R2 ra[LA], rb[LB];
#pragma unroll
        for (int i = 0; i < LA; ++i) ra[i] = __ldg(Ap + aOff[i]);
#pragma unroll
        for (int i = 0; i < LB; ++i) rb[i] = __ldg(Bp + bOff[i]);
        for (int ch = 0; ch < nchunks; ++ch) {
            __syncthreads();                     // previous chunk fully consumed
#pragma unroll
            for (int i = 0; i < LA; ++i) sA[aSm[i]] = ra[i];
#pragma unroll
            for (int i = 0; i < LB; ++i) sB[bSm[i]] = rb[i];
            __syncthreads();
            if (ch + 1 < nchunks) {              // prefetch the next chunk while computing this one
                const unsigned long long kb = (unsigned long long)(ch + 1) << KCB;
                const R2* An = Ap + segeval(p.kA, p.nkA, kb);
                const R2* Bn = Bp + segeval(p.kB, p.nkB, kb);
#pragma unroll
                for (int i = 0; i < LA; ++i) ra[i] = __ldg(An + aOff[i]);
#pragma unroll
                for (int i = 0; i < LB; ++i) rb[i] = __ldg(Bn + bOff[i]);
            }
#pragma unroll
            for (int kk = 0; kk < BK; ++kk) {
                R2 a[4], b[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) a[i] = sA[kk * BM + ty + i * TY];
#pragma unroll
                for (int j = 0; j < 4; ++j) b[j] = sB[kk * BN + tx + j * TX];
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) cmac(acc[i][j], a[i], b[j]);
            }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) Cp[cMo[i] + cNo[j]] = acc[i][j];
    }
}

int gemm_kcb(int dtype) { return dtype == 0 ? 4 : 3; }

template <typename R2, int KCB>
static const void* pick_gemm(int tmb, int tnb) {
    if (tmb == 6 && tnb == 6) return (const void*)&gemm_kernel<R2, 6, 6, KCB>;
    if (tmb == 5 && tnb == 6) return (const void*)&gemm_kernel<R2, 5, 6, KCB>;
    if (tmb == 6 && tnb == 5) return (const void*)&gemm_kernel<R2, 6, 5, KCB>;
    if (tmb == 5 && tnb == 5) return (const void*)&gemm_kernel<R2, 5, 5, KCB>;
    return nullptr;
}
const void* gemm_func(int dtype, int tmb, int tnb) {
    return dtype == 0 ? pick_gemm<float2, 4>(tmb, tnb) : pick_gemm<double2, 3>(tmb, tnb);
}

// ---------------------------------------------------------------------------------------------
// Tensor-core complex GEMMs (warp-level mma.sync), same GemmParams / tile enumeration as
// gemm_kernel.  Block tile 64 x 64 complex, K chunk 16 (c32) / 8 (c64).
//
//  ComplexF64 (gemm_dmma_kernel): DMMA mma.sync.m8n8k4.f64 -- 4 real MMAs per complex tile
//      product (re += Ar*Br, re += (-Ai)*Bi, im += Ar*Bi, im += Ai*Br); warp tile 16 x 32, 8 warps.
//      Shared tiles are [k][m] double2 with a row stride of 64 + 2 elements: the fragment loads
//      of a quarter warp (lanes g = 0..1, t = 0..3 -> element t * 66 + g) hit 8 distinct
//      16-byte bank groups.
//  ComplexF32 (gemm_tf32x3_kernel): 3xTF32 mma.sync.m16n8k8 -- every fp32 operand is split
//      once, while it is staged into shared memory, into a tf32 head and a tf32 tail
//      (x = hi + lo, |lo| <= 2^-11 |x|); each of the 4 real products is hi*hi + lo*hi + hi*lo
//      accumulated in fp32 (the dropped lo*lo term is ~2^-22 relative, i.e. fp32 rounding level):
//      12 MMAs per complex tile product against 4 FFMA per complex MAC, so per complex MAC the
//      tensor pipe has (tf32 MAC rate / 12) against (FFMA rate / 4).  Warp tile 32 x 32, 4 warps.
//      Shared tiles are stored FRAGMENT-MAJOR: for every (k8 step, m16 tile, plane) the 128 floats
//      a warp needs are laid out [lane][a0 a1 a2 a3], so one conflict-free LDS.128 per lane yields
//      the four consecutive registers of an MMA A operand (planes: re_hi, im_hi, re_lo, im_lo);
//      B likewise as [lane][re b0 b1, im b0 b1] for the hi and the lo halves.
__device__ __forceinline__ void mma_tf32(float (&d)[4], const float4 a, const float b0, const float b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(__float_as_uint(a.x)), "r"(__float_as_uint(a.y)), "r"(__float_as_uint(a.z)),
                   "r"(__float_as_uint(a.w)), "r"(__float_as_uint(b0)), "r"(__float_as_uint(b1)));
}
__device__ __forceinline__ void mma_tf32_z(float (&d)[4], const float4 a, const float b0, const float b1) {   // C = 0
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%10,%10,%10};\n"
                 : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3])
                 : "r"(__float_as_uint(a.x)), "r"(__float_as_uint(a.y)), "r"(__float_as_uint(a.z)),
                   "r"(__float_as_uint(a.w)), "r"(__float_as_uint(b0)), "r"(__float_as_uint(b1)), "f"(0.f));
}
__device__ __forceinline__ void mma_f64(double (&d)[2], const double a, const double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d[0]), "+d"(d[1]) : "d"(a), "d"(b));
}
// round-to-nearest (ties away) to the 10-bit tf32 mantissa with two integer ops (what cvt.rna.tf32.f32 does,
// which compiles to ~5 instructions here); finite inputs only
__device__ __forceinline__ float to_tf32(float x) {
    return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u);
}
__device__ __forceinline__ float fneg_bits(float x) { return __uint_as_float(__float_as_uint(x) ^ 0x80000000u); }

template <int NBITS>
__device__ __forceinline__ int bit_offset(int idx, const long long* tbl) {
    int o = 0;
#pragma unroll
    for (int t = 0; t < NBITS; ++t) if ((idx >> t) & 1) o += (int)tbl[t];
    return o;
}

// Load-slot bookkeeping shared by both kernels: slot e = tid + i * NT (i < L); bit j of e selects a
// global-offset contribution off[j] and a shared-memory contribution sm(j).  The thread part (bits
// below LOGNT) is summed into registers once; the slot part is uniform and re-derived from the
// parameter block where it is used.
template <int LOGNT, int NB_, typename F>
__device__ __forceinline__ int slot_sum(int i, F f) {
    int o = 0;
#pragma unroll
    for (int j = LOGNT; j < NB_; ++j) if ((i >> (j - LOGNT)) & 1) o += f(j);
    return o;
}

// bank swizzle of the fragment-major tiles: within a 128-float block (32 lane slots of 16 bytes) lane slot
// bits 0..1 are XORed with bits 3..4, so operand stores whose lanes run over m (n) land in 8 distinct
// 16-byte bank groups (8- and 16-way conflicts otherwise) while a quarter warp's LDS.128 stays conflict-free
__device__ __forceinline__ int frag_swz(int off) { return off ^ (((off >> 5) & 3) << 2); }

template <int TMB, int TNB>
__global__ void __launch_bounds__(((1 << (TMB + TNB)) / 1024) * 32, 2)
gemm_tf32x3_kernel(const __grid_constant__ GemmParams p) {
    using R2 = float2;
    constexpr int KCB = 4;
    constexpr int BM = 1 << TMB, BN = 1 << TNB, BK = 1 << KCB;
    constexpr int WGM = BM / 32, WGN = BN / 32;
    constexpr int NT = 32 * WGM * WGN;
    constexpr int LOGNT = TMB + TNB - 5;
    constexpr int LA = (BM * BK) / NT, LB = (BN * BK) / NT;
    constexpr int MT = 2, NTL = 4;                    // m16 / n8 MMA tiles per 32 x 32 warp tile
    constexpr int MTC = BM / 16, NTC = BN / 8;        // MMA tiles per block tile
    constexpr int SA = (BK / 8) * MTC * 512, SB = (BK / 8) * NTC * 256;   // floats per stage
    static_assert(NT == (1 << LOGNT), "thread count");
    extern __shared__ __align__(16) float smem_f[];   // two stages of [A tile | B tile]
    const R2* __restrict__ A = reinterpret_cast<const R2*>(p.A);
    const R2* __restrict__ B = reinterpret_cast<const R2*>(p.B);
    R2* __restrict__ C = reinterpret_cast<R2*>(p.C);
    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int wm0 = (warp % WGM) * 32, wn0 = (warp / WGM) * 32;
    int aOffT = 0, aSmT = 0, bOffT = 0, bSmT = 0;
#pragma unroll
    for (int j = 0; j < LOGNT; ++j) {
        if ((tid >> j) & 1) {
            aOffT += (int)p.aLoadOff[j]; aSmT += p.aLoadSmT[j];
            bOffT += (int)p.bLoadOff[j]; bSmT += p.bLoadSmT[j];
        }
    }
    aSmT = frag_swz(aSmT); bSmT = frag_swz(bSmT);     // the swizzle is XOR-linear over the disjoint bit fields
    auto aOffI = [&](int i) { return slot_sum<LOGNT, TMB + KCB>(i, [&](int j) { return (int)p.aLoadOff[j]; }); };
    auto bOffI = [&](int i) { return slot_sum<LOGNT, TNB + KCB>(i, [&](int j) { return (int)p.bLoadOff[j]; }); };
    auto aSmI = [&](int i) { return frag_swz(slot_sum<LOGNT, TMB + KCB>(i, [&](int j) { return p.aLoadSmT[j]; })); };
    auto bSmI = [&](int i) { return frag_swz(slot_sum<LOGNT, TNB + KCB>(i, [&](int j) { return p.bLoadSmT[j]; })); };
    const long long hmask = (1ll << p.hb) - 1ll;
    const int nchunks = 1 << (p.nK - KCB);
    const int lsw = lane ^ (lane >> 3);
    const int fAo = (wm0 / 16) * 128 + lsw;           // float4 index: + (ks*MTC + mt)*128 + plane*32
    const int fBo = (wn0 / 8) * 64 + lsw;             // float4 index: + (ks*NTC + nt)*64 + half*32
    R2 ra[LA], rb[LB];
    auto stage = [&](float* sA, float* sB) {          // split into tf32 head / tail + scatter
#pragma unroll
        for (int i = 0; i < LA; ++i) {                // A: four planes re_hi, im_hi, re_lo, im_lo
            float* d = sA + (aSmT ^ aSmI(i));
            const float hr = to_tf32(ra[i].x), hi = to_tf32(ra[i].y);
            d[0] = hr; d[128] = hi; d[256] = to_tf32(ra[i].x - hr); d[384] = to_tf32(ra[i].y - hi);
        }
#pragma unroll
        for (int i = 0; i < LB; ++i) {                // B: [re b0 b1, im b0 b1] x {hi, lo}
            float* d = sB + (bSmT ^ bSmI(i));
            const float hr = to_tf32(rb[i].x), hi = to_tf32(rb[i].y);
            d[0] = hr; d[2] = hi; d[128] = to_tf32(rb[i].x - hr); d[130] = to_tf32(rb[i].y - hi);
        }
    };
    for (long long tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
        const long long u = tile >> p.hb;
        const unsigned long long hh = (unsigned long long)(tile & hmask);
        const R2* Ap = A + u * p.sUA + segeval(p.sAhi, p.nsAhi, hh);
        const R2* Bp = B + u * p.sUB + segeval(p.sBhi, p.nsBhi, hh);
        R2* Cp = C + u * p.sUC + segeval(p.sChi, p.nsChi, hh);
        float cre[MT][NTL][4], cim[MT][NTL][4];
#pragma unroll
        for (int i = 0; i < MT; ++i)
#pragma unroll
            for (int j = 0; j < NTL; ++j)
#pragma unroll
                for (int e = 0; e < 4; ++e) { cre[i][j][e] = 0.f; cim[i][j][e] = 0.f; }
#pragma unroll
        for (int i = 0; i < LA; ++i) ra[i] = __ldg(Ap + aOffT + aOffI(i));
#pragma unroll
        for (int i = 0; i < LB; ++i) rb[i] = __ldg(Bp + bOffT + bOffI(i));
        stage(smem_f, smem_f + SA);              // (the previous tile ended on a barrier: both stages are free)
        __syncthreads();
        // two-stage pipeline, one barrier per chunk: global loads of chunk ch+1 are in flight while chunk ch
        // is multiplied out of stage ch&1, then split + stored into the other stage
        for (int ch = 0; ch < nchunks; ++ch) {
            const float* sA = smem_f + (ch & 1) * (SA + SB);
            const float* sB = sA + SA;
            if (ch + 1 < nchunks) {
                const unsigned long long kb = (unsigned long long)(ch + 1) << KCB;
                const R2* An = Ap + segeval(p.kA, p.nkA, kb);
                const R2* Bn = Bp + segeval(p.kB, p.nkB, kb);
#pragma unroll
                for (int i = 0; i < LA; ++i) ra[i] = __ldg(An + aOffT + aOffI(i));
#pragma unroll
                for (int i = 0; i < LB; ++i) rb[i] = __ldg(Bn + bOffT + bOffI(i));
            }
            const float4* fA = reinterpret_cast<const float4*>(sA) + fAo;
            const float4* fB = reinterpret_cast<const float4*>(sB) + fBo;
#pragma unroll 1
            for (int ks = 0; ks < BK / 8; ++ks) {
                float4 ahr[MT], ahi[MT], alr[MT], ali[MT];
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) {
                    const float4* q = fA + (ks * MTC + mt) * 128;
                    ahr[mt] = q[0]; ahi[mt] = q[32]; alr[mt] = q[64]; ali[mt] = q[96];
                }
#pragma unroll
                for (int nt = 0; nt < NTL; ++nt) {
                    const float4* q = fB + (ks * NTC + nt) * 64;
                    const float4 bh = q[0], bl = q[32];            // {re b0, re b1, im b0, im b1}
                    const float nh0 = fneg_bits(bh.z), nh1 = fneg_bits(bh.w);
                    const float nl0 = fneg_bits(bl.z), nl1 = fneg_bits(bl.w);
                    // Each k8 partial product is chained inside the tensor core from a ZERO accumulator (cross
                    // terms first, head x head last) and then added to the running sums with FADD: the MMA adds
                    // with truncation, and chaining all of K through it biases every sum towards zero by
                    // ~2^-24 per MMA (measured 2.4e-5 relative on K = 2^9 trees, against 3e-6 this way).
                    float dre[MT][4], dim[MT][4];
#pragma unroll
                    for (int mt = 0; mt < MT; ++mt) {
                        mma_tf32_z(dre[mt], alr[mt], bh.x, bh.y);
                        mma_tf32_z(dim[mt], alr[mt], bh.z, bh.w);
                    }
#pragma unroll
                    for (int mt = 0; mt < MT; ++mt) {
                        mma_tf32(dre[mt], ahr[mt], bl.x, bl.y);
                        mma_tf32(dim[mt], ahr[mt], bl.z, bl.w);
                    }
#pragma unroll
                    for (int mt = 0; mt < MT; ++mt) {
                        mma_tf32(dre[mt], ali[mt], nh0, nh1);
                        mma_tf32(dim[mt], ali[mt], bh.x, bh.y);
                    }
#pragma unroll
                    for (int mt = 0; mt < MT; ++mt) {
                        mma_tf32(dre[mt], ahi[mt], nl0, nl1);
                        mma_tf32(dim[mt], ahi[mt], bl.x, bl.y);
                    }
#pragma unroll
                    for (int mt = 0; mt < MT; ++mt) {
                        mma_tf32(dre[mt], ahr[mt], bh.x, bh.y);
                        mma_tf32(dim[mt], ahr[mt], bh.z, bh.w);
                    }
#pragma unroll
                    for (int mt = 0; mt < MT; ++mt) {
                        mma_tf32(dre[mt], ahi[mt], nh0, nh1);
                        mma_tf32(dim[mt], ahi[mt], bh.x, bh.y);
                    }
#pragma unroll
                    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
                        for (int e = 0; e < 4; ++e) { cre[mt][nt][e] += dre[mt][e]; cim[mt][nt][e] += dim[mt][e]; }
                }
            }
            if (ch + 1 < nchunks) {
                float* nA = smem_f + ((ch + 1) & 1) * (SA + SB);
                stage(nA, nA + SA);
            }
            __syncthreads();
        }
        // epilogue: accumulator c0 (g, 2t) c1 (g, 2t+1) c2 (g+8, 2t) c3 (g+8, 2t+1) -> C through the tile-bit tables
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int ro = bit_offset<TMB>(wm0 + mt * 16 + g + h * 8, p.cM);
#pragma unroll
                for (int nt = 0; nt < NTL; ++nt) {
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        const int co = bit_offset<TNB>(wn0 + nt * 8 + 2 * t + j, p.cN);
                        Cp[ro + co] = make_float2(cre[mt][nt][h * 2 + j], cim[mt][nt][h * 2 + j]);
                    }
                }
            }
        }
    }
}

template <int TMB, int TNB>
__global__ void __launch_bounds__(((1 << (TMB + TNB)) / 512) * 32, 2)
gemm_dmma_kernel(const __grid_constant__ GemmParams p) {
    using R2 = double2;
    constexpr int KCB = 3;
    constexpr int BM = 1 << TMB, BN = 1 << TNB, BK = 1 << KCB;
    constexpr int WGM = BM / 16, WGN = BN / 32;
    constexpr int NT = 32 * WGM * WGN;
    constexpr int LOGNT = TMB + TNB - 4;
    constexpr int LDA = BM + 2, LDB = BN + 2;
    constexpr int LA = (BM * BK) / NT, LB = (BN * BK) / NT;
    constexpr int MT = 2, NTL = 4;                    // m8 / n8 MMA tiles per 16 x 32 warp tile
    static_assert(NT == (1 << LOGNT), "thread count");
    constexpr int SA = BK * LDA, SB = BK * LDB;       // elements per stage
    __shared__ __align__(16) R2 smem_d[2 * (SA + SB)]; // two stages of [A tile | B tile]
    const R2* __restrict__ A = reinterpret_cast<const R2*>(p.A);
    const R2* __restrict__ B = reinterpret_cast<const R2*>(p.B);
    R2* __restrict__ C = reinterpret_cast<R2*>(p.C);
    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int wm0 = (warp % WGM) * 16, wn0 = (warp / WGM) * 32;
    int aOffT = 0, aSmT = 0, bOffT = 0, bSmT = 0;
#pragma unroll
    for (int j = 0; j < LOGNT; ++j) {
        if ((tid >> j) & 1) {
            aOffT += (int)p.aLoadOff[j]; aSmT += p.aLoadSmT[j];
            bOffT += (int)p.bLoadOff[j]; bSmT += p.bLoadSmT[j];
        }
    }
    auto aOffI = [&](int i) { return slot_sum<LOGNT, TMB + KCB>(i, [&](int j) { return (int)p.aLoadOff[j]; }); };
    auto bOffI = [&](int i) { return slot_sum<LOGNT, TNB + KCB>(i, [&](int j) { return (int)p.bLoadOff[j]; }); };
    auto aSmI = [&](int i) { return slot_sum<LOGNT, TMB + KCB>(i, [&](int j) { return p.aLoadSmT[j]; }); };
    auto bSmI = [&](int i) { return slot_sum<LOGNT, TNB + KCB>(i, [&](int j) { return p.bLoadSmT[j]; }); };
    const long long hmask = (1ll << p.hb) - 1ll;
    const int nchunks = 1 << (p.nK - KCB);
    for (long long tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
        const long long u = tile >> p.hb;
        const unsigned long long hh = (unsigned long long)(tile & hmask);
        const R2* Ap = A + u * p.sUA + segeval(p.sAhi, p.nsAhi, hh);
        const R2* Bp = B + u * p.sUB + segeval(p.sBhi, p.nsBhi, hh);
        R2* Cp = C + u * p.sUC + segeval(p.sChi, p.nsChi, hh);
        double cre[MT][NTL][2], cim[MT][NTL][2];
#pragma unroll
        for (int i = 0; i < MT; ++i)
#pragma unroll
            for (int j = 0; j < NTL; ++j) { cre[i][j][0] = cre[i][j][1] = 0.0; cim[i][j][0] = cim[i][j][1] = 0.0; }
        R2 ra[LA], rb[LB];
#pragma unroll
        for (int i = 0; i < LA; ++i) ra[i] = __ldg(Ap + aOffT + aOffI(i));
#pragma unroll
        for (int i = 0; i < LB; ++i) rb[i] = __ldg(Bp + bOffT + bOffI(i));
#pragma unroll
        for (int i = 0; i < LA; ++i) smem_d[aSmT + aSmI(i)] = ra[i];     // (the previous tile ended on a barrier)
#pragma unroll
        for (int i = 0; i < LB; ++i) smem_d[SA + bSmT + bSmI(i)] = rb[i];
        __syncthreads();
        // two-stage pipeline, one barrier per chunk (see gemm_tf32x3_kernel)
        for (int ch = 0; ch < nchunks; ++ch) {
            const R2* sA = smem_d + (ch & 1) * (SA + SB);
            const R2* sB = sA + SA;
            if (ch + 1 < nchunks) {
                const unsigned long long kb = (unsigned long long)(ch + 1) << KCB;
                const R2* An = Ap + segeval(p.kA, p.nkA, kb);
                const R2* Bn = Bp + segeval(p.kB, p.nkB, kb);
#pragma unroll
                for (int i = 0; i < LA; ++i) ra[i] = __ldg(An + aOffT + aOffI(i));
#pragma unroll
                for (int i = 0; i < LB; ++i) rb[i] = __ldg(Bn + bOffT + bOffI(i));
            }
#pragma unroll
            for (int ks = 0; ks < BK / 4; ++ks) {
                // A fragment of an m8 tile: a0 (row g, k = t); B fragment of an n8 tile: b0 (k = t, n = g)
                const R2* ab = sA + (ks * 4 + t) * LDA + wm0 + g;
                const R2* bb = sB + (ks * 4 + t) * LDB + wn0 + g;
                R2 af[MT];
                double nai[MT];
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) { af[mt] = ab[mt * 8]; nai[mt] = -af[mt].y; }
#pragma unroll
                for (int nt = 0; nt < NTL; ++nt) {
                    const R2 bv = bb[nt * 8];
#pragma unroll
                    for (int mt = 0; mt < MT; ++mt) {
                        mma_f64(cre[mt][nt], af[mt].x, bv.x);
                        mma_f64(cim[mt][nt], af[mt].x, bv.y);
                        mma_f64(cre[mt][nt], nai[mt], bv.y);
                        mma_f64(cim[mt][nt], af[mt].y, bv.x);
                    }
                }
            }
            if (ch + 1 < nchunks) {
                R2* nA = smem_d + ((ch + 1) & 1) * (SA + SB);
#pragma unroll
                for (int i = 0; i < LA; ++i) nA[aSmT + aSmI(i)] = ra[i];
#pragma unroll
                for (int i = 0; i < LB; ++i) nA[SA + bSmT + bSmI(i)] = rb[i];
            }
            __syncthreads();
        }
        // epilogue: accumulator c0 (g, 2t) c1 (g, 2t+1) -> C through the tile-bit tables
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
            const int ro = bit_offset<TMB>(wm0 + mt * 8 + g, p.cM);
#pragma unroll
            for (int nt = 0; nt < NTL; ++nt) {
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const int co = bit_offset<TNB>(wn0 + nt * 8 + 2 * t + j, p.cN);
                    Cp[ro + co] = make_double2(cre[mt][nt][j], cim[mt][nt][j]);
                }
            }
        }
    }
}

// Shared-memory contribution of tile-index bit `b` (bit positions: [0, tb) = M (N) tile bits, [tb, tb + kcb)
// = K chunk bits) in the layouts of the tensor-core kernels; filled into GemmParams::a/bLoadSmT by the host.
int gemm_mma_smem_bit(int dtype, bool is_b, int tb, int b) {
    if (dtype != 0) {                                            // c64: [k][m] double2, row stride 2^tb + 2
        return b < tb ? (1 << b) : (1 << (b - tb)) * ((1 << tb) + 2);
    }
    const int c = b - tb;
    if (!is_b) {                                                 // c32 A: [k8 step][m16 tile][plane][lane][a0..a3] floats
        if (b < 3) return 16 << b;                               // g
        if (b == 3) return 1;                                    // row + 8 -> a1 / a3
        if (b < tb) return 512 << (b - 4);                       // m16 tile (4 planes x 128 floats)
        if (c < 2) return 4 << c;                                // t
        if (c == 2) return 2;                                    // k + 4 -> a2 / a3
        return ((1 << tb) / 16) * 512;                           // k8 step
    }
    if (b < 3) return 16 << b;                                   // c32 B: [k8 step][n8 tile][hi/lo][lane][re b0 b1, im b0 b1]
    if (b < tb) return 256 << (b - 3);
    if (c < 2) return 4 << c;
    if (c == 2) return 1;                                        // k + 4 -> b1
    return ((1 << tb) / 8) * 256;
}

int gemm_mma_threads(int dtype) { return dtype == 0 ? 128 : 256; }
// dynamic shared memory: two stages of (64 x 16 A + 64 x 16 B) x 4 tf32 planes for c32; c64 uses static storage
size_t gemm_mma_smem_bytes(int dtype) { return dtype == 0 ? 2 * (4096 + 4096) * sizeof(float) : 0; }

const void* gemm_mma_func(int dtype, int tmb, int tnb) {
    if (tmb != 6 || tnb != 6) return nullptr;
    return dtype == 0 ? (const void*)&gemm_tf32x3_kernel<6, 6> : (const void*)&gemm_dmma_kernel<6, 6>;
}

// Reduction-shaped nodes (few C elements, long K -- e.g. the root after the batched
// slice variables were summed early): one warp per C element, lanes stride over k,
// shuffle reduction.  Requires nC <= 8 (all C bits are "thread bits" in OpParams).
template <typename R2>
__global__ void __launch_bounds__(kThreads)
kreduce_kernel(const __grid_constant__ OpParams p) {
    const R2* __restrict__ A = reinterpret_cast<const R2*>(p.A);
    const R2* __restrict__ B = reinterpret_cast<const R2*>(p.B);
    R2* __restrict__ C = reinterpret_cast<R2*>(p.C);
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    const long long total = (long long)p.U << p.nC;
    const unsigned cmask = (1u << p.nC) - 1u;
    const long long K = 1ll << p.nK;
    for (long long o = warp; o < total; o += nwarps) {
        const long long u = o >> p.nC;
        const unsigned c = (unsigned)o & cmask;
        const R2* Ap = A + u * p.sUA + segeval(p.sAlo, p.nsAlo, c);
        const R2* Bp = B + u * p.sUB + segeval(p.sBlo, p.nsBlo, c);
        R2 acc0, acc1; acc0.x = acc0.y = acc1.x = acc1.y = 0;
        long long k = lane;
        for (; k + 32 < K; k += 64) {
            const R2 a0 = __ldg(Ap + segeval(p.kA, p.nkA, (unsigned long long)k));
            const R2 b0 = __ldg(Bp + segeval(p.kB, p.nkB, (unsigned long long)k));
            const R2 a1 = __ldg(Ap + segeval(p.kA, p.nkA, (unsigned long long)(k + 32)));
            const R2 b1 = __ldg(Bp + segeval(p.kB, p.nkB, (unsigned long long)(k + 32)));
            cmac(acc0, a0, b0); cmac(acc1, a1, b1);
        }
        for (; k < K; k += 32)
            cmac(acc0, __ldg(Ap + segeval(p.kA, p.nkA, (unsigned long long)k)),
                 __ldg(Bp + segeval(p.kB, p.nkB, (unsigned long long)k)));
        acc0.x += acc1.x; acc0.y += acc1.y;
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) {
            acc0.x += __shfl_xor_sync(0xffffffffu, acc0.x, s);
            acc0.y += __shfl_xor_sync(0xffffffffu, acc0.y, s);
        }
        if (lane == 0) C[u * p.sUC + c] = acc0;
    }
}

// Same, one 256-thread block per C element: for very long K with few outputs (the root of a
// GEMM-shaped tree: K = 2^20, one output per bitstring).
template <typename R2>
__global__ void __launch_bounds__(kThreads)
kreduce_block_kernel(const __grid_constant__ OpParams p) {
    const R2* __restrict__ A = reinterpret_cast<const R2*>(p.A);
    const R2* __restrict__ B = reinterpret_cast<const R2*>(p.B);
    R2* __restrict__ C = reinterpret_cast<R2*>(p.C);
    __shared__ R2 part[kThreads / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long total = (long long)p.U << p.nC;
    const unsigned cmask = (1u << p.nC) - 1u;
    const long long K = 1ll << p.nK;
    for (long long o = blockIdx.x; o < total; o += gridDim.x) {
        const long long u = o >> p.nC;
        const unsigned c = (unsigned)o & cmask;
        const R2* Ap = A + u * p.sUA + segeval(p.sAlo, p.nsAlo, c);
        const R2* Bp = B + u * p.sUB + segeval(p.sBlo, p.nsBlo, c);
        R2 acc0, acc1; acc0.x = acc0.y = acc1.x = acc1.y = 0;
        long long k = threadIdx.x;
        for (; k + kThreads < K; k += 2 * kThreads) {
            const R2 a0 = __ldg(Ap + segeval(p.kA, p.nkA, (unsigned long long)k));
            const R2 b0 = __ldg(Bp + segeval(p.kB, p.nkB, (unsigned long long)k));
            const R2 a1 = __ldg(Ap + segeval(p.kA, p.nkA, (unsigned long long)(k + kThreads)));
            const R2 b1 = __ldg(Bp + segeval(p.kB, p.nkB, (unsigned long long)(k + kThreads)));
            cmac(acc0, a0, b0); cmac(acc1, a1, b1);
        }
        for (; k < K; k += kThreads)
            cmac(acc0, __ldg(Ap + segeval(p.kA, p.nkA, (unsigned long long)k)),
                 __ldg(Bp + segeval(p.kB, p.nkB, (unsigned long long)k)));
        acc0.x += acc1.x; acc0.y += acc1.y;
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) {
            acc0.x += __shfl_xor_sync(0xffffffffu, acc0.x, s);
            acc0.y += __shfl_xor_sync(0xffffffffu, acc0.y, s);
        }
        __syncthreads();                         // part[] free again
        if (lane == 0) part[warp] = acc0;
        __syncthreads();
        if (threadIdx.x == 0) {
            R2 t = part[0];
#pragma unroll
            for (int w = 1; w < kThreads / 32; ++w) { t.x += part[w].x; t.y += part[w].y; }
            C[u * p.sUC + c] = t;
        }
    }
}

// Split-K variant for the root-like nodes of GEMM-shaped trees (K = 2^20 .. 2^30 against a handful of
// outputs): 2^p.kc blocks share one output, each reduces a contiguous 1 / 2^p.kc of the k range and adds its
// partial sum into C with atomicAdd (C is zeroed by a memset node in front of this launch).  With one block per
// output such a node leaves most of the SMs idle and is latency-bound on a few KB of loads in flight.
template <typename R2>
__global__ void __launch_bounds__(kThreads)
kreduce_split_kernel(const __grid_constant__ OpParams p) {
    const R2* __restrict__ A = reinterpret_cast<const R2*>(p.A);
    const R2* __restrict__ B = reinterpret_cast<const R2*>(p.B);
    R2* __restrict__ C = reinterpret_cast<R2*>(p.C);
    __shared__ R2 part[kThreads / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int sb = p.kc;                                         // log2 of the splits per output
    const long long total = ((long long)p.U << p.nC) << sb;
    const unsigned cmask = (1u << p.nC) - 1u;
    const long long Kc = 1ll << (p.nK - sb);
    for (long long w = blockIdx.x; w < total; w += gridDim.x) {
        const long long o = w >> sb;
        const long long k0 = (w & ((1ll << sb) - 1ll)) * Kc;
        const long long u = o >> p.nC;
        const unsigned c = (unsigned)o & cmask;
        const R2* Ap = A + u * p.sUA + segeval(p.sAlo, p.nsAlo, c);
        const R2* Bp = B + u * p.sUB + segeval(p.sBlo, p.nsBlo, c);
        R2 acc0, acc1, acc2, acc3;
        acc0.x = acc0.y = acc1.x = acc1.y = acc2.x = acc2.y = acc3.x = acc3.y = 0;
        long long k = threadIdx.x;
        for (; k + 3 * kThreads < Kc; k += 4 * kThreads) {       // 8 loads in flight per thread
            const R2 a0 = __ldg(Ap + segeval(p.kA, p.nkA, (unsigned long long)(k0 + k)));
            const R2 b0 = __ldg(Bp + segeval(p.kB, p.nkB, (unsigned long long)(k0 + k)));
            const R2 a1 = __ldg(Ap + segeval(p.kA, p.nkA, (unsigned long long)(k0 + k + kThreads)));
            const R2 b1 = __ldg(Bp + segeval(p.kB, p.nkB, (unsigned long long)(k0 + k + kThreads)));
            const R2 a2 = __ldg(Ap + segeval(p.kA, p.nkA, (unsigned long long)(k0 + k + 2 * kThreads)));
            const R2 b2 = __ldg(Bp + segeval(p.kB, p.nkB, (unsigned long long)(k0 + k + 2 * kThreads)));
            const R2 a3 = __ldg(Ap + segeval(p.kA, p.nkA, (unsigned long long)(k0 + k + 3 * kThreads)));
            const R2 b3 = __ldg(Bp + segeval(p.kB, p.nkB, (unsigned long long)(k0 + k + 3 * kThreads)));
            cmac(acc0, a0, b0); cmac(acc1, a1, b1); cmac(acc2, a2, b2); cmac(acc3, a3, b3);
        }
        for (; k < Kc; k += kThreads)
            cmac(acc0, __ldg(Ap + segeval(p.kA, p.nkA, (unsigned long long)(k0 + k))),
                 __ldg(Bp + segeval(p.kB, p.nkB, (unsigned long long)(k0 + k))));
        acc0.x += acc1.x + acc2.x + acc3.x; acc0.y += acc1.y + acc2.y + acc3.y;
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) {
            acc0.x += __shfl_xor_sync(0xffffffffu, acc0.x, s);
            acc0.y += __shfl_xor_sync(0xffffffffu, acc0.y, s);
        }
        __syncthreads();                         // part[] free again
        if (lane == 0) part[warp] = acc0;
        __syncthreads();
        if (threadIdx.x == 0) {
            R2 t = part[0];
#pragma unroll
            for (int w2 = 1; w2 < kThreads / 32; ++w2) { t.x += part[w2].x; t.y += part[w2].y; }
            atomicAdd(&C[u * p.sUC + c].x, t.x);
            atomicAdd(&C[u * p.sUC + c].y, t.y);
        }
    }
}

const void* kreduce_split_func(int dtype) {
    return dtype == 0 ? (const void*)&kreduce_split_kernel<float2> : (const void*)&kreduce_split_kernel<double2>;
}

const void* kreduce_block_func(int dtype) {
    return dtype == 0 ? (const void*)&kreduce_block_kernel<float2> : (const void*)&kreduce_block_kernel<double2>;
}

const void* kreduce_func(int dtype) {
    return dtype == 0 ? (const void*)&kreduce_kernel<float2> : (const void*)&kreduce_kernel<double2>;
}

// Output leaves: one-hot (or +/-) vectors selected by the bitstring
// (docs/src/users_guide.md:149-158, docs/src/basics.md:55-63).
template <typename R2>
__global__ void outleaf_kernel(R2* base, const OutLeafDesc* __restrict__ d, const unsigned char* __restrict__ bits,
                               int n_outputs, long long amp0, long long n) {
    const OutLeafDesc L = d[blockIdx.y];
    const long long total = n << L.span_bits;
    const long long mask = (1ll << L.span_bits) - 1ll;
    R2* out = base + L.offset_per_amp * n;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const long long u = i >> L.span_bits;
        const long long j = i & mask;
        const unsigned char val = bits[(amp0 + u) * n_outputs + (L.out_idx - 1)];
        R2 v; v.x = 0; v.y = 0;
        if (val == 0) v.x = (j == 0);
        else if (val == 1) v.x = (j == 1);
        else if (val == 2) v.x = (j < 2);
        else v.x = j == 0 ? 1 : (j == 1 ? -1 : 0);
        out[i] = v;
    }
}

const void* outleaf_func(int dtype) {
    return dtype == 0 ? (const void*)&outleaf_kernel<float2> : (const void*)&outleaf_kernel<double2>;
}

// Sum of the saved scalar over the batched slice bits, in double, one warp per
// bitstring (deterministic order).
template <typename R2>
__global__ void reduce_root_kernel(const R2* __restrict__ root, long long sU, int span_bits, long long n,
                                   double scale, double* __restrict__ acc, long long amp0) {
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    const long long len = 1ll << span_bits;
    for (long long u = warp; u < n; u += nwarps) {
        double sx = 0, sy = 0;
        const R2* r = root + u * sU;
        for (long long i = lane; i < len; i += 32) {
            const R2 v = __ldg(r + i);
            sx += (double)v.x; sy += (double)v.y;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            sx += __shfl_xor_sync(0xffffffffu, sx, o);
            sy += __shfl_xor_sync(0xffffffffu, sy, o);
        }
        if (lane == 0) {
            acc[2 * (amp0 + u)] += scale * sx;
            acc[2 * (amp0 + u) + 1] += scale * sy;
        }
    }
}

const void* reduce_root_func(int dtype) {
    return dtype == 0 ? (const void*)&reduce_root_kernel<float2> : (const void*)&reduce_root_kernel<double2>;
}

// Tensor-valued root (open network): one thread per (bitstring, output element).  The output index runs over the
// saved tensor's TRUE extents in Julia order; its modes sit at arbitrary bit positions of the (power-of-two padded)
// lowered root, and the batched slice bits still open in the root are summed here, in double.
template <typename R2>
__global__ void reduce_root_open_kernel(const R2* __restrict__ root, long long sU, long long n, double scale,
                                        double* __restrict__ acc, long long amp0, const __grid_constant__ RootDesc d) {
    const long long total = n * d.elems;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long u = i / d.elems, o = i - u * d.elems;
        long long rest = o, addr = 0;
        for (int m = 0; m < d.n_modes; ++m) {
            const long long idx = rest % d.ext[m];
            rest /= d.ext[m];
            addr |= idx << d.pos[m];
        }
        const R2* r = root + u * sU + addr;
        double sx = 0, sy = 0;
        for (long long v = 0; v < (1ll << d.vtotal); ++v) {
            long long va = 0;
            int sh = 0;
            for (int s = 0; s < d.n_vseg; ++s) {
                va |= ((v >> sh) & ((1ll << d.vbits[s]) - 1ll)) << d.vpos[s];
                sh += d.vbits[s];
            }
            const R2 x = __ldg(r + va);
            sx += (double)x.x; sy += (double)x.y;
        }
        acc[2 * ((amp0 + u) * d.elems + o)] += scale * sx;
        acc[2 * ((amp0 + u) * d.elems + o) + 1] += scale * sy;
    }
}

const void* reduce_root_open_func(int dtype) {
    return dtype == 0 ? (const void*)&reduce_root_open_kernel<float2> : (const void*)&reduce_root_open_kernel<double2>;
}

template <typename R2>
__global__ void finalize_kernel(const double* __restrict__ acc, R2* __restrict__ out, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) {
        R2 v;
        v.x = (typename Real<R2>::T)acc[2 * i];
        v.y = (typename Real<R2>::T)acc[2 * i + 1];
        out[i] = v;
    }
}

const void* finalize_func(int dtype) {
    return dtype == 0 ? (const void*)&finalize_kernel<float2> : (const void*)&finalize_kernel<double2>;
}

}  // namespace qxb
