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

template <typename R2, int KC, int MA, int NB, bool ONE>
__global__ void __launch_bounds__(kThreads, 2)
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
    for (long long u = blockIdx.x; u < p.U; u += gridDim.x) {
        __syncthreads();                       // previous row fully consumed
        const R2* Au = A + u * p.sUA;
        const R2* Bu = B + u * p.sUB;
        for (long long i = tid; i < nA; i += kThreads) sA[i] = __ldg(Au + i);
        for (long long i = tid; i < nB; i += kThreads) sB[i] = __ldg(Bu + i);
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
    return nullptr;
}
template <typename R2>
static const void* pick_smem_kc(int kc, int ma, int nb, bool one) {
    switch (kc) {
    case 0: return pick_smem_tile<R2, 0>(ma, nb, one);
    case 1: return pick_smem_tile<R2, 1>(ma, nb, one);
    case 2: return pick_smem_tile<R2, 2>(ma, nb, one);
    default: return nullptr;
    }
}

// nullptr when there is no shared-memory variant for this shape
const void* contract_smem_func(int dtype, int kc, int ma, int nb, bool single_chunk) {
    return dtype == 0 ? pick_smem_kc<float2>(kc, ma, nb, single_chunk) : pick_smem_kc<double2>(kc, ma, nb, single_chunk);
}

template <typename R2, int KC, int MA, int NB>
static const void* pick_one(bool one) {
    return one ? (const void*)&contract_kernel<R2, KC, MA, NB, true> : (const void*)&contract_kernel<R2, KC, MA, NB, false>;
}
template <typename R2, int KC, int MA>
static const void* pick_nb(int nb, bool one) {
    switch (nb) {
    case 0: return pick_one<R2, KC, MA, 0>(one);
    case 1: return pick_one<R2, KC, MA, 1>(one);
    default: return pick_one<R2, KC, MA, 2>(one);
    }
}
template <typename R2, int KC>
static const void* pick_ma(int ma, int nb, bool one) {
    switch (ma) {
    case 0: return pick_nb<R2, KC, 0>(nb, one);
    case 1: return pick_nb<R2, KC, 1>(nb, one);
    default: return pick_nb<R2, KC, 2>(nb, one);
    }
}
template <typename R2>
static const void* pick_kc(int kc, int ma, int nb, bool one) {
    switch (kc) {
    case 0: return pick_ma<R2, 0>(ma, nb, one);
    case 1: return pick_ma<R2, 1>(ma, nb, one);
    case 2: return pick_ma<R2, 2>(ma, nb, one);
    default: return pick_ma<R2, 3>(ma, nb, one);
    }
}

const void* contract_func(int dtype, int kc, int ma, int nb, bool single_chunk) {
    return dtype == 0 ? pick_kc<float2>(kc, ma, nb, single_chunk) : pick_kc<double2>(kc, ma, nb, single_chunk);
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
