// qxb200 -- tiled split-K reduction for reduction-shaped nodes (sm_100a): few outputs (<= 2^8 per bitstring row), very
// long K.  Root-like nodes of GEMM-shaped trees (Sycamore-like 53 qubits, 12 cycles: 16 x 16 outputs, K = 2^24) went
// through kreduce_split_kernel, which walks the whole K range once PER OUTPUT: every operand element is read 16 times
// and the node ran at 58 GB/s (74.6 ms of a 227 ms slice, profiles/r2_summary.md).  Here a CTA takes a chunk of K,
// stages the operand tiles A[rows of A x chunk], B[rows of B x chunk] in shared memory ONCE (a "row" = the distinct
// operand addresses the C bits select), every thread accumulates one output over the chunk from shared memory, and the
// partial sums are combined with atomicAdd into a zeroed C.  HBM traffic = |A| + |B|, once.
#include <cuda_runtime.h>

#include "qxb_kernels.cuh"

namespace qxb {

namespace {
__device__ __forceinline__ long long kseg(const DSeg* s, int n, unsigned long long x) {
    long long r = 0;
    for (int i = 0; i < n; ++i) {
        const DSeg g = s[i];
        r |= (long long)(((x >> g.src) & ((1ull << g.len) - 1ull)) << g.dst);
    }
    return r;
}
template <typename R2>
__device__ __forceinline__ void kmac(R2& acc, const R2 a, const R2 b) {
    acc.x = fma(a.x, b.x, acc.x);
    acc.x = fma(-a.y, b.y, acc.x);
    acc.y = fma(a.x, b.y, acc.y);
    acc.y = fma(a.y, b.x, acc.y);
}
}  // namespace

template <typename R2>
__global__ void __launch_bounds__(kThreads)
kreduce_tile_kernel(const __grid_constant__ OpParams p, const __grid_constant__ KredTile t) {
    constexpr int KT = kKredTileK;
    constexpr int LD = KT + 1;                        // padded row: threads of a warp read different rows at the same k
    extern __shared__ __align__(16) unsigned char kred_smem[];
    R2* sA = reinterpret_cast<R2*>(kred_smem);
    R2* sB = sA + t.n_rows_a * LD;
    const R2* __restrict__ A = reinterpret_cast<const R2*>(p.A);
    const R2* __restrict__ B = reinterpret_cast<const R2*>(p.B);
    R2* __restrict__ C = reinterpret_cast<R2*>(p.C);
    const int tid = threadIdx.x;
    const int c = tid & ((1 << p.nC) - 1);            // this thread's output; the spare thread bits split the chunk's k
    const int grp = tid >> p.nC, ngrp = kThreads >> p.nC;
    const int ra = t.row_a[c] * LD, rb = t.row_b[c] * LD;
    // staging slots: element e = tid + 256 * i of a tile -> row e / KT, k e % KT (KT divides 256: kk is fixed per thread)
    const int kk = tid % KT, r0 = tid / KT;
    constexpr int RSTEP = kThreads / KT;
    const long long gAk = kseg(p.kA, p.nkA, (unsigned long long)kk), gBk = kseg(p.kB, p.nkB, (unsigned long long)kk);
    const long long nchunks = 1ll << (p.nK - kKredTileKBits);
    for (long long u = 0; u < p.U; ++u) {
        const R2* Au = A + u * p.sUA;
        const R2* Bu = B + u * p.sUB;
        R2 acc; acc.x = 0; acc.y = 0;
        for (long long ch = blockIdx.x; ch < nchunks; ch += gridDim.x) {
            const unsigned long long k0 = (unsigned long long)ch << kKredTileKBits;
            const long long gA0 = kseg(p.kA, p.nkA, k0) + gAk, gB0 = kseg(p.kB, p.nkB, k0) + gBk;   // disjoint bits: k0 | kk
            __syncthreads();                           // the previous chunk is consumed
            for (int r = r0; r < t.n_rows_a; r += RSTEP) sA[r * LD + kk] = __ldg(Au + t.off_a[r] + gA0);
            for (int r = r0; r < t.n_rows_b; r += RSTEP) sB[r * LD + kk] = __ldg(Bu + t.off_b[r] + gB0);
            __syncthreads();
#pragma unroll 8
            for (int k = grp; k < KT; k += ngrp) kmac(acc, sA[ra + k], sB[rb + k]);
        }
        // one atomic per thread and row (the spare-bit groups of an output simply add up in C)
        R2* dst = C + u * p.sUC + kseg(p.sClo, p.nsClo, (unsigned long long)c);
        atomicAdd(&dst->x, acc.x);
        atomicAdd(&dst->y, acc.y);
    }
}

// chunk base offsets: one segment of the K address map per lane of warp 0, OR-reduced (the maps of a scrambled operand
// have 20+ segments; evaluated by every thread per chunk they were most of the kernel's instructions)
__device__ __forceinline__ long long kseg_warp(const DSeg* sgs, int n, unsigned long long x, int lane) {
    unsigned long long v = 0;
    if (lane < n) { const DSeg g = sgs[lane]; v = ((x >> g.src) & ((1ull << g.len) - 1ull)) << g.dst; }
    const unsigned lo = __reduce_or_sync(0xffffffffu, (unsigned)v), hi = __reduce_or_sync(0xffffffffu, (unsigned)(v >> 32));
    return (long long)(((unsigned long long)hi << 32) | lo);
}

template <typename R2, int T>
__global__ void __launch_bounds__(kThreads)
kreduce_grid_kernel(const __grid_constant__ OpParams p, const __grid_constant__ KredTile t) {
    constexpr int KT = kKredTileK;
    constexpr int LD = KT + 1;
    extern __shared__ __align__(16) unsigned char kred_smem[];
    __shared__ long long s_off[2][2];                  // [slot][A, B]: base offsets of the chunk issued next
    const int nA = t.n_rows_a, nB = t.n_rows_b;
    const int stage_elems = (nA + nB) * LD;
    R2* sbuf = reinterpret_cast<R2*>(kred_smem);
    const R2* __restrict__ A = reinterpret_cast<const R2*>(p.A);
    const R2* __restrict__ B = reinterpret_cast<const R2*>(p.B);
    R2* __restrict__ C = reinterpret_cast<R2*>(p.C);
    const int tid = threadIdx.x, lane = tid & 31;
    const int n_tiles = (nA / T) * (nB / T);           // a power of two <= 256
    const int tile = tid & (n_tiles - 1), grp = tid / n_tiles, ngrp = kThreads / n_tiles;
    const int ia = (tile % (nA / T)) * T, ib = (tile / (nA / T)) * T;
    const int kk = tid % KT, r0 = tid / KT;
    constexpr int RSTEP = kThreads / KT;
    int kkB = kk;                                      // the threads staging B walk the chunk in B's address order
    if (t.kperm_set) {
        kkB = 0;
#pragma unroll
        for (int j = 0; j < kKredTileKBits; ++j) kkB |= ((kk >> j) & 1) << t.kperm_b[j];
    }
    const long long gAk = kseg(p.kA, p.nkA, (unsigned long long)kk), gBk = kseg(p.kB, p.nkB, (unsigned long long)kkB);
    const long long nchunks = 1ll << (p.nK - kKredTileKBits);
    const long long G = gridDim.x;
    const unsigned sbase = (unsigned)__cvta_generic_to_shared(sbuf);
    for (long long u = 0; u < p.U; ++u) {
        const R2* Au = A + u * p.sUA;
        const R2* Bu = B + u * p.sUB;
        auto offsets = [&](long long ch, int slot) {   // warp 0
            const unsigned long long k0 = (unsigned long long)ch << kKredTileKBits;
            const long long a = kseg_warp(p.kA, p.nkA, k0, lane), b = kseg_warp(p.kB, p.nkB, k0, lane);
            if (lane == 0) { s_off[slot][0] = a; s_off[slot][1] = b; }
        };
        auto issue = [&](int slot, int stage) {
            const long long gA0 = s_off[slot][0] + gAk, gB0 = s_off[slot][1] + gBk;   // disjoint bits
            const unsigned sa = sbase + (unsigned)(stage * stage_elems * sizeof(R2)), sb = sa + (unsigned)(nA * LD * sizeof(R2));
            for (int r = r0; r < nA; r += RSTEP) {
                const R2* src = Au + t.off_a[r] + gA0;
                const unsigned dst = sa + (unsigned)((r * LD + kk) * sizeof(R2));
                if constexpr (sizeof(R2) == 8) asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
                else asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
            }
            for (int r = r0; r < nB; r += RSTEP) {
                const R2* src = Bu + t.off_b[r] + gB0;
                const unsigned dst = sb + (unsigned)((r * LD + kkB) * sizeof(R2));
                if constexpr (sizeof(R2) == 8) asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
                else asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
            }
        };
        R2 acc[T * T];
#pragma unroll
        for (int q = 0; q < T * T; ++q) { acc[q].x = 0; acc[q].y = 0; }
        __syncthreads();                               // the previous row's last chunk and offsets are consumed
        if (tid < 32) { offsets(blockIdx.x, 0); offsets(blockIdx.x + G, 1); }
        __syncthreads();
        if ((long long)blockIdx.x < nchunks) issue(0, 0);
        asm volatile("cp.async.commit_group;" ::: "memory");
        int it = 0;
        for (long long ch = blockIdx.x; ch < nchunks; ch += G, ++it) {
            const int stage = it & 1;
            if (ch + G < nchunks) issue((it + 1) & 1, stage ^ 1);
            asm volatile("cp.async.commit_group;" ::: "memory");
            asm volatile("cp.async.wait_group 1;" ::: "memory");
            __syncthreads();
            const R2* sA = sbuf + stage * stage_elems + ia * LD;
            const R2* sB = sbuf + stage * stage_elems + (nA + ib) * LD;
#pragma unroll 2
            for (int k = grp; k < KT; k += ngrp) {
                R2 x[T], y[T];
#pragma unroll
                for (int q = 0; q < T; ++q) { x[q] = sA[q * LD + k]; y[q] = sB[q * LD + k]; }
#pragma unroll
                for (int j = 0; j < T; ++j)
#pragma unroll
                    for (int i = 0; i < T; ++i) kmac(acc[j * T + i], x[i], y[j]);
            }
            if (tid < 32) offsets(ch + 2 * G, it & 1);  // for the chunk issued in the next iteration
            __syncthreads();                           // this stage is free for the chunk after next
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
#pragma unroll
        for (int q = 0; q < T * T; ++q) {
            const int c = t.grid_c[(ia + (q % T)) * nB + ib + (q / T)];
            R2* dst = C + u * p.sUC + kseg(p.sClo, p.nsClo, (unsigned long long)c);
            atomicAdd(&dst->x, acc[q].x);
            atomicAdd(&dst->y, acc[q].y);
        }
    }
}

const void* kreduce_grid_func(int dtype, int tile) {
    if (tile == 4) return dtype == 0 ? (const void*)&kreduce_grid_kernel<float2, 4> : (const void*)&kreduce_grid_kernel<double2, 4>;
    return dtype == 0 ? (const void*)&kreduce_grid_kernel<float2, 2> : (const void*)&kreduce_grid_kernel<double2, 2>;
}

const void* kreduce_tile_func(int dtype) {
    return dtype == 0 ? (const void*)&kreduce_tile_kernel<float2> : (const void*)&kreduce_tile_kernel<double2>;
}


// ---------------------------------------------------------------------------------------------------------------
// FMA-pipe peak, measured on the device the library runs on: the roofline denominator for the launches whose
// intermediates live in shared memory (fused chain, row programs) -- those are bound by the FP64 / FP32 FMA pipe, not by
// HBM.  16 independent accumulators per thread, no memory traffic in the loop.
template <typename T>
__global__ void __launch_bounds__(256) fma_peak_kernel(T* out, int iters, T a, T b) {
    T acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = (T)(threadIdx.x + i);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[i] = fma(acc[i], a, b);
    }
    T s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += acc[i];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256) fma2_peak_kernel(float* out, int iters, float a, float b) {
    unsigned long long acc[16], av, bv;
    asm("mov.b64 %0, {%1, %1};" : "=l"(av) : "f"(a));
    asm("mov.b64 %0, {%1, %1};" : "=l"(bv) : "f"(b));
#pragma unroll
    for (int i = 0; i < 16; ++i) { const float x = (float)(threadIdx.x + i); asm("mov.b64 %0, {%1, %1};" : "=l"(acc[i]) : "f"(x)); }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(acc[i]) : "l"(av), "l"(bv));
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) { float x, y; asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(acc[i])); s += x + y; }
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// TFLOP/s (2 flops per FMA) of the FFMA (dtype 0), DFMA (dtype 1) or packed FFMA2 (dtype 2) pipe: best of 5 timed launches on `st`
double fma_peak_tflops(int dtype, int num_sms, cudaStream_t st) {
    const int blocks = num_sms * 8, iters = 4096;
    void* buf = nullptr;
    if (cudaMalloc(&buf, (size_t)blocks * 256 * 8) != cudaSuccess) return 0.0;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    double best = 0.0;
    for (int rep = 0; rep < 6; ++rep) {
        cudaEventRecord(e0, st);
        if (dtype == 0) fma_peak_kernel<float><<<blocks, 256, 0, st>>>((float*)buf, iters, 1.0000001f, 1e-9f);
        else if (dtype == 2) fma2_peak_kernel<<<blocks, 256, 0, st>>>((float*)buf, iters, 1.0000001f, 1e-9f);
        else fma_peak_kernel<double><<<blocks, 256, 0, st>>>((double*)buf, iters, 1.0000001, 1e-9);
        cudaEventRecord(e1, st);
        cudaEventSynchronize(e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        const double tf = (dtype == 2 ? 2.0 : 1.0) * 2.0 * 16.0 * iters * (double)blocks * 256.0 / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(buf);
    return best;
}

}  // namespace qxb

namespace qxb {

// ---------------------------------------------------------------------------------------------------------------
// "big x small" streaming nodes: a huge operand (2^27 - 2^30 elements) times a tiny one (<= 2^10), few N-only bits
// (<= 5) and few K bits (<= 5), no batch bits -- the bulk of a Sycamore-like depth-12 slice (24 nodes, 16 GB each).
// contract_kernel gives every thread a 4-wide register tile over N and walks the remaining N values as separate tiles:
// the big operand is read once PER N-TILE (4 - 8 times) and the node ran at 1.7 - 3.2 TB/s algorithmic.  Here a thread owns
// one position of the big operand's free index space and ALL 2^N outputs of it: it reads its 2^K elements of the big
// operand once (coalesced along the low bits, all loads in flight), multiplies by the small operand from shared memory
// ([k][n], broadcast reads) and writes 2^N outputs (coalesced).  HBM traffic = |big| + |C|, once.
// Accumulators of one position: NN complex outputs.  ComplexF32 uses the packed FFMA2 (fma.rn.f32x2, sm_100): per output
// P += (a.x, a.x) * (b.x, b.y) and Q += (a.y, a.y) * (b.x, b.y), result = (P.x - Q.y, P.y + Q.x) -- two issue slots per
// complex MAC instead of four (the scalar version of this kernel was issue-bound: 0.76 IPC per scheduler with 59 % FFMA,
// profiles/r2_summary.md), and the small operand needs no swizzled copy.
template <typename R2, int NN, bool PACK>
struct BsAcc {
    R2 acc[NN];
    __device__ __forceinline__ void clear() {
#pragma unroll
        for (int n = 0; n < NN; ++n) { acc[n].x = 0; acc[n].y = 0; }
    }
    __device__ __forceinline__ void mac_row(const R2 a, const R2* bk) {
#pragma unroll
        for (int n = 0; n < NN; ++n) kmac(acc[n], a, bk[n]);
    }
    __device__ __forceinline__ R2 get(int n) const { return acc[n]; }
};
template <int NN>
struct BsAcc<float2, NN, true> {
    unsigned long long P[NN], Q[NN];
    __device__ __forceinline__ void clear() {
#pragma unroll
        for (int n = 0; n < NN; ++n) { P[n] = 0ull; Q[n] = 0ull; }
    }
    __device__ __forceinline__ void mac_row(const float2 a, const float2* bk) {
        unsigned long long ax, ay;
        asm("mov.b64 %0, {%1, %1};" : "=l"(ax) : "f"(a.x));
        asm("mov.b64 %0, {%1, %1};" : "=l"(ay) : "f"(a.y));
        const unsigned long long* b = reinterpret_cast<const unsigned long long*>(bk);
#pragma unroll
        for (int n = 0; n < NN; ++n) {
            const unsigned long long bv = b[n];
            asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(P[n]) : "l"(ax), "l"(bv));
            asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(Q[n]) : "l"(ay), "l"(bv));
        }
    }
    __device__ __forceinline__ float2 get(int n) const {
        float px, py, qx, qy;
        asm("mov.b64 {%0, %1}, %2;" : "=f"(px), "=f"(py) : "l"(P[n]));
        asm("mov.b64 {%0, %1}, %2;" : "=f"(qx), "=f"(qy) : "l"(Q[n]));
        return make_float2(px - qy, py + qx);
    }
};

// KKT > 0: the K extent is a compile-time constant (every loop unrolls, the offset tables are read as immediate
// constant-bank operands, no tail predicates: ~18 % fewer instructions on an issue-bound kernel); KKT = 0: run time.
template <typename R2, int NN, int RK, bool PACK, int KKT>
__global__ void __launch_bounds__(256)
bigsmall_kernel(const __grid_constant__ BigSmallParams p) {
    extern __shared__ __align__(16) unsigned char bs_smem[];
    R2* sB = reinterpret_cast<R2*>(bs_smem);                       // [n_hi][k][n]
    const R2* __restrict__ A = reinterpret_cast<const R2*>(p.big);
    const R2* __restrict__ B = reinterpret_cast<const R2*>(p.small_);
    R2* __restrict__ C = reinterpret_cast<R2*>(p.C);
    const int KK = KKT > 0 ? KKT : (1 << p.nK);
    const int n_small = (KK * NN) << p.nNhi;
    // position index = [8 thread bits][n_hi bits][block bits]: the thread part of both addresses is evaluated once, the
    // block part (with the N bits beyond the register tile) is uniform over the CTA
    const long long a_lo = kseg(p.tA, p.ntA, (unsigned long long)threadIdx.x);
    const long long c_lo = kseg(p.tC, p.ntC, (unsigned long long)threadIdx.x);
    const long long n_blk = p.n_pos >> 8;
    for (long long u = 0; u < p.U; ++u) {
        __syncthreads();
        for (int i = threadIdx.x; i < n_small; i += blockDim.x) {
            const int n = i % NN, k = (i / NN) & (KK - 1), nh = i / (NN * KK);
            sB[i] = __ldg(B + u * p.sUsmall + p.bK[k] + p.bN[n] + p.bH[nh]);
        }
        __syncthreads();
        const R2* Au = A + u * p.sUbig + a_lo;
        R2* Cu = C + u * p.sUC + c_lo;
        const R2* Ab = Au + kseg(p.tA, p.ntA, (unsigned long long)blockIdx.x << 8);
        if constexpr (KKT > 0) {
            // RK elements of the big operand per round, next round (or the next block's first) in flight during the FMAs
            constexpr int NR = (KKT + RK - 1) / RK;
            R2 a[RK], an[RK];
            if ((long long)blockIdx.x < n_blk) {
#pragma unroll
                for (int q = 0; q < RK; ++q) if (q < KKT) a[q] = __ldg(Ab + p.aK[q]);
            }
            for (long long blk = blockIdx.x; blk < n_blk; blk += gridDim.x) {
                R2* Cb = Cu + kseg(p.tC, p.ntC, (unsigned long long)blk << 8);
                const R2* sBh = sB + (int)(blk & ((1 << p.nNhi) - 1)) * (KKT * NN);
                const long long nxt = blk + gridDim.x;
                const R2* Anext = Au + kseg(p.tA, p.ntA, (unsigned long long)(nxt < n_blk ? nxt : blk) << 8);
                BsAcc<R2, NN, PACK> acc;
                acc.clear();
#pragma unroll
                for (int r = 0; r < NR; ++r) {
                    if (r + 1 < NR) {
#pragma unroll
                        for (int q = 0; q < RK; ++q) if ((r + 1) * RK + q < KKT) an[q] = __ldg(Ab + p.aK[(r + 1) * RK + q]);
                    } else {
#pragma unroll
                        for (int q = 0; q < RK; ++q) if (q < KKT) an[q] = __ldg(Anext + p.aK[q]);
                    }
#pragma unroll
                    for (int q = 0; q < RK; ++q) if (r * RK + q < KKT) acc.mac_row(a[q], sBh + (r * RK + q) * NN);
#pragma unroll
                    for (int q = 0; q < RK; ++q) a[q] = an[q];
                }
                Ab = Anext;
#pragma unroll
                for (int n = 0; n < NN; ++n) Cb[p.cN[n]] = acc.get(n);
            }
        } else {
        // RK elements of the big operand per round (enough bytes in flight per SM to cover the HBM latency); the next
        // round's loads -- the first round of the CTA's NEXT block of positions after the last one -- are issued before
        // this round's FMAs, so loads stay in flight across blocks
        R2 a[RK], an[RK];
        if ((long long)blockIdx.x < n_blk) {
#pragma unroll
            for (int q = 0; q < RK; ++q) a[q] = (q < KK) ? __ldg(Ab + p.aK[q]) : R2{0, 0};
        }
        for (long long blk = blockIdx.x; blk < n_blk; blk += gridDim.x) {
            R2* Cb = Cu + kseg(p.tC, p.ntC, (unsigned long long)blk << 8);
            const R2* sBh = sB + (int)(blk & ((1 << p.nNhi) - 1)) * (KK * NN);
            const long long nxt = blk + gridDim.x;
            const R2* Anext = Au + kseg(p.tA, p.ntA, (unsigned long long)(nxt < n_blk ? nxt : blk) << 8);
            BsAcc<R2, NN, PACK> acc;
            acc.clear();
            for (int k0 = 0; k0 < KK; k0 += RK) {
                const bool last = k0 + RK >= KK;
                const R2* src = last ? Anext : Ab;
                const int kb = last ? 0 : k0 + RK;
                if (!last || nxt < n_blk) {
#pragma unroll
                    for (int q = 0; q < RK; ++q) an[q] = (kb + q < KK) ? __ldg(src + p.aK[kb + q]) : R2{0, 0};
                }
#pragma unroll
                for (int q = 0; q < RK; ++q) {
                    if (k0 + q >= KK) break;
                    acc.mac_row(a[q], sBh + (k0 + q) * NN);
                }
#pragma unroll
                for (int q = 0; q < RK; ++q) a[q] = an[q];
            }
            Ab = Anext;
#pragma unroll
            for (int n = 0; n < NN; ++n) Cb[p.cN[n]] = acc.get(n);
        }
        }
    }
}

// Same node, the big operand through 1-D TMA bulk copies: when the 8 thread bits of the position index are the 8 lowest
// address bits of the big operand, the 256 positions of a CTA at one k are one contiguous 2 KB (4 KB) run, and a stage
// of 16 (8) k values is 32 KB brought in by one elected thread with mbarrier complete_tx.  Three stages per CTA and two
// CTAs per SM keep ~190 KB of reads in flight per SM without holding them in registers (the register version has
// 24 - 32 KB in flight and ran at 4.4 TB/s on the 2^4 x 2^4 nodes: FMA time + HBM time, not overlapped).
namespace {
__device__ __forceinline__ unsigned bs_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bs_mbar_wait(unsigned bar, unsigned parity) {
    unsigned done = 0;
    while (!done)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
}
}  // namespace

template <typename R2, int NN>
__global__ void __launch_bounds__(256, 2)
bigsmall_tma_kernel(const __grid_constant__ BigSmallParams p, const int n_stages) {
    constexpr int KC = kBigSmallStageBytes / (256 * (int)sizeof(R2));     // k values per stage
    extern __shared__ __align__(128) unsigned char bs_smem[];
    const int KK = 1 << p.nK;
    const int kcn = KK < KC ? KK : KC, nkc = KK / kcn;                    // k values per fill, fills per block of positions
    const int n_small = (KK * NN) << p.nNhi;
    R2* sB = reinterpret_cast<R2*>(bs_smem + (size_t)n_stages * kBigSmallStageBytes);
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(bs_smem + (size_t)n_stages * kBigSmallStageBytes +
                                                                     (((size_t)n_small * sizeof(R2) + 15) & ~(size_t)15));
    const R2* __restrict__ A = reinterpret_cast<const R2*>(p.big);
    const R2* __restrict__ B = reinterpret_cast<const R2*>(p.small_);
    R2* __restrict__ C = reinterpret_cast<R2*>(p.C);
    const int tid = threadIdx.x;
    if (tid == 0) {
        for (int s = 0; s < n_stages; ++s)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bs_u32(bars + s)), "r"(1u) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < n_small; i += 256) {
        const int n = i % NN, k = (i / NN) & (KK - 1), nh = i / (NN * KK);
        sB[i] = __ldg(B + p.bK[k] + p.bN[n] + p.bH[nh]);
    }
    __syncthreads();
    const long long n_blk = p.n_pos >> 8;
    const long long my_blocks = (long long)blockIdx.x < n_blk ? (n_blk - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const long long n_fill = my_blocks * nkc;
    const unsigned run_bytes = 256u * (unsigned)sizeof(R2);
    auto issue = [&](long long f) {                                       // thread 0
        const long long blk = blockIdx.x + (f / nkc) * (long long)gridDim.x;
        const int kc = (int)(f % nkc), s = (int)(f % n_stages);
        const R2* Ab = A + kseg(p.tA, p.ntA, (unsigned long long)blk << 8);
        const unsigned bar = bs_u32(bars + s), dst = bs_u32(bs_smem + (size_t)s * kBigSmallStageBytes);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(run_bytes * (unsigned)kcn) : "memory");
        for (int q = 0; q < kcn; ++q)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(dst + (unsigned)q * run_bytes), "l"(Ab + p.aK[kc * kcn + q]), "r"(run_bytes), "r"(bar) : "memory");
    };
    if (tid == 0)
        for (long long f = 0; f < n_fill && f < n_stages; ++f) issue(f);
    const long long c_lo = kseg(p.tC, p.ntC, (unsigned long long)tid);
    R2 acc[NN];
    for (long long f = 0; f < n_fill; ++f) {
        const long long blk = blockIdx.x + (f / nkc) * (long long)gridDim.x;
        const int kc = (int)(f % nkc), s = (int)(f % n_stages);
        if (kc == 0) {
#pragma unroll
            for (int n = 0; n < NN; ++n) { acc[n].x = 0; acc[n].y = 0; }
        }
        bs_mbar_wait(bs_u32(bars + s), (unsigned)((f / n_stages) & 1));
        const R2* st = reinterpret_cast<const R2*>(bs_smem + (size_t)s * kBigSmallStageBytes) + tid;
        const R2* bk = sB + (int)(blk & ((1 << p.nNhi) - 1)) * (KK * NN) + kc * kcn * NN;
#pragma unroll 4
        for (int q = 0; q < kcn; ++q) {
            const R2 a = st[q * 256];
#pragma unroll
            for (int n = 0; n < NN; ++n) kmac(acc[n], a, bk[q * NN + n]);
        }
        __syncthreads();                                                  // every thread is done with stage s
        if (tid == 0 && f + n_stages < n_fill) issue(f + n_stages);
        if (kc == nkc - 1) {
            R2* Cb = C + c_lo + kseg(p.tC, p.ntC, (unsigned long long)blk << 8);
#pragma unroll
            for (int n = 0; n < NN; ++n) Cb[p.cN[n]] = acc[n];
        }
    }
}

const void* bigsmall_tma_func(int dtype, int n_bits) {
    if (dtype == 0) {
        switch (n_bits) {
        case 1: return (const void*)&bigsmall_tma_kernel<float2, 2>;
        case 2: return (const void*)&bigsmall_tma_kernel<float2, 4>;
        case 3: return (const void*)&bigsmall_tma_kernel<float2, 8>;
        case 4: return (const void*)&bigsmall_tma_kernel<float2, 16>;
        case 5: return (const void*)&bigsmall_tma_kernel<float2, 32>;
        default: return nullptr;
        }
    }
    switch (n_bits) {
    case 1: return (const void*)&bigsmall_tma_kernel<double2, 2>;
    case 2: return (const void*)&bigsmall_tma_kernel<double2, 4>;
    case 3: return (const void*)&bigsmall_tma_kernel<double2, 8>;
    case 4: return (const void*)&bigsmall_tma_kernel<double2, 16>;
    default: return nullptr;
    }
}

template <typename R2, int NN, int RK, bool PACK>
static const void* bigsmall_pick_k(int k_bits) {
    switch (k_bits) {
    case 3: return (const void*)&bigsmall_kernel<R2, NN, RK, PACK, 8>;
    case 4: return (const void*)&bigsmall_kernel<R2, NN, RK, PACK, 16>;
    case 5: return (const void*)&bigsmall_kernel<R2, NN, RK, PACK, 32>;
    default: return (const void*)&bigsmall_kernel<R2, NN, RK, PACK, 0>;
    }
}

const void* bigsmall_func(int dtype, int n_bits, int k_bits, bool packed) {
    if (dtype == 0 && packed) {
        switch (n_bits) {
        case 1: return bigsmall_pick_k<float2, 2, 8, true>(k_bits);
        case 2: return bigsmall_pick_k<float2, 4, 8, true>(k_bits);
        case 3: return bigsmall_pick_k<float2, 8, 8, true>(k_bits);
        case 4: return bigsmall_pick_k<float2, 16, 4, true>(k_bits);
        default: return nullptr;
        }
    }
    if (dtype == 0) {
        switch (n_bits) {
        case 1: return bigsmall_pick_k<float2, 2, 8, false>(k_bits);
        case 2: return bigsmall_pick_k<float2, 4, 8, false>(k_bits);
        case 3: return bigsmall_pick_k<float2, 8, 8, false>(k_bits);
        case 4: return bigsmall_pick_k<float2, 16, 8, false>(k_bits);
        case 5: return bigsmall_pick_k<float2, 32, 4, false>(0);   // measured: the unrolled-K variant is slower at 2^5 outputs (4.9 vs 3.4 ms)
        default: return nullptr;
        }
    }
    switch (n_bits) {
    case 1: return bigsmall_pick_k<double2, 2, 8, false>(k_bits);
    case 2: return bigsmall_pick_k<double2, 4, 8, false>(k_bits);
    case 3: return bigsmall_pick_k<double2, 8, 4, false>(k_bits);
    case 4: return bigsmall_pick_k<double2, 16, 4, false>(k_bits);
    default: return nullptr;
    }
}

}  // namespace qxb
