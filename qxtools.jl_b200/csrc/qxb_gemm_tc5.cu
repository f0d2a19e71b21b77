// qxb200 -- tcgen05 / TMEM ComplexF32 GEMM for GEMM-shaped contraction nodes (sm_100a).
//
// One real TF32 UMMA GEMM does the whole complex product (brought up stand-alone in scripts/microbench/tc5_cgemm.cu,
// first run on a B200 in round 2: 1.5e-6 relative against fp64, 81 TFLOP/s complex-equivalent with this staging).
// With k' = 2k + p (p = 0 re, 1 im):
//   A'[m][k']             = the interleaved complex row of A
//   B'[n     ][2k, 2k+1]  = ( B_re, -B_im)   -> real parts of C in accumulator columns [0, 64)
//   B'[64 + n][2k, 2k+1]  = ( B_im,  B_re)   -> imaginary parts in columns [64, 128)
// fp32 accuracy from the 3xTF32 split x = hi + lo: hi*hi accumulates in TMEM columns [0, 128), hi*lo + lo*hi in
// columns [128, 256) (the small terms do not ride on the large partial sums); the epilogue adds the two.
//
// CTA tile = 128 M-only bits-rows x 64 complex N columns = UMMA 128 x 128 x 8 (kind::tf32, cta_group::1), K chunk of
// 16 complex = 32 real k' = 4 UMMA steps x 3 products per stage.  The operands of a contraction node are NOT matrices:
// row / column / k indices are scattered address bits (GemmParams: one global offset per tile-index bit).  So the
// staging is a gather: every thread loads 8 + 4 complex elements through the bit tables (registers, one chunk ahead),
// splits them and writes the canonical K-major no-swizzle core-matrix layout (8 rows x 16 bytes contiguous,
// LBO = 128 B between 16-byte K chunks, SBO = 1024 B between 8-row groups) -- "the permute fused into the load".
// One thread issues the MMAs; tcgen05.commit -> mbarrier frees a stage / signals the epilogue (bounded spins: a
// descriptor the hardware rejects must trap, not hang the GPU).  Epilogue: tcgen05.ld 32x32b.x32 -> registers -> C
// through the tile-bit tables.  Persistent over the tiles of the node.
#include <cuda_runtime.h>
#include <stdint.h>

#include "qxb_kernels.cuh"

namespace qxb {

namespace {

constexpr int kTmb = 7, kTnb = 6, kKcb = 4;
constexpr int kBM = 128, kBNC = 64;
constexpr int kTileBytes = 128 * 32 * 4;              // one operand plane: 128 rows x 32 floats
constexpr int kStageBytes = 4 * kTileBytes;           // A_hi, A_lo, B_hi, B_lo
constexpr int kStages = 2;
constexpr int kTmemCols = 256;
constexpr int kThreads5 = 256, kLogNT = 8;
constexpr int kLA = (kBM * 16) / kThreads5, kLB = (kBNC * 16) / kThreads5;   // complex elements per thread and chunk

__device__ __forceinline__ uint32_t s_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ float tf32_rna(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}
__device__ __forceinline__ void mbar_init5(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_wait5(uint32_t bar, uint32_t parity) {
    for (uint32_t spin = 0; spin < (1u << 27); ++spin) {
        uint32_t done;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (done) return;
    }
    __trap();
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_c, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_c), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
// K-major, no swizzle: start >> 4 at [0,14), LBO >> 4 at [16,30), SBO >> 4 at [32,46), version 1 at [46,48)
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr) {
    return (uint64_t)((addr & 0x3FFFF) >> 4) | ((uint64_t)(128u >> 4) << 16) | ((uint64_t)(1024u >> 4) << 32) | (1ull << 46);
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
          "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
          "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
}
__device__ __forceinline__ long long segeval5(const DSeg* s, int n, unsigned long long x) {
    long long r = 0;
    for (int i = 0; i < n; ++i) {
        const DSeg g = s[i];
        r |= (long long)(((x >> g.src) & ((1ull << g.len) - 1ull)) << g.dst);
    }
    return r;
}
template <int LO, int HI, typename F>
__device__ __forceinline__ int slot_sum5(int i, F f) {
    int o = 0;
#pragma unroll
    for (int j = LO; j < HI; ++j) if ((i >> (j - LO)) & 1) o += f(j);
    return o;
}
__device__ __forceinline__ void split_store(uint8_t* hi_plane, uint32_t off, float x, float y) {
    const float hx = tf32_rna(x), hy = tf32_rna(y);
    *reinterpret_cast<float2*>(hi_plane + off) = make_float2(hx, hy);
    *reinterpret_cast<float2*>(hi_plane + kTileBytes + off) = make_float2(tf32_rna(x - hx), tf32_rna(y - hy));
}

}  // namespace

__global__ void __launch_bounds__(kThreads5, 1)
gemm_tc5_kernel(const __grid_constant__ GemmParams p) {
    extern __shared__ __align__(1024) uint8_t smem5[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint8_t* ctl = smem5 + kStages * kStageBytes;
    const uint32_t bar_free0 = s_u32(ctl), bar_free1 = s_u32(ctl + 8), bar_done = s_u32(ctl + 16);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(ctl + 32);
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_u32(tmem_slot)), "r"(kTmemCols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid == 0) {
        mbar_init5(bar_free0, 1); mbar_init5(bar_free1, 1); mbar_init5(bar_done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tmem_slot;
    // instruction descriptor: D = f32 (1 << 4), A = B = tf32 (2 << 7, 2 << 10), both K-major, N >> 3 at [17,23), M >> 4 at [24,29)
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24);

    const float2* __restrict__ A = reinterpret_cast<const float2*>(p.A);
    const float2* __restrict__ B = reinterpret_cast<const float2*>(p.B);
    float2* __restrict__ C = reinterpret_cast<float2*>(p.C);
    // this thread's part of the gather: load index e = tid + 256 * i; bit j of e contributes a global offset and a
    // shared-memory byte offset (tables sorted by ascending global offset: consecutive threads read ascending addresses)
    int aOffT = 0, aSmT = 0, bOffT = 0, bSmT = 0;
#pragma unroll
    for (int j = 0; j < kLogNT; ++j) {
        if ((tid >> j) & 1) {
            aOffT += (int)p.aLoadOff[j]; aSmT += p.aLoadSmT[j];
            bOffT += (int)p.bLoadOff[j]; bSmT += p.bLoadSmT[j];
        }
    }
    auto aOffI = [&](int i) { return slot_sum5<kLogNT, kTmb + kKcb>(i, [&](int j) { return (int)p.aLoadOff[j]; }); };
    auto bOffI = [&](int i) { return slot_sum5<kLogNT, kTnb + kKcb>(i, [&](int j) { return (int)p.bLoadOff[j]; }); };
    auto aSmI = [&](int i) { return slot_sum5<kLogNT, kTmb + kKcb>(i, [&](int j) { return p.aLoadSmT[j]; }); };
    auto bSmI = [&](int i) { return slot_sum5<kLogNT, kTnb + kKcb>(i, [&](int j) { return p.bLoadSmT[j]; }); };

    const long long hmask = (1ll << p.hb) - 1ll;
    const int nchunks = 1 << (p.nK - kKcb);
    unsigned it = 0;                                   // chunks issued so far by this CTA (stage = it & 1)
    unsigned tile_no = 0;
    for (long long tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++tile_no) {
        const long long u = tile >> p.hb;
        const unsigned long long hh = (unsigned long long)(tile & hmask);
        const float2* Ap = A + u * p.sUA + segeval5(p.sAhi, p.nsAhi, hh);
        const float2* Bp = B + u * p.sUB + segeval5(p.sBhi, p.nsBhi, hh);
        float2* Cp = C + u * p.sUC + segeval5(p.sChi, p.nsChi, hh);
        float2 ra[kLA], rb[kLB];
#pragma unroll
        for (int i = 0; i < kLA; ++i) ra[i] = __ldg(Ap + aOffT + aOffI(i));
#pragma unroll
        for (int i = 0; i < kLB; ++i) rb[i] = __ldg(Bp + bOffT + bOffI(i));
        for (int c = 0; c < nchunks; ++c, ++it) {
            const int s = (int)(it & 1u);
            uint8_t* st = smem5 + s * kStageBytes;
            if (it >= (unsigned)kStages)                // the MMAs that read this stage two chunks ago are done
                mbar_wait5(s ? bar_free1 : bar_free0, ((it - kStages) >> 1) & 1u);
#pragma unroll
            for (int i = 0; i < kLA; ++i) split_store(st, (uint32_t)(aSmT + aSmI(i)), ra[i].x, ra[i].y);
#pragma unroll
            for (int i = 0; i < kLB; ++i) {
                const uint32_t o = (uint32_t)(bSmT + bSmI(i));
                split_store(st + 2 * kTileBytes, o, rb[i].x, -rb[i].y);              // row n:      (re, -im)
                split_store(st + 2 * kTileBytes, o + 8 * 1024, rb[i].y, rb[i].x);    // row 64 + n: (im,  re)
            }
            if (c + 1 < nchunks) {                      // next chunk's operands into registers while the MMAs run
                const unsigned long long kb = (unsigned long long)(c + 1) << kKcb;
                const float2* An = Ap + segeval5(p.kA, p.nkA, kb);
                const float2* Bn = Bp + segeval5(p.kB, p.nkB, kb);
#pragma unroll
                for (int i = 0; i < kLA; ++i) ra[i] = __ldg(An + aOffT + aOffI(i));
#pragma unroll
                for (int i = 0; i < kLB; ++i) rb[i] = __ldg(Bn + bOffT + bOffI(i));
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic stores -> visible to the MMA's async proxy
            __syncthreads();
            if (tid == 0) {
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t a_hi = s_u32(st), a_lo = a_hi + kTileBytes, b_hi = a_hi + 2 * kTileBytes, b_lo = a_hi + 3 * kTileBytes;
#pragma unroll
                for (int j = 0; j < 4; ++j) {              // UMMA K = 8 tf32 = two 16-byte chunks = 256 B further on
                    const uint32_t o = (uint32_t)j * 256u;
                    const uint32_t acc = (c > 0 || j > 0) ? 1u : 0u;
                    umma_tf32(tmem, smem_desc(a_hi + o), smem_desc(b_hi + o), idesc, acc);
                    umma_tf32(tmem + 128, smem_desc(a_hi + o), smem_desc(b_lo + o), idesc, acc);
                    umma_tf32(tmem + 128, smem_desc(a_lo + o), smem_desc(b_hi + o), idesc, 1u);
                }
                umma_commit(s ? bar_free1 : bar_free0);    // arrives when the MMAs above have consumed this stage
                if (c == nchunks - 1) umma_commit(bar_done);
            }
        }
        mbar_wait5(bar_done, tile_no & 1u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        {
            // warp w reads TMEM lanes (w % 4) * 32 .. +31 (row m = lane), complex columns (w / 4) * 32 .. +31
            const int q = warp & 3, h = warp >> 2;
            const uint32_t base = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(h * 32);
            uint32_t re[32], im[32], xre[32], xim[32];
            tmem_ld32(base, re); tmem_ld32(base + 64, im); tmem_ld32(base + 128, xre); tmem_ld32(base + 192, xim);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            int ro = 0;
            const int m = q * 32 + lane;
#pragma unroll
            for (int t = 0; t < kTmb; ++t) if ((m >> t) & 1) ro += (int)p.cM[t];
            const int coh = h ? (int)p.cN[5] : 0;
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                int co = coh;
#pragma unroll
                for (int t = 0; t < 5; ++t) if ((i >> t) & 1) co += (int)p.cN[t];
                Cp[ro + co] = make_float2(__uint_as_float(re[i]) + __uint_as_float(xre[i]), __uint_as_float(im[i]) + __uint_as_float(xim[i]));
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();                                // the accumulators may be overwritten by the next tile
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kTmemCols));
}

const void* gemm_tc5_func() { return (const void*)&gemm_tc5_kernel; }
size_t gemm_tc5_smem_bytes() { return (size_t)kStages * kStageBytes + 64; }
int gemm_tc5_threads() { return kThreads5; }

// Shared-memory BYTE offset contributed by tile-index bit `b` (bits [0, tb) = M (N) tile bits, [tb, tb + 4) = the
// 16 complex k of a chunk) in the canonical K-major core-matrix layout: row r, complex k c ->
// (r >> 3) * 1024 + (c >> 1) * 128 + (r & 7) * 16 + (c & 1) * 8.  Same for A (128 rows) and B (rows n and 64 + n).
int gemm_tc5_smem_bit(int tb, int b) {
    if (b < 3) return 16 << b;
    if (b < tb) return 1024 << (b - 3);
    const int c = b - tb;
    return c == 0 ? 8 : 128 << (c - 1);
}

}  // namespace qxb
