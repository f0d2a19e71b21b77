// tcgen05 3xTF32 complex GEMM -- stand-alone bring-up kernel for the ComplexF32 GEMM-shaped nodes.
//
// STATUS: written and compiled for sm_100a in a session WITHOUT GPU access (ptxas: 134 registers, no spills; SASS: 12 UTCHMMA per K chunk, UTCBAR commits, LDTM epilogue loads);
// it has never run.  It is not linked into libqxb200.so.  Purpose: debug descriptor / layout choices and measure the
// ceiling in isolation (one `gpurun` call) before the operand staging of gemm_tf32x3_kernel is moved onto it.
//
//   C[m][n] = sum_k A[m][k] * B[n][k]          A: [M][K], B: [N][K] (K contiguous), C: [M][N], all complex<float>
//
// One real TF32 GEMM does the whole complex product: with k' = 2k + p (p = 0 re, 1 im)
//   A'[m][k']            = the interleaved complex row of A as it lies in memory
//   B'[n     ][2k, 2k+1] = ( B_re, -B_im)    -> real parts of C in accumulator columns [0, 64)
//   B'[64 + n][2k, 2k+1] = ( B_im,  B_re)    -> imaginary parts in columns [64, 128)
// and fp32 accuracy comes from the 3xTF32 split x = hi + lo (both tf32): hi*hi into one TMEM accumulator, hi*lo + lo*hi
// into a second one (the small terms do not ride on the large partial sums), added in the epilogue.
//
// CTA tile 128 (m) x 64 complex (n) = UMMA 128 x 128 x 8 (kind::tf32, cta_group::1), K chunk of 16 complex = 32 real k'
// = 4 UMMA steps x 3 products per stage; two shared-memory stages, operands staged by all 256 threads into the
// canonical K-major INTERLEAVE (no swizzle) layout of cute/arch/mma_sm100_desc.hpp: core matrix = 8 rows x 16 bytes
// stored contiguously, LBO = 128 B between the core matrices of consecutive 16-byte K chunks, SBO = 1024 B between
// 8-row groups; one thread issues the MMAs, tcgen05.commit releases a stage / signals the epilogue through mbarriers
// (bounded spins: a wrong descriptor must trap, not hang the box).
//
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tc5_cgemm tc5_cgemm.cu
// run:   ./tc5_cgemm [M N K reps]      (defaults 256 128 256 for the check, then 4096 4096 512 timed)
//        TC5_LBO / TC5_SBO (bytes) override the descriptor offsets while debugging
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

constexpr int BM = 128, BNC = 64, BKC = 16;           // rows, complex columns, complex k per stage
constexpr int TILE_BYTES = 128 * 32 * 4;              // one operand tile: 128 rows x 32 floats
constexpr int STAGE_BYTES = 4 * TILE_BYTES;           // A_hi, A_lo, B_hi, B_lo
constexpr int STAGES = 2;
constexpr int TMEM_COLS = 256;                        // accumulators: [0,128) hi*hi, [128,256) cross terms
constexpr size_t SMEM_BYTES = (size_t)STAGES * STAGE_BYTES + 64;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ float tf32_rna(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    for (uint32_t spin = 0; spin < (1u << 27); ++spin) {
        uint32_t done;
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (done) return;
    }
    __trap();                                          // never hang: a descriptor the hardware rejects ends up here
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_c, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_c), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
// K-major, no swizzle (cute::UMMA::SmemDescriptor): start >> 4 at [0,14), LBO >> 4 at [16,30), SBO >> 4 at [32,46), version 1 at [46,48)
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((addr & 0x3FFFF) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
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

// element (row r, 16-byte chunk c) of a 128 x 32-float tile in the canonical layout
__device__ __forceinline__ uint32_t tile_off(int r, int c) { return (uint32_t)((r >> 3) * 1024 + c * 128 + (r & 7) * 16); }

__global__ void __launch_bounds__(256, 1)
cgemm_tc5_kernel(const float2* __restrict__ A, const float2* __restrict__ B, float2* __restrict__ C, int M, int N, int K,
                 uint32_t lbo, uint32_t sbo) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BNC;
    uint8_t* ctl = smem + STAGES * STAGE_BYTES;
    const uint32_t bar_free0 = smem_u32(ctl), bar_free1 = smem_u32(ctl + 8), bar_done = smem_u32(ctl + 16);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(ctl + 32);

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid == 0) {
        mbar_init(bar_free0, 1); mbar_init(bar_free1, 1); mbar_init(bar_done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tmem_slot;
    // instruction descriptor (cute::UMMA::InstrDescriptor): D = f32 (1 << 4), A = B = tf32 (2 << 7, 2 << 10), both K-major,
    // N >> 3 at [17,23), M >> 4 at [24,29)
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24);

    const int nchunks = K / BKC;
    for (int c = 0; c < nchunks; ++c) {
        const int s = c & 1;
        uint8_t* st = smem + s * STAGE_BYTES;
        if (c >= STAGES) mbar_wait(s ? bar_free1 : bar_free0, (uint32_t)(((c - STAGES) >> 1) & 1));   // MMAs of chunk c-2 done reading
        const int k0 = c * BKC;
        // A: 128 rows x 8 chunks of 2 complex; lanes = 8 rows x 4 chunks (64 contiguous bytes per row, conflict-free 16 B stores)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int idx = tid + 256 * i, r = ((idx >> 5) & 15) * 8 + (idx & 7), ch = (idx >> 9) * 4 + ((idx >> 3) & 3);
            const float4 v = __ldg(reinterpret_cast<const float4*>(A + (size_t)(m0 + r) * K + k0 + 2 * ch));
            float4 hi, lo;
            hi.x = tf32_rna(v.x); hi.y = tf32_rna(v.y); hi.z = tf32_rna(v.z); hi.w = tf32_rna(v.w);
            lo.x = tf32_rna(v.x - hi.x); lo.y = tf32_rna(v.y - hi.y); lo.z = tf32_rna(v.z - hi.z); lo.w = tf32_rna(v.w - hi.w);
            *reinterpret_cast<float4*>(st + tile_off(r, ch)) = hi;
            *reinterpret_cast<float4*>(st + TILE_BYTES + tile_off(r, ch)) = lo;
        }
        // B: 64 complex rows x 8 chunks, each feeding row n (re, -im) and row 64 + n (im, re)
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int idx = tid + 256 * i, r = ((idx >> 5) & 7) * 8 + (idx & 7), ch = (idx >> 8) * 4 + ((idx >> 3) & 3);
            const float4 v = __ldg(reinterpret_cast<const float4*>(B + (size_t)(n0 + r) * K + k0 + 2 * ch));
            const float4 re = make_float4(v.x, -v.y, v.z, -v.w), im = make_float4(v.y, v.x, v.w, v.z);
            float4 hi, lo;
            hi.x = tf32_rna(re.x); hi.y = tf32_rna(re.y); hi.z = tf32_rna(re.z); hi.w = tf32_rna(re.w);
            lo.x = tf32_rna(re.x - hi.x); lo.y = tf32_rna(re.y - hi.y); lo.z = tf32_rna(re.z - hi.z); lo.w = tf32_rna(re.w - hi.w);
            *reinterpret_cast<float4*>(st + 2 * TILE_BYTES + tile_off(r, ch)) = hi;
            *reinterpret_cast<float4*>(st + 3 * TILE_BYTES + tile_off(r, ch)) = lo;
            hi.x = tf32_rna(im.x); hi.y = tf32_rna(im.y); hi.z = tf32_rna(im.z); hi.w = tf32_rna(im.w);
            lo.x = tf32_rna(im.x - hi.x); lo.y = tf32_rna(im.y - hi.y); lo.z = tf32_rna(im.z - hi.z); lo.w = tf32_rna(im.w - hi.w);
            *reinterpret_cast<float4*>(st + 2 * TILE_BYTES + tile_off(64 + r, ch)) = hi;
            *reinterpret_cast<float4*>(st + 3 * TILE_BYTES + tile_off(64 + r, ch)) = lo;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> visible to the MMA's async proxy
        __syncthreads();
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t a_hi = smem_u32(st), a_lo = a_hi + TILE_BYTES, b_hi = a_hi + 2 * TILE_BYTES, b_lo = a_hi + 3 * TILE_BYTES;
#pragma unroll
            for (int j = 0; j < 4; ++j) {                  // UMMA K = 8 tf32 = two 16-byte chunks = 256 B further on
                const uint32_t o = (uint32_t)j * 256u;      // tile_off(): two 128-byte core-matrix columns
                const uint32_t acc = (c > 0 || j > 0) ? 1u : 0u;
                umma_tf32(tmem, smem_desc(a_hi + o, lbo, sbo), smem_desc(b_hi + o, lbo, sbo), idesc, acc);
                umma_tf32(tmem + 128, smem_desc(a_hi + o, lbo, sbo), smem_desc(b_lo + o, lbo, sbo), idesc, acc);
                umma_tf32(tmem + 128, smem_desc(a_lo + o, lbo, sbo), smem_desc(b_hi + o, lbo, sbo), idesc, 1u);
            }
            umma_commit(s ? bar_free1 : bar_free0);        // arrives when the MMAs above have consumed this stage
            if (c == nchunks - 1) umma_commit(bar_done);
        }
    }
    mbar_wait(bar_done, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // epilogue: warp w reads TMEM lanes (w % 4) * 32 .. +31 (row m = lane), complex columns (w / 4) * 32 .. +31
    {
        const int q = warp & 3, h = warp >> 2;
        const uint32_t base = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(h * 32);
        uint32_t re[32], im[32], xre[32], xim[32];
        tmem_ld32(base, re); tmem_ld32(base + 64, im); tmem_ld32(base + 128, xre); tmem_ld32(base + 192, xim);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        float2* crow = C + (size_t)(m0 + q * 32 + lane) * N + n0 + h * 32;
#pragma unroll
        for (int i = 0; i < 32; ++i)
            crow[i] = make_float2(__uint_as_float(re[i]) + __uint_as_float(xre[i]), __uint_as_float(im[i]) + __uint_as_float(xim[i]));
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TMEM_COLS));
}

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)

static int run(int M, int N, int K, int reps, bool check, uint32_t lbo, uint32_t sbo) {
    if (M % BM || N % BNC || K % BKC) { fprintf(stderr, "M %% 128, N %% 64, K %% 16 must be 0\n"); return 1; }
    std::vector<float2> hA((size_t)M * K), hB((size_t)N * K), hC((size_t)M * N);
    uint64_t s = 0x9E3779B97F4A7C15ull;
    auto rnd = [&] { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return (float)((double)(s >> 11) / 9007199254740992.0 * 2.0 - 1.0); };
    for (auto& v : hA) v = make_float2(rnd(), rnd());
    for (auto& v : hB) v = make_float2(rnd(), rnd());
    float2 *dA, *dB, *dC;
    CK(cudaMalloc(&dA, hA.size() * 8)); CK(cudaMalloc(&dB, hB.size() * 8)); CK(cudaMalloc(&dC, hC.size() * 8));
    CK(cudaMemcpy(dA, hA.data(), hA.size() * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, hB.data(), hB.size() * 8, cudaMemcpyHostToDevice));
    CK(cudaMemset(dC, 0xFF, hC.size() * 8));
    CK(cudaFuncSetAttribute(cgemm_tc5_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
    dim3 grid(N / BNC, M / BM);
    cgemm_tc5_kernel<<<grid, 256, SMEM_BYTES>>>(dA, dB, dC, M, N, K, lbo, sbo);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    if (check) {
        CK(cudaMemcpy(hC.data(), dC, hC.size() * 8, cudaMemcpyDeviceToHost));
        double worst = 0, scale = 0;
        for (int m = 0; m < M; ++m)
            for (int n = 0; n < N; ++n) {
                double re = 0, im = 0;
                for (int k = 0; k < K; ++k) {
                    const float2 a = hA[(size_t)m * K + k], b = hB[(size_t)n * K + k];
                    re += (double)a.x * b.x - (double)a.y * b.y;
                    im += (double)a.x * b.y + (double)a.y * b.x;
                }
                const float2 c = hC[(size_t)m * N + n];
                worst = std::fmax(worst, std::fmax(std::fabs(c.x - re), std::fabs(c.y - im)));
                scale = std::fmax(scale, std::fmax(std::fabs(re), std::fabs(im)));
            }
        printf("{\"check\": {\"M\": %d, \"N\": %d, \"K\": %d, \"max_abs_err\": %.3e, \"max_abs_ref\": %.3e, \"rel\": %.3e, \"lbo\": %u, \"sbo\": %u}}\n",
               M, N, K, worst, scale, worst / scale, lbo, sbo);
    } else {
        cudaEvent_t e0, e1;
        CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        for (int i = 0; i < 3; ++i) cgemm_tc5_kernel<<<grid, 256, SMEM_BYTES>>>(dA, dB, dC, M, N, K, lbo, sbo);
        CK(cudaEventRecord(e0));
        for (int i = 0; i < reps; ++i) cgemm_tc5_kernel<<<grid, 256, SMEM_BYTES>>>(dA, dB, dC, M, N, K, lbo, sbo);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        const double flops = 8.0 * M * N * (double)K;
        printf("{\"timed\": {\"M\": %d, \"N\": %d, \"K\": %d, \"ms\": %.4f, \"tflops_complex_equivalent\": %.2f, \"tf32_mma_tflops\": %.2f}}\n",
               M, N, K, ms / reps, flops / (ms / reps * 1e-3) / 1e12, 3.0 * flops / (ms / reps * 1e-3) / 1e12);
    }
    cudaFree(dA); cudaFree(dB); cudaFree(dC);
    return 0;
}

int main(int argc, char** argv) {
    const uint32_t lbo = getenv("TC5_LBO") ? (uint32_t)atoi(getenv("TC5_LBO")) : 128u;
    const uint32_t sbo = getenv("TC5_SBO") ? (uint32_t)atoi(getenv("TC5_SBO")) : 1024u;
    if (argc >= 4) return run(atoi(argv[1]), atoi(argv[2]), atoi(argv[3]), argc > 4 ? atoi(argv[4]) : 20, argc <= 4, lbo, sbo);
    if (int rc = run(256, 128, 256, 0, true, lbo, sbo)) return rc;
    return run(4096, 4096, 512, 20, false, lbo, sbo);
}
