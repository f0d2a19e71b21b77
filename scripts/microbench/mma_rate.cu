// Micro-benchmark: issue rate of the warp-level MMA instructions the tensor-core GEMM kernels use
// (mma.sync m16n8k8 tf32 -> HMMA.1688.F32.TF32, mma.sync m8n8k4 f64 -> DMMA.8) and of FFMA / DFMA,
// as a function of independent accumulator chains per warp (ILP) and warps per SM.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate.bin mma_rate.cu ; run on the GPU box.
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void k_tf32(float* out, int iters) {
    float d[ILP][4];
    unsigned a[4] = {threadIdx.x, threadIdx.x + 1, threadIdx.x + 2, threadIdx.x + 3}, b[2] = {threadIdx.x * 3, threadIdx.x * 5};
#pragma unroll
    for (int i = 0; i < ILP; ++i) for (int e = 0; e < 4; ++e) d[i][e] = 0.f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i)
            asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                         : "+f"(d[i][0]), "+f"(d[i][1]), "+f"(d[i][2]), "+f"(d[i][3])
                         : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) for (int e = 0; e < 4; ++e) s += d[i][e];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ILP>
__global__ void k_f64(double* out, int iters) {
    double d[ILP][2];
    double a = threadIdx.x * 1e-3, b = threadIdx.x * 2e-3;
#pragma unroll
    for (int i = 0; i < ILP; ++i) d[i][0] = d[i][1] = 0.0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                         : "+d"(d[i][0]), "+d"(d[i][1]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += d[i][0] + d[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ILP, typename T>
__global__ void k_fma(T* out, int iters) {
    T d[ILP];
    T a = (T)(threadIdx.x * 1e-3), b = (T)1.0001;
#pragma unroll
    for (int i = 0; i < ILP; ++i) d[i] = (T)i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) d[i] = fma(d[i], b, a);
    }
    T s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += d[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
float time_ms(F launch) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    launch(); cudaDeviceSynchronize();
    cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    return ms;
}

int main() {
    cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
    const int sms = pr.multiProcessorCount;
    int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_khz\": %d, \"rows\": [\n", pr.name, sms, khz);
    void* buf; cudaMalloc(&buf, 1 << 26);
    const int iters = 4096;
    bool first = true;
    auto row = [&](const char* what, int ilp, int warps, float ms, double flops_per_instr) {
        const double instr = (double)sms * warps * ilp * iters;
        const double per_sm_clk = ms * 1e-3 * khz * 1e3 / ((double)warps * ilp * iters);     // clocks per instruction per SM
        printf("%s {\"op\": \"%s\", \"ilp\": %d, \"warps_per_sm\": %d, \"ms\": %.4f, \"tflops\": %.2f, \"clk_per_instr_per_sm\": %.3f}",
               first ? "" : ",\n", what, ilp, warps, ms, instr * flops_per_instr / (ms * 1e-3) / 1e12, per_sm_clk);
        first = false;
    };
#define RUN_TF32(ILP, W) row("hmma_tf32_m16n8k8", ILP, W, time_ms([&] { k_tf32<ILP><<<sms, 32 * W>>>((float*)buf, iters); }), 2.0 * 16 * 8 * 8)
#define RUN_F64(ILP, W) row("dmma_m8n8k4", ILP, W, time_ms([&] { k_f64<ILP><<<sms, 32 * W>>>((double*)buf, iters); }), 2.0 * 8 * 8 * 4)
#define RUN_FFMA(ILP, W) row("ffma", ILP, W, time_ms([&] { k_fma<ILP, float><<<sms, 32 * W>>>((float*)buf, iters); }), 2.0 * 32)
#define RUN_DFMA(ILP, W) row("dfma", ILP, W, time_ms([&] { k_fma<ILP, double><<<sms, 32 * W>>>((double*)buf, iters); }), 2.0 * 32)
    RUN_TF32(1, 4); RUN_TF32(2, 4); RUN_TF32(4, 4); RUN_TF32(8, 4); RUN_TF32(16, 4);
    RUN_TF32(4, 8); RUN_TF32(8, 8); RUN_TF32(16, 8); RUN_TF32(8, 16); RUN_TF32(16, 16); RUN_TF32(8, 32);
    RUN_F64(1, 4); RUN_F64(2, 4); RUN_F64(4, 4); RUN_F64(8, 4); RUN_F64(16, 4);
    RUN_F64(4, 8); RUN_F64(8, 8); RUN_F64(8, 16); RUN_F64(16, 16); RUN_F64(8, 32);
    RUN_FFMA(8, 16); RUN_FFMA(8, 32); RUN_FFMA(16, 32);
    RUN_DFMA(8, 16); RUN_DFMA(8, 32); RUN_DFMA(16, 32);
    printf("\n]}\n");
    return 0;
}
