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

// Generic batched bit-segment contraction.  One thread block covers 2^tb
// consecutive C elements per step (or several smaller tiles when tb < 8); the
// low-bit part of the address maps is computed once per thread, the high-bit
// part once per tile.
template <typename R2>
__global__ void __launch_bounds__(kThreads)
contract_kernel(const __grid_constant__ OpParams p) {
    const R2* __restrict__ A = reinterpret_cast<const R2*>(p.A);
    const R2* __restrict__ B = reinterpret_cast<const R2*>(p.B);
    R2* __restrict__ C = reinterpret_cast<R2*>(p.C);
    const int tid = threadIdx.x;
    const int tb = p.tb;
    const int lob = tb < 8 ? tb : 8;
    const int sub_bits = 8 - lob;
    const int sub = tid >> lob;
    const int ept = tb > 8 ? (1 << (tb - 8)) : 1;
    const unsigned lo0 = tid & ((1u << lob) - 1u);
    long long aLo[kMaxEpt], bLo[kMaxEpt];
#pragma unroll
    for (int j = 0; j < kMaxEpt; ++j) {
        aLo[j] = 0; bLo[j] = 0;
        if (j < ept) {
            const unsigned lo = lo0 + j * kThreads;
            aLo[j] = segeval(p.sA, p.nsA, lo);
            bLo[j] = segeval(p.sB, p.nsB, lo);
        }
    }
    const int hb = p.nC - tb;
    const long long hmask = (1ll << hb) - 1ll;
    const int nk = 1 << p.nK;
    const bool ktab = p.nK <= 4;
    for (long long t0 = ((long long)blockIdx.x << sub_bits); t0 < p.tiles;
         t0 += ((long long)gridDim.x << sub_bits)) {
        const long long tile = t0 + sub;
        if (tile >= p.tiles) continue;
        const long long u = tile >> hb;
        const unsigned long long hi = (unsigned long long)(tile & hmask) << tb;
        const long long aHi = u * p.sUA + segeval(p.sA, p.nsA, hi);
        const long long bHi = u * p.sUB + segeval(p.sB, p.nsB, hi);
        R2* Cp = C + u * p.sUC + (long long)hi;
#pragma unroll
        for (int j = 0; j < kMaxEpt; ++j) {
            if (j < ept) {
                const long long a = aHi + aLo[j];
                const long long b = bHi + bLo[j];
                R2 acc; acc.x = 0; acc.y = 0;
                if (ktab) {
                    for (int k = 0; k < nk; ++k)
                        cmac(acc, __ldg(A + a + p.ktabA[k]), __ldg(B + b + p.ktabB[k]));
                } else {
                    for (int k = 0; k < nk; ++k) {
                        const long long ak = segeval(p.kA, p.nkA, (unsigned long long)k);
                        const long long bk = segeval(p.kB, p.nkB, (unsigned long long)k);
                        cmac(acc, __ldg(A + a + ak), __ldg(B + b + bk));
                    }
                }
                Cp[lo0 + j * kThreads] = acc;
            }
        }
    }
}

void launch_contract(int dtype, const OpParams& p, int grid, cudaStream_t st) {
    if (dtype == 0) contract_kernel<float2><<<grid, kThreads, 0, st>>>(p);
    else contract_kernel<double2><<<grid, kThreads, 0, st>>>(p);
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

void launch_output_leaves(int dtype, void* chunk_base, const OutLeafDesc* d_desc, int n_leaves,
                          const unsigned char* d_bits, int n_outputs, long long amp0, long long n,
                          cudaStream_t st) {
    if (n_leaves == 0 || n == 0) return;
    dim3 grid((unsigned)((n * 2 + 255) / 256 > 64 ? 64 : (n * 2 + 255) / 256), (unsigned)n_leaves);
    if (grid.x == 0) grid.x = 1;
    if (dtype == 0) outleaf_kernel<float2><<<grid, 256, 0, st>>>((float2*)chunk_base, d_desc, d_bits, n_outputs, amp0, n);
    else outleaf_kernel<double2><<<grid, 256, 0, st>>>((double2*)chunk_base, d_desc, d_bits, n_outputs, amp0, n);
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

void launch_reduce_root(int dtype, const void* root, long long sU, int span_bits, long long n,
                        double scale, double* acc, long long amp0, cudaStream_t st) {
    if (n == 0) return;
    long long blocks = (n * 32 + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    if (dtype == 0)
        reduce_root_kernel<float2><<<(unsigned)blocks, 256, 0, st>>>((const float2*)root, sU, span_bits, n, scale, acc, amp0);
    else
        reduce_root_kernel<double2><<<(unsigned)blocks, 256, 0, st>>>((const double2*)root, sU, span_bits, n, scale, acc, amp0);
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

void launch_finalize(int dtype, const double* acc, void* out, long long n, cudaStream_t st) {
    if (n == 0) return;
    long long blocks = (n + 255) / 256;
    if (blocks > 1024) blocks = 1024;
    if (dtype == 0) finalize_kernel<float2><<<(unsigned)blocks, 256, 0, st>>>(acc, (float2*)out, n);
    else finalize_kernel<double2><<<(unsigned)blocks, 256, 0, st>>>(acc, (double2*)out, n);
}

}  // namespace qxb
