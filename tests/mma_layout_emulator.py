"""CPU emulation of the data movement of the tensor-core GEMM kernels
(qxb_kernels.cu: gemm_tf32x3_kernel / gemm_dmma_kernel) for ONE 64 x 64 block tile and one K chunk:
staging through the host-built shared-memory tables, per-lane fragment loads, the PTX-documented
mma.sync fragment layouts, and the epilogue's accumulator -> (row, col) map.  It checks the index
arithmetic (the part that cannot be seen by the compiler), not the MMA instruction itself.
Test infrastructure only."""
import numpy as np

TMB = TNB = 6
BM = BN = 64


def smem_bit(dtype, is_b, tb, b):
    """mirror of gemm_mma_smem_bit (checked against the library by test_mma_layout.py)"""
    if dtype != 0:
        return (1 << b) if b < tb else (1 << (b - tb)) * ((1 << tb) + 2)
    c = b - tb
    if not is_b:
        if b < 3: return 16 << b
        if b == 3: return 1
        if b < tb: return 512 << (b - 4)
        if c < 2: return 4 << c
        if c == 2: return 2
        return ((1 << tb) // 16) * 512
    if b < 3: return 16 << b
    if b < tb: return 256 << (b - 3)
    if c < 2: return 4 << c
    if c == 2: return 1
    return ((1 << tb) // 8) * 256


def stage(tile, dtype, is_b, perm, smem_bit_fn=smem_bit):
    """tile[m][k] complex; perm = order of the tile-index bits over the load-slot bits (the host sorts
    them by global offset).  Returns the shared-memory image the kernel would build."""
    kcb = 4 if dtype == 0 else 3
    tb = 6
    nbits = tb + kcb
    BK = 1 << kcb
    if dtype == 0:
        sm = np.full(((BK // 8) * (64 // (8 if is_b else 16)) * (256 if is_b else 512),), np.nan, np.float32)
    else:
        sm = np.full((BK * 66,), np.nan, np.complex128)
    for e in range(1 << nbits):            # slot e = tid + i * NT: the split into tid / i does not matter here
        idx, off = 0, 0
        for j in range(nbits):
            if (e >> j) & 1:
                idx |= 1 << perm[j]
                off += smem_bit_fn(dtype, is_b, tb, perm[j])
        m, k = idx & 63, idx >> tb
        v = tile[m, k]
        if dtype == 0:
            off = frag_swz(off)
            re, im = np.float32(v.real), np.float32(v.imag)
            hr, hi = split_hi(re), split_hi(im)
            lr, li = split_hi(np.float32(re - hr)), split_hi(np.float32(im - hi))
            if is_b:
                sm[off + 0], sm[off + 2], sm[off + 128], sm[off + 130] = hr, hi, lr, li
            else:
                sm[off + 0], sm[off + 128], sm[off + 256], sm[off + 384] = hr, hi, lr, li
        else:
            sm[off] = v
    assert not np.isnan(sm.view(np.float32 if dtype == 0 else np.float64)).any() or dtype != 0
    return sm


def frag_swz(off):
    """bank swizzle of the fragment-major c32 tiles (qxb_kernels.cu: frag_swz)"""
    return off ^ (((off >> 5) & 3) << 2)


def split_hi(x):
    b = np.array([x], np.float32).view(np.uint32)
    b = (b + np.uint32(0x1000)) & np.uint32(0xffffe000)
    return b.view(np.float32)[0]


def mma_m16n8k8(d, a, b0, b1):
    """d[lane][4] += A(16x8) B(8x8); a[lane][4], b0/b1[lane] -- PTX ISA fragment layout for .tf32"""
    A = np.zeros((16, 8)); B = np.zeros((8, 8))
    for lane in range(32):
        g, t = lane >> 2, lane & 3
        A[g, t], A[g + 8, t], A[g, t + 4], A[g + 8, t + 4] = a[lane]
        B[t, g], B[t + 4, g] = b0[lane], b1[lane]
    D = A @ B
    for lane in range(32):
        g, t = lane >> 2, lane & 3
        d[lane] += [D[g, 2 * t], D[g, 2 * t + 1], D[g + 8, 2 * t], D[g + 8, 2 * t + 1]]


def mma_m8n8k4(d, a, b):
    A = np.zeros((8, 4)); B = np.zeros((4, 8))
    for lane in range(32):
        g, t = lane >> 2, lane & 3
        A[g, t] = a[lane]; B[t, g] = b[lane]
    D = A @ B
    for lane in range(32):
        g, t = lane >> 2, lane & 3
        d[lane] += [D[g, 2 * t], D[g, 2 * t + 1]]


def run_c32(sA, sB):
    """4 warps, warp tile 32 x 32; returns C[64][64] complex from the emulated fragments"""
    MTC, NTC = 4, 8
    C = np.zeros((64, 64), np.complex128)
    fA = sA.reshape(-1, 4); fB = sB.reshape(-1, 4)       # float4 views
    lanes = np.arange(32)
    for warp in range(4):
        wm0, wn0 = (warp % 2) * 32, (warp // 2) * 32
        cre = np.zeros((2, 4, 32, 4)); cim = np.zeros((2, 4, 32, 4))
        for ks in range(2):
            lsw = lanes ^ (lanes >> 3)
            baseA = (wm0 // 16) * 128 + lsw
            baseB = (wn0 // 8) * 64 + lsw
            for nt in range(4):
                q = baseB + (ks * NTC + nt) * 64
                bh, bl = fB[q], fB[q + 32]
                for mt in range(2):
                    qa = baseA + (ks * MTC + mt) * 128
                    ahr, ahi, alr, ali = fA[qa], fA[qa + 32], fA[qa + 64], fA[qa + 96]
                    r, i = cre[mt, nt], cim[mt, nt]
                    mma_m16n8k8(r, alr, bh[:, 0], bh[:, 1]); mma_m16n8k8(i, alr, bh[:, 2], bh[:, 3])
                    mma_m16n8k8(r, ahr, bl[:, 0], bl[:, 1]); mma_m16n8k8(i, ahr, bl[:, 2], bl[:, 3])
                    mma_m16n8k8(r, ali, -bh[:, 2], -bh[:, 3]); mma_m16n8k8(i, ali, bh[:, 0], bh[:, 1])
                    mma_m16n8k8(r, ahi, -bl[:, 2], -bl[:, 3]); mma_m16n8k8(i, ahi, bl[:, 0], bl[:, 1])
                    mma_m16n8k8(r, ahr, bh[:, 0], bh[:, 1]); mma_m16n8k8(i, ahr, bh[:, 2], bh[:, 3])
                    mma_m16n8k8(r, ahi, -bh[:, 2], -bh[:, 3]); mma_m16n8k8(i, ahi, bh[:, 0], bh[:, 1])
        for mt in range(2):
            for h in range(2):
                for nt in range(4):
                    for j in range(2):
                        for lane in range(32):
                            g, t = lane >> 2, lane & 3
                            row = wm0 + mt * 16 + g + h * 8
                            col = wn0 + nt * 8 + 2 * t + j
                            C[row, col] = cre[mt, nt, lane, h * 2 + j] + 1j * cim[mt, nt, lane, h * 2 + j]
    return C


def run_c64(sA, sB):
    """8 warps, warp tile 16 x 32"""
    LD = 66
    C = np.zeros((64, 64), np.complex128)
    lanes = np.arange(32)
    g, t = lanes >> 2, lanes & 3
    for warp in range(8):
        wm0, wn0 = (warp % 4) * 16, (warp // 4) * 32
        cre = np.zeros((2, 4, 32, 2)); cim = np.zeros((2, 4, 32, 2))
        for ks in range(2):
            ab = (ks * 4 + t) * LD + wm0 + g
            bb = (ks * 4 + t) * LD + wn0 + g
            for nt in range(4):
                bv = sB[bb + nt * 8]
                for mt in range(2):
                    af = sA[ab + mt * 8]
                    mma_m8n8k4(cre[mt, nt], af.real, bv.real); mma_m8n8k4(cim[mt, nt], af.real, bv.imag)
                    mma_m8n8k4(cre[mt, nt], -af.imag, bv.imag); mma_m8n8k4(cim[mt, nt], af.imag, bv.real)
        for mt in range(2):
            for nt in range(4):
                for j in range(2):
                    for lane in range(32):
                        row = wm0 + mt * 8 + (lane >> 2)
                        col = wn0 + nt * 8 + 2 * (lane & 3) + j
                        C[row, col] = cre[mt, nt, lane, j] + 1j * cim[mt, nt, lane, j]
    return C
