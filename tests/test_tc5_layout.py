"""numpy replay of scripts/microbench/tc5_cgemm.cu (the tcgen05 3xTF32 complex GEMM bring-up kernel, not yet run on
hardware): the staging index arithmetic, the shared-memory image it writes, what a UMMA with the kernel's descriptors
reads from that image under the canonical K-major no-swizzle layout of cute/arch/mma_sm100_desc.hpp
(((8,m),(T,2)) : ((T,SBO),(1,LBO)) in elements, T = 4 for tf32), the real-GEMM formulation of the complex product,
the 3xTF32 split with a separate accumulator for the cross terms, and the TMEM -> C epilogue mapping.
It pins the MATH and the layout bookkeeping; the PTX itself can only be checked on a B200."""
import numpy as np

BM, BNC, BKC = 128, 64, 16
TILE_BYTES = 128 * 32 * 4
LBO, SBO = 128, 1024


def tf32_rna(x):
    """cvt.rna.tf32.f32: round to nearest (ties away) on a 10-bit mantissa, as float32."""
    u = np.asarray(x, dtype=np.float32).view(np.uint32).astype(np.uint64)
    u = (u + 0x1000) & 0xFFFFE000
    return u.astype(np.uint32).view(np.float32)


def tile_off(r, c):
    return (r >> 3) * 1024 + c * 128 + (r & 7) * 16


def stage_image(A, B, m0, n0, k0):
    """The 64 KB stage the 256 threads write for one K chunk: [A_hi | A_lo | B_hi | B_lo], bytes as float32 words."""
    img = np.zeros(4 * TILE_BYTES // 4, dtype=np.float32)

    def put(tile, r, ch, v4):
        o = (tile * TILE_BYTES + tile_off(r, ch)) // 4
        hi = tf32_rna(v4)
        img[o:o + 4] = hi
        lo = tf32_rna(v4 - hi)
        o = ((tile + 1) * TILE_BYTES + tile_off(r, ch)) // 4
        img[o:o + 4] = lo

    seen_a, seen_b = set(), set()
    for tid in range(256):
        for i in range(4):
            idx = tid + 256 * i
            r, ch = ((idx >> 5) & 15) * 8 + (idx & 7), (idx >> 9) * 4 + ((idx >> 3) & 3)
            seen_a.add((r, ch))
            v = A[m0 + r, k0 + 2 * ch: k0 + 2 * ch + 2]                   # float4 = two complex
            put(0, r, ch, np.array([v[0].real, v[0].imag, v[1].real, v[1].imag], dtype=np.float32))
        for i in range(2):
            idx = tid + 256 * i
            r, ch = ((idx >> 5) & 7) * 8 + (idx & 7), (idx >> 8) * 4 + ((idx >> 3) & 3)
            seen_b.add((r, ch))
            v = B[n0 + r, k0 + 2 * ch: k0 + 2 * ch + 2]
            x, y, z, w = v[0].real, v[0].imag, v[1].real, v[1].imag
            put(2, r, ch, np.array([x, -y, z, -w], dtype=np.float32))       # row n     : ( re, -im)
            put(2, 64 + r, ch, np.array([y, x, w, z], dtype=np.float32))    # row 64 + n: ( im,  re)
    assert len(seen_a) == 128 * 8 and len(seen_b) == 64 * 8                 # every (row, 16-byte chunk) exactly once
    return img


def umma_operand(img, tile, j):
    """128 x 8 tf32 operand a K-major no-swizzle descriptor (start = tile base + j * 256, LBO, SBO) addresses:
    element (row, k) at start + (row % 8) * 16 + (row // 8) * SBO + (k // 4) * LBO + (k % 4) * 4 bytes."""
    start = tile * TILE_BYTES + j * 256
    out = np.zeros((128, 8), dtype=np.float32)
    for row in range(128):
        for k in range(8):
            out[row, k] = img[(start + (row % 8) * 16 + (row // 8) * SBO + (k // 4) * LBO + (k % 4) * 4) // 4]
    return out


def test_complex_gemm_through_one_real_tf32_gemm():
    rng = np.random.default_rng(0)
    M, N, K = 128, 64, 48
    A = (rng.uniform(-1, 1, (M, K)) + 1j * rng.uniform(-1, 1, (M, K))).astype(np.complex64)
    B = (rng.uniform(-1, 1, (N, K)) + 1j * rng.uniform(-1, 1, (N, K))).astype(np.complex64)
    acc_main = np.zeros((128, 128), dtype=np.float32)                       # TMEM columns [0, 128)
    acc_cross = np.zeros((128, 128), dtype=np.float32)                      # TMEM columns [128, 256)
    for c in range(K // BKC):
        img = stage_image(A, B, 0, 0, c * BKC)
        for j in range(4):                                                  # UMMA K = 8 per step, D[m][n] += sum_k A[m][k] B[n][k]
            a_hi, a_lo, b_hi, b_lo = (umma_operand(img, t, j) for t in range(4))
            acc_main += a_hi @ b_hi.T
            acc_cross += a_hi @ b_lo.T
            acc_cross += a_lo @ b_hi.T
    # epilogue: warp w reads lanes (w % 4) * 32.., complex columns (w / 4) * 32..: re at column n, im at column 64 + n
    C = np.zeros((M, N), dtype=np.complex64)
    for w in range(8):
        q, h = w & 3, w >> 2
        for lane in range(32):
            m = q * 32 + lane
            cols = np.arange(h * 32, h * 32 + 32)
            C[m, cols] = (acc_main[m, cols] + acc_cross[m, cols]) + 1j * (acc_main[m, 64 + cols] + acc_cross[m, 64 + cols])
    ref = A.astype(np.complex128) @ B.astype(np.complex128).T               # C[m][n] = sum_k A[m][k] B[n][k]
    err = np.max(np.abs(C - ref)) / np.max(np.abs(ref))
    assert err < 2e-6, err                                                  # fp32-level; plain tf32 would be ~1e-3
    # and the split matters: hi * hi alone is tf32 accuracy
    C1 = acc_main[:, :64] + 1j * acc_main[:, 64:]
    assert np.max(np.abs(C1 - ref)) / np.max(np.abs(ref)) > 1e-4


def test_descriptor_fields():
    """smem_desc() / idesc of the kernel against the bit fields of cute::UMMA::SmemDescriptor / InstrDescriptor."""
    addr = 0x12340
    desc = ((addr & 0x3FFFF) >> 4) | ((LBO >> 4) << 16) | ((SBO >> 4) << 32) | (1 << 46)
    assert desc & 0x3FFF == addr >> 4                    # start_address_, bits [0, 14)
    assert (desc >> 16) & 0x3FFF == 8                    # leading_byte_offset_ = 128 B
    assert (desc >> 32) & 0x3FFF == 64                   # stride_byte_offset_ = 1024 B
    assert (desc >> 46) & 3 == 1 and desc >> 61 == 0     # version_ = 1 (Blackwell), layout_type_ = SWIZZLE_NONE
    idesc = (1 << 4) | (2 << 7) | (2 << 10) | ((128 >> 3) << 17) | ((128 >> 4) << 24)
    assert (idesc >> 4) & 3 == 1                         # c_format_ = F32
    assert (idesc >> 7) & 7 == 2 and (idesc >> 10) & 7 == 2      # a_format_ = b_format_ = TF32
    assert (idesc >> 15) & 1 == 0 and (idesc >> 16) & 1 == 0     # both K-major
    assert (idesc >> 17) & 0x3F == 16 and (idesc >> 24) & 0x1F == 8   # N = 128, M = 128
