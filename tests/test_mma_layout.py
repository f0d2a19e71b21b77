"""Index arithmetic of the tensor-core GEMM kernels, checked on the CPU (tests/mma_layout_emulator.py):
shared-memory staging tables (from the library), fragment loads, mma.sync fragment layouts, epilogue."""
import ctypes as C

import numpy as np
import pytest

from qxb200 import _lib
import mma_layout_emulator as em


def lib_smem_bit():
    lib = _lib.load()
    f = lib.qxb_debug_mma_smem_bit
    f.restype = C.c_int
    f.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int]
    return lambda dtype, is_b, tb, b: f(dtype, 1 if is_b else 0, tb, b)


def test_smem_tables_match_library():
    f = lib_smem_bit()
    for dtype in (0, 1):
        kcb = 4 if dtype == 0 else 3
        for is_b in (False, True):
            for b in range(6 + kcb):
                assert f(dtype, is_b, 6, b) == em.smem_bit(dtype, is_b, 6, b)


@pytest.mark.parametrize("dtype", [0, 1])
@pytest.mark.parametrize("seed", [0, 1, 2])
def test_tile_product(dtype, seed):
    rng = np.random.default_rng(seed)
    kcb = 4 if dtype == 0 else 3
    BK = 1 << kcb
    A = rng.standard_normal((64, BK)) + 1j * rng.standard_normal((64, BK))
    B = rng.standard_normal((64, BK)) + 1j * rng.standard_normal((64, BK))
    if dtype == 0:
        A = A.astype(np.complex64); B = B.astype(np.complex64)
    permA = list(rng.permutation(6 + kcb)); permB = list(rng.permutation(6 + kcb))
    f = lib_smem_bit()
    sA = em.stage(A, dtype, False, permA, f)
    sB = em.stage(B, dtype, True, permB, f)
    got = em.run_c32(sA, sB) if dtype == 0 else em.run_c64(sA, sB)
    want = A.astype(np.complex128) @ B.astype(np.complex128).T
    tol = 2e-6 if dtype == 0 else 1e-13       # 3xTF32: dropped lo*lo terms ~ 2^-22 per product
    assert np.max(np.abs(got - want)) / np.max(np.abs(want)) < tol


def test_tc5_staging_layout(lib_built):
    """tcgen05 kernel (csrc/qxb_gemm_tc5.cu): the per-bit byte offsets the gather adds up must reproduce the canonical
    K-major no-swizzle core-matrix layout the UMMA descriptors announce (LBO 128 B, SBO 1024 B): element (row r,
    complex k c) of a 2^tb-row tile at (r >> 3) * 1024 + (c >> 1) * 128 + (r & 7) * 16 + (c & 1) * 8."""
    lib = lib_built
    for tb in (7, 6):
        seen = set()
        for r in range(1 << tb):
            for c in range(16):
                idx = r | (c << tb)
                off = sum(lib.qxb_debug_tc5_smem_bit(tb, b) for b in range(tb + 4) if (idx >> b) & 1)
                assert off == (r >> 3) * 1024 + (c >> 1) * 128 + (r & 7) * 16 + (c & 1) * 8
                seen.add(off)
        assert len(seen) == (1 << tb) * 16 and max(seen) + 8 <= 128 * 32 * 4      # distinct, inside one 16 KB plane
    # B: rows n and 64 + n of the 128-row real operand; 64 rows further = 8 groups of 8 rows = 8 * 1024 bytes
    assert lib.qxb_debug_tc5_smem_bit(7, 6) == 8 * 1024
