"""CPU replay of ``contract_kernel``'s index arithmetic on the launch templates the library builds.

Test infrastructure.  ``qxb_debug_templates`` returns the ``OpParams`` array exactly as
``build_templates`` (csrc/qxb_exec.cu) composes it for the kernel: which C bits are thread bits,
register-tile bits and tile-index ("hi") bits, and the address maps for each part.  This module
mirrors the struct (csrc/qxb_kernels.cuh) and walks the same loops as the kernel
(csrc/qxb_kernels.cu: contract_kernel + tile_compute) -- block, thread, tile, register tile, K chunk,
k -- producing for every (c, k) the element addresses of A, B and C the GPU would touch.  The tests
compare them with the lowered op's own definition ``C[c] = sum_k A[fA(c) + gA(k)] B[fB(c) + gB(k)]``.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

K_MAX_SEG, K_MAX_KSEG, K_KTAB, K_THREADS = 32, 32, 16, 256


class DSeg(C.Structure):
    _fields_ = [("src", C.c_ubyte), ("dst", C.c_ubyte), ("len", C.c_ubyte), ("pad", C.c_ubyte)]


class OpParams(C.Structure):
    _fields_ = [("A", C.c_void_p), ("B", C.c_void_p), ("C", C.c_void_p),
                ("sUA", C.c_longlong), ("sUB", C.c_longlong), ("sUC", C.c_longlong), ("tiles", C.c_longlong),
                ("nC", C.c_int), ("nK", C.c_int), ("U", C.c_int), ("lob", C.c_int), ("ma", C.c_int), ("nb", C.c_int),
                ("kc", C.c_int), ("hb", C.c_int), ("aBits", C.c_int), ("bBits", C.c_int),
                ("nsAlo", C.c_int), ("nsBlo", C.c_int), ("nsClo", C.c_int), ("nsAhi", C.c_int), ("nsBhi", C.c_int),
                ("nsChi", C.c_int), ("nkA", C.c_int), ("nkB", C.c_int),
                ("ktabA", C.c_longlong * K_KTAB), ("ktabB", C.c_longlong * K_KTAB),
                ("aT", C.c_longlong * 4), ("bT", C.c_longlong * 4), ("cT", C.c_longlong * 16),
                ("sAlo", DSeg * 8), ("sBlo", DSeg * 8), ("sClo", DSeg * 8),
                ("sAhi", DSeg * K_MAX_SEG), ("sBhi", DSeg * K_MAX_SEG), ("sChi", DSeg * K_MAX_SEG),
                ("kA", DSeg * K_MAX_KSEG), ("kB", DSeg * K_MAX_KSEG)]


def templates(graph, n_free: int = -1):
    """``[OpParams]`` of every lowered op of ``graph`` (an ``executor.Graph``), pointers unset."""
    lib = graph._lib
    need = lib.qxb_debug_templates(graph._h, n_free, None, 0)
    assert need >= 0 and need % C.sizeof(OpParams) == 0, (need, C.sizeof(OpParams))
    n = need // C.sizeof(OpParams)
    arr = (OpParams * n)()
    assert lib.qxb_debug_templates(graph._h, n_free, arr, need) == need
    return list(arr)


def _segeval(segs, n, x):
    r = np.zeros_like(x)
    for i in range(n):
        s = segs[i]
        r |= ((x >> s.src) & ((1 << s.len) - 1)) << s.dst
    return r


def kernel_addresses(p: OpParams):
    """Element offsets (within one bitstring row) that contract_kernel computes for op ``p``:
    ``(c_addr[n_out], a_addr[n_out, K], b_addr[n_out, K])`` in the kernel's enumeration order --
    tile index hh, thread bits lo, register tile (jm, jn); K = chunk-major, ktab inside a chunk."""
    lob, hb, ma, nb, kc = p.lob, p.hb, p.ma, p.nb, p.kc
    assert 0 <= lob <= 8 and kc <= min(p.nK, 4) and ma <= 2 and nb <= 2
    assert lob + ma + nb + hb == p.nC, "thread + tile + hi bits must cover C exactly"
    TM, TN, KK = 1 << ma, 1 << nb, 1 << kc
    hh = np.arange(1 << hb, dtype=np.int64)[:, None, None, None]
    lo = np.arange(1 << lob, dtype=np.int64)[None, :, None, None]
    jm = np.arange(TM)[None, None, :, None]
    jn = np.arange(TN)[None, None, None, :]
    aT = np.array(list(p.aT), dtype=np.int64)
    bT = np.array(list(p.bT), dtype=np.int64)
    cT = np.array(list(p.cT), dtype=np.int64)
    ap = _segeval(p.sAhi, p.nsAhi, hh) + _segeval(p.sAlo, p.nsAlo, lo)        # Ap of the kernel (row offset apart)
    bp = _segeval(p.sBhi, p.nsBhi, hh) + _segeval(p.sBlo, p.nsBlo, lo)
    cp = _segeval(p.sChi, p.nsChi, hh) + _segeval(p.sClo, p.nsClo, lo)
    c_addr = (cp + cT[jm * TN + jn]).reshape(-1)
    shape = np.broadcast_shapes(ap.shape, jm.shape, jn.shape)
    a_base = np.broadcast_to(ap + aT[jm], shape).reshape(-1)
    b_base = np.broadcast_to(bp + bT[jn], shape).reshape(-1)
    # K: chunks of 2^kc; chunk offset through the kA/kB segments, in-chunk offset through ktab (tile_compute)
    ktabA = np.array(list(p.ktabA), dtype=np.int64)
    ktabB = np.array(list(p.ktabB), dtype=np.int64)
    nch = 1 << (p.nK - kc)
    kb = (np.arange(nch, dtype=np.int64) << kc)
    ka = (_segeval(p.kA, p.nkA, kb)[:, None] + ktabA[None, :KK]).reshape(-1)
    kbb = (_segeval(p.kB, p.nkB, kb)[:, None] + ktabB[None, :KK]).reshape(-1)
    return c_addr, a_base[:, None] + ka[None, :], b_base[:, None] + kbb[None, :]


def reference_addresses(op: dict, c_addr):
    """The lowered op's definition (qxb_graph_describe): for C element ``c`` and k, A at fA(c) + gA(k), B at fB(c) + gB(k)."""
    def seg(segs, x):
        r = np.zeros_like(x)
        for src, dst, ln in segs:
            r |= ((x >> src) & ((1 << ln) - 1)) << dst
        return r
    k = np.arange(1 << op["nK"], dtype=np.int64)
    a = seg(op["segA"], c_addr)[:, None] + seg(op["segKA"], k)[None, :]
    b = seg(op["segB"], c_addr)[:, None] + seg(op["segKB"], k)[None, :]
    return a, b


def loads_per_output(p: OpParams):
    """Operand loads the kernel issues per C element: ((2^ma + 2^nb) * K) / 2^(ma+nb)."""
    return ((1 << p.ma) + (1 << p.nb)) * (1 << p.nK) / (1 << (p.ma + p.nb))
