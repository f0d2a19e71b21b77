"""CPU replay of the row-program interpreter kernel (csrc/qxb_rowprog.cu) on the descriptors the library builds.

Test infrastructure.  ``qxb_debug_rowprog`` serialises the program of one phase exactly as the executor hands it to
``rowprog_kernel``: dependency levels, warp units, the size-aligned shared-memory arena plan, the 128-byte hot
descriptors (register-tile / K tables, added to the lane bases) and the per-unit lane tables (base offsets of each lane's
thread-tile, evaluated on the host from the thread-tile -> address segments).  This
module mirrors the structs (csrc/qxb_rowprog.h) and walks the same loops as the kernel -- level, unit, lane,
thread-tile, register tile, K chunk, k; the K-splitting lanes of the small-op path -- on a numpy arena per bitstring
row, with the const phase and the leaves taken from ``lowered_emulator``.  Levels are executed with a barrier
in between (all units of a level read the state the previous level left), so a missing dependency or an arena
overlap between tensors that are live in the same level shows up as a wrong amplitude.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

import lowered_emulator as le

K_MAX_SEG, K_MAX_KSEG = 12, 8


class RSeg(C.Structure):
    _fields_ = [("src", C.c_ubyte), ("dst", C.c_ubyte), ("len", C.c_ubyte), ("pad", C.c_ubyte)]


class RowOpHot(C.Structure):
    _fields_ = [("aT", C.c_uint16 * 4), ("bT", C.c_uint16 * 4), ("cT", C.c_uint16 * 16),
                ("ktA", C.c_uint16 * 16), ("ktB", C.c_uint16 * 16),
                ("nK", C.c_uint8), ("kc", C.c_uint8), ("ma", C.c_uint8), ("nb", C.c_uint8),
                ("ntt", C.c_uint8), ("kind", C.c_uint8), ("ks", C.c_uint8), ("gen", C.c_uint8),
                ("pad_", C.c_uint8 * 8)]          # alignas(16)


class RowUnitDesc(C.Structure):
    _fields_ = [("hot", RowOpHot),
                ("gA", C.c_uint64), ("gB", C.c_uint64), ("gC", C.c_uint64),
                ("kA", RSeg * K_MAX_KSEG), ("kB", RSeg * K_MAX_KSEG),
                ("nkA", C.c_uint8), ("nkB", C.c_uint8), ("pad", C.c_uint8 * 6),
                ("lA", C.c_uint16 * 32), ("lB", C.c_uint16 * 32), ("lC", C.c_uint16 * 32)]


class RowLeaf(C.Structure):
    _fields_ = [("off", C.c_int32), ("span_bits", C.c_int32), ("out_idx", C.c_int32)]


assert C.sizeof(RowOpHot) == 128 and C.sizeof(RowUnitDesc) == 416
K_NULL, K_WARPS = 0xFFFF, 8


class RowProgram:
    pass


class Unavailable(Exception):
    """The library builds no row program for this phase (it then runs the per-op kernels); str = its reason."""


def dump(graph, free_mask: int, phase: int):
    """Row program of ``phase`` (1 block, 2 chunk) for the batched variables in ``free_mask``; None when the library
    would not build one (reason in ``qxb_last_error``)."""
    lib = graph._lib
    need = lib.qxb_debug_rowprog(graph._h, free_mask, phase, None, 0)
    assert need >= 32, need
    buf = (C.c_char * need)()
    assert lib.qxb_debug_rowprog(graph._h, free_mask, phase, buf, need) == need
    raw = bytes(buf)
    hdr = np.frombuffer(raw, dtype=np.int32, count=8)
    if not hdr[0]:
        return None
    rp = RowProgram()
    _, n_ops, n_descs, n_levels, n_leaves, rp.arena_elems, rp.root_off, rp.root_span = (int(x) for x in hdr)
    off = 32
    n_slots = int(np.frombuffer(raw, dtype=np.int32, count=1, offset=off)[0]); off += 4
    rp.level_start = np.frombuffer(raw, dtype=np.int32, count=n_levels + 1, offset=off).tolist(); off += 4 * (n_levels + 1)
    rp.slots = np.frombuffer(raw, dtype=np.uint16, count=n_slots, offset=off).tolist(); off += 2 * (n_slots + n_slots % 2)
    rp.descs = [RowUnitDesc.from_buffer_copy(raw, off + i * C.sizeof(RowUnitDesc)) for i in range(n_descs)]
    off += n_descs * C.sizeof(RowUnitDesc)
    rp.desc_op = np.frombuffer(raw, dtype=np.int32, count=n_descs, offset=off).tolist(); off += 4 * n_descs
    rp.leaves = [RowLeaf.from_buffer_copy(raw, off + i * 12) for i in range(n_leaves)]; off += 12 * n_leaves
    names = ("lop", "ref_a", "ref_b", "ref_c", "in_arena_a", "in_arena_b", "in_arena_c")
    for nm in names:
        setattr(rp, nm, np.frombuffer(raw, dtype=np.int32, count=n_ops, offset=off).tolist()); off += 4 * n_ops
    rp.fused_names = raw[off:need].decode().split() if phase == 3 else []
    if phase != 3:
        assert off == need
    rp.n_levels = n_levels
    assert rp.level_start[-1] == n_slots and all(x % K_WARPS == 0 for x in rp.level_start)
    return rp


def _rseg(segs, n, x):
    r = np.zeros_like(x)
    for i in range(n):
        s = segs[i]
        r |= ((x >> s.src) & ((1 << s.len) - 1)) << s.dst
    return r


class _Mem:
    """One tensor's storage as the kernel sees it: a flat array + base element offset."""
    def __init__(self, buf, base):
        self.buf, self.base = buf, base

    def ld(self, idx):
        return self.buf[self.base + idx]

    def st(self, idx, val):
        self.buf[self.base + idx] = val


def _run_unit(op: RowUnitDesc, mA: _Mem, mB: _Mem, mC: _Mem, fixA, fixB, stores, dtype):
    """One warp unit as rowprog_kernel runs it from its descriptor; stores are appended to ``stores`` (applied at the
    level barrier).  fixA / fixB: what the executor XORs into the lane bases of arena leaves with fixed variables."""
    h = op.hot
    lane = np.arange(32, dtype=np.int64)
    nK = h.nK
    if h.kind == 255:                 # staged operand: 2^ntt elements from the tensor's base (+ chunk offset) into the arena
        n = 1 << h.ntt
        es = 16 if dtype == np.complex128 else 8
        first = op.gA // es           # the dump leaves the pointer null: gA = byte offset of this unit's first element
        dst = op.lC[0] | (op.lC[1] << 16)
        i = np.arange(n, dtype=np.int64)
        stores.append((mC, dst + i, mA.ld(first + i)))
        return
    bA = np.array(op.lA[:], dtype=np.int64) + fixA
    bB = np.array(op.lB[:], dtype=np.int64) + fixB
    bC = np.array(op.lC[:], dtype=np.int64)
    act = bC != K_NULL

    def koff(k):                      # k: int array
        ka = np.array([h.ktA[int(x) & 15] for x in k], dtype=np.int64)
        kb = np.array([h.ktB[int(x) & 15] for x in k], dtype=np.int64)
        if nK > 4:
            ka += _rseg(op.kA, op.nkA, k >> 4)
            kb += _rseg(op.kB, op.nkB, k >> 4)
        return ka, kb

    if h.kind == 254:                 # DMMA: 8 x 8 output tiles, lane = 4 g + t holds A[g][t], B[t][g], C[g][2t], C[g][2t + 1]
        assert dtype == np.complex128, "the FP64 tensor-pipe path exists for ComplexF64 only"
        g_, t_ = lane >> 2, lane & 3
        for i in range(1 << h.ma):
            Ct = np.zeros((8, 8), dtype=dtype)
            for s_ in range(1 << (nK - 2)):
                ka, kb = koff(np.full(32, s_ << 2, dtype=np.int64))
                a = mA.ld(bA + h.aT[i] + ka)
                b = mB.ld(bB + h.bT[i] + kb)
                Af = np.zeros((8, 4), dtype=dtype); Bf = np.zeros((4, 8), dtype=dtype)
                Af[g_, t_] = a
                Bf[t_, g_] = b
                Ct += Af @ Bf
            stores.append((mC, bC + h.cT[i], Ct[g_, 2 * t_]))
            stores.append((mC, bC + h.cT[i] + h.cT[4], Ct[g_, 2 * t_ + 1]))
        return
    if h.kind == 0:                   # kred: lanes split K
        ntt, ks = h.ntt, h.ks
        ksub = lane >> ntt
        assert np.array_equal(act, ksub < (1 << ks)), "kred lane mask"
        acc = np.zeros(32, dtype=dtype)
        for kl in range(1 << (nK - ks)):
            k = ksub | (kl << ks)
            ka, kb = koff(k)
            a = mA.ld(np.where(act, bA + ka, 0))
            b = mB.ld(np.where(act, bB + kb, 0))
            acc += np.where(act, a * b, 0)
        for i in range(ks):
            acc = acc + acc[lane ^ (1 << (ntt + i))]
        sel = act & (ksub == 0)
        stores.append((mC, bC[sel], acc[sel]))
        return
    ma, nb, kc = h.ma, h.nb, h.kc
    assert h.kind == 1 + (((ma * 3 + nb) * 3 + kc) * 2 + 0), "tile kind does not encode (ma, nb, kc)"
    regs = (4 if dtype == np.complex128 else 2) * (((1 << ma) + (1 << nb)) * (1 << kc) + (1 << (ma + nb)))
    assert regs <= 100, "tile variant not instantiated in the kernel"
    bA = np.where(act, bA, 0); bB = np.where(act, bB, 0)
    TM, TN, KK = 1 << ma, 1 << nb, 1 << kc
    acc = np.zeros((TM, TN, 32), dtype=dtype)
    for ch in range(1 << (nK - kc)):
        kb0 = ch << kc
        ka, kbo = koff(np.full(32, kb0, dtype=np.int64))
        for k in range(KK):
            av = [mA.ld(bA + h.aT[j] + ka + h.ktA[k]) for j in range(TM)]
            bv = [mB.ld(bB + h.bT[j] + kbo + h.ktB[k]) for j in range(TN)]
            for jm in range(TM):
                for jn in range(TN):
                    acc[jm, jn] += av[jm] * bv[jn]
    for jm in range(TM):
        for jn in range(TN):
            stores.append((mC, (bC + h.cT[jm * TN + jn])[act], acc[jm, jn][act]))


def run_program(rp: RowProgram, resolve, arena, dtype, fix=None, check_races=True):
    """Execute the levels of ``rp`` the way the kernel walks them: per level, warp w takes slots level_start + w,
    + 8, ...; a barrier after each level.  resolve(tensor index) -> _Mem of a tensor that is not in the row arena.
    fix: {op index: (xorA, xorB)} for arena leaves whose fixed-variable offsets the executor folds in per launch."""
    fix = fix or {}
    for lv in range(rp.n_levels):
        stores = []
        s0, s1 = rp.level_start[lv], rp.level_start[lv + 1]
        order = [s for w in range(K_WARPS) for s in range(s0 + w, s1, K_WARPS)]
        assert sorted(order) == list(range(s0, s1))
        for sl in order:
            di = rp.slots[sl]
            if di == K_NULL:
                continue
            op = rp.descs[di]
            j = rp.desc_op[di]
            if rp.lop[j] < 0:             # copy pseudo-op: source = the whole tensor from its base
                mA, mB, mC = resolve(rp.ref_a[j], base=True), None, _Mem(arena, 0)
            else:
                mA = _Mem(arena, 0) if rp.in_arena_a[j] else resolve(rp.ref_a[j])
                mB = _Mem(arena, 0) if rp.in_arena_b[j] else resolve(rp.ref_b[j])
                mC = _Mem(arena, 0) if rp.in_arena_c[j] else resolve(rp.ref_c[j])
            assert op.hot.gen == (0 if (rp.in_arena_a[j] and rp.in_arena_b[j] and rp.in_arena_c[j]) else 1)
            fa, fb = fix.get(j, (0, 0))
            _run_unit(op, mA, mB, mC, fa, fb, stores, dtype)
        # barrier: the stores of the level land now (no unit of the level may read what another one writes)
        seen = {}
        for mem, idx, val in stores:
            if check_races:
                key = id(mem.buf)
                s = seen.setdefault(key, set())
                addrs = (mem.base + idx).tolist()
                assert not (s & set(addrs)), "two stores of one level hit the same element"
                s.update(addrs)
            mem.st(idx, val)


def run_block_rows(graph, desc, free_mask, data, bits, fixed_vals, dtype=np.complex128):
    """One aligned block the way the row-program path executes it: const phase as lowered ops (folded at compile
    time by the per-op kernels), block phase = single-CTA row program over the global arenas, chunk phase = one
    shared-memory arena per bitstring row; returns the per-bitstring partial sums."""
    n = bits.shape[0]
    T = desc["tensors"]
    ops = desc["ops"]
    ar = desc["arena_elems"]
    arenas = {"const": np.zeros(max(ar["const"], 2), dtype=dtype), "block": np.zeros(max(ar["block"], 2), dtype=dtype)}
    leaves = {}

    def locate(t):
        off = sum(int(fixed_vals[v]) << pos for v, pos in t["fixed"])
        if t["leaf"] and not t["output_leaf"]:
            lab = t["data_label"]
            if lab not in leaves:
                leaves[lab] = le.pad_leaf(data[lab]).astype(dtype)
            return leaves[lab], off
        assert t["phase"] != "chunk"
        return arenas[t["phase"]], t["offset"] + off

    for op in ops:                                      # const phase (per-op kernels at compile time)
        if op["phase"] != "const":
            continue
        A, a0 = locate(T[op["a"]]); B, b0 = locate(T[op["b"]]); Cb, c0 = locate(T[op["c"]])
        c = np.arange(1 << op["nC"], dtype=np.int64)
        fa, fb = le.segeval(op["segA"], c), le.segeval(op["segB"], c)
        ks = np.arange(1 << op["nK"], dtype=np.int64)
        ka, kb = le.segeval(op["segKA"], ks), le.segeval(op["segKB"], ks)
        acc = np.zeros(1 << op["nC"], dtype=dtype)
        for k in range(len(ks)):
            acc += A[a0 + fa + ka[k]] * B[b0 + fb + kb[k]]
        Cb[c0 + c] = acc

    def resolve(ti, base=False):
        buf, off = locate({**T[ti], "fixed": []} if base else T[ti])
        return _Mem(buf, off)

    rp_block = dump(graph, free_mask, 1)
    if any(o["phase"] == "block" for o in ops):
        if rp_block is None:
            raise Unavailable(graph._lib.qxb_last_error().decode())
        run_program(rp_block, resolve, None, dtype)
    rp = dump(graph, free_mask, 2)
    if rp is None:
        raise Unavailable(graph._lib.qxb_last_error().decode())
    out = np.zeros(n, dtype=np.complex128)
    for u in range(n):
        arena = np.full(max(rp.arena_elems, 1), np.nan + 0j, dtype=dtype)      # NaN: reading an unwritten element shows
        for lf in rp.leaves:
            val = bits[u, lf.out_idx - 1]
            v = np.zeros(1 << lf.span_bits, dtype=dtype)
            if val == 0: v[0] = 1
            elif val == 1: v[1] = 1
            elif val == 2: v[:2] = 1
            else: v[0], v[1] = 1, -1
            arena[lf.off: lf.off + (1 << lf.span_bits)] = v
        # fixed-variable offsets of arena-resident output leaves are XORed into the lane bases per launch by the executor
        fix = {}
        for j in range(len(rp.ref_a)):
            if rp.lop[j] < 0:
                continue
            fa = sum(int(fixed_vals[v]) << pos for v, pos in T[rp.ref_a[j]]["fixed"]) if rp.in_arena_a[j] else 0
            fb = sum(int(fixed_vals[v]) << pos for v, pos in T[rp.ref_b[j]]["fixed"]) if rp.in_arena_b[j] else 0
            if fa or fb:
                fix[j] = (fa, fb)
        run_program(rp, resolve, arena, dtype, fix)
        root = arena[rp.root_off: rp.root_off + (1 << rp.root_span)].astype(np.complex128)
        out[u] = desc["root_scale"] * np.sum(root)
    return out


def amplitudes(graph, data, bits, slice_begin=0, slice_end=None, dtype=np.complex128):
    dims = graph.slice_dims
    if slice_end is None:
        slice_end = graph.n_slices
    total = np.zeros(bits.shape[0], dtype=np.complex128)
    for n_free, vals in le.decompose(dims, slice_begin, slice_end):
        mask = (1 << n_free) - 1
        total = total + run_block_rows(graph, graph.describe(n_free), mask, data, bits, vals, dtype)
    return total.astype(dtype)
