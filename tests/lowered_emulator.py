"""numpy emulator of the LOWERED program that ``qxb_graph_describe`` reports.

Test infrastructure: it executes, on the CPU, exactly what the CUDA executor is
told to do -- the bit-segment address maps, the arena offsets of the memory plan,
the per-phase op order, the aligned-block decomposition of a slice range -- so
the host-side lowering can be checked against the oracle without a GPU.
"""
from __future__ import annotations

import numpy as np


def ceil_log2(x):
    b = 0
    while (1 << b) < x:
        b += 1
    return b


def segeval(segs, x):
    r = np.zeros_like(x)
    for src, dst, ln in segs:
        r |= ((x >> src) & ((1 << ln) - 1)) << dst
    return r


def pad_leaf(arr):
    """Column-major leaf -> power-of-two padded flat buffer (mode m at bits [pos_m, pos_m+nbits_m))."""
    arr = np.asarray(arr, dtype=np.complex128)
    dims = arr.shape
    nbits = [ceil_log2(d) for d in dims]
    pos = np.cumsum([0] + nbits[:-1]) if dims else []
    buf = np.zeros(1 << int(sum(nbits)), dtype=np.complex128)
    if arr.ndim == 0:
        buf[0] = arr
        return buf
    idx = np.indices(dims).reshape(len(dims), -1)
    addr = np.zeros(idx.shape[1], dtype=np.int64)
    for m in range(len(dims)):
        addr |= idx[m].astype(np.int64) << int(pos[m])
    buf[addr] = arr.reshape(-1)        # C-order flatten matches np.indices C-order enumeration
    return buf


def decompose(dims, b, e):
    """Aligned blocks of a slice range: (n_free, fixed values[k])."""
    k = len(dims)
    place = [1]
    for d in dims:
        place.append(place[-1] * d)
    out = []
    while b < e:
        j = 0
        while j < k and b % place[j + 1] == 0 and b + place[j + 1] <= e:
            j += 1
        vals, s = [], b
        for d in dims:
            vals.append(s % d)
            s //= d
        out.append((j, vals))
        b += place[j]
    return out


def run_block(desc, data, bits, fixed_vals, dtype=np.complex128, shuffle_seed=None, state=None):
    """Execute one aligned block; returns the per-bitstring partial sums.
    shuffle_seed: run the non-const ops in a RANDOM topological order of their `deps`
    edges (what a CUDA graph is allowed to do) instead of program order.
    state: a dict that receives ``locate`` (tensor description -> (buffer, base element, row stride)) so that a caller
    can read every tensor of the block afterwards (the fused-chain test of tests/test_rowprog.py)."""
    n = bits.shape[0]
    T = desc["tensors"]
    ar = desc["arena_elems"]
    arenas = {"const": np.zeros(max(ar["const"], 2), dtype=dtype),
              "block": np.zeros(max(ar["block"], 2), dtype=dtype),
              "chunk": np.zeros(max(ar["chunk_per_amp"], 2) * n, dtype=dtype)}
    leaves = {}

    def locate(t):
        off = sum(int(fixed_vals[v]) << pos for v, pos in t["fixed"])
        if t["leaf"] and not t["output_leaf"]:
            lab = t["data_label"]
            if lab not in leaves:
                leaves[lab] = pad_leaf(data[lab]).astype(dtype)
            return leaves[lab], off, 0
        sU = (1 << t["span_bits"]) if t["amp"] else 0
        if t["phase"] == "chunk":
            return arenas["chunk"], t["offset"] * n + off, sU
        return arenas[t["phase"]], t["offset"] + off, sU

    if state is not None:
        state["locate"] = locate
    for t in T:
        if t["output_leaf"]:
            buf, base, sU = locate({**t, "fixed": []})
            span = t["span_bits"]
            for u in range(n):
                val = bits[u, t["out_idx"] - 1]
                v = np.zeros(1 << span, dtype=dtype)
                if val == 0: v[0] = 1
                elif val == 1: v[1] = 1
                elif val == 2: v[:2] = 1
                else: v[0], v[1] = 1, -1
                buf[base + u * sU: base + u * sU + (1 << span)] = v

    ops = desc["ops"]
    order = [i for i, o in enumerate(ops) if o["phase"] == "const"]
    rest = [i for i, o in enumerate(ops) if o["phase"] != "const"]
    if shuffle_seed is None:
        order += [i for i in rest if ops[i]["phase"] == "block"] + [i for i in rest if ops[i]["phase"] == "chunk"]
    else:
        rng = np.random.default_rng(shuffle_seed)
        done, pending = set(order), set(rest)
        while pending:
            ready = sorted(i for i in pending if all(d in done or ops[d]["phase"] == "const" for d in ops[i]["deps"]))
            assert ready, "dependency cycle"
            pick = ready[int(rng.integers(len(ready)))]
            order.append(pick); done.add(pick); pending.discard(pick)
    for _phase in (0,):
        for oi in order:
            op = ops[oi]
            A, a0, sUA = locate(T[op["a"]])
            B, b0, sUB = locate(T[op["b"]])
            Cb, c0, sUC = locate(T[op["c"]])
            c = np.arange(1 << op["nC"], dtype=np.int64)
            fa, fb = segeval(op["segA"], c), segeval(op["segB"], c)
            ks = np.arange(1 << op["nK"], dtype=np.int64)
            ka, kb = segeval(op["segKA"], ks), segeval(op["segKB"], ks)
            U = n if T[op["c"]]["amp"] else 1
            for u in range(U):
                acc = np.zeros(1 << op["nC"], dtype=dtype)
                for k in range(len(ks)):
                    acc += A[a0 + u * sUA + fa + ka[k]] * B[b0 + u * sUB + fb + kb[k]]
                Cb[c0 + u * sUC + c] = acc
    R = T[desc["root"]]
    buf, base, sU = locate(R)
    modes = desc.get("root_modes", [])
    if any(nbits > 0 for _ext, nbits, _pos in modes):
        # tensor-valued root (reduce_root_open_kernel): output index over the true extents in Julia order, the
        # remaining address bits (batched slice variables still open in the root) are summed
        elems = int(np.prod([m[0] for m in modes], dtype=np.int64))
        o = np.arange(elems, dtype=np.int64)
        addr, rest, mode_mask = np.zeros_like(o), o.copy(), 0
        for ext, nbits, pos in modes:
            addr |= (rest % ext) << pos
            rest //= ext
            mode_mask |= ((1 << nbits) - 1) << pos
        vbits = [b for b in range(R["span_bits"]) if not (mode_mask >> b) & 1]
        out = np.zeros((n, elems), dtype=np.complex128)
        for u in range(n):
            row = buf[base + u * sU: base + u * sU + (1 << R["span_bits"])].astype(np.complex128)
            for v in range(1 << len(vbits)):
                va = sum(((v >> i) & 1) << b for i, b in enumerate(vbits))
                out[u] += row[addr | va]
        return desc["root_scale"] * out
    out = np.zeros(n, dtype=np.complex128)
    for u in range(n):
        out[u] = desc["root_scale"] * np.sum(buf[base + u * sU: base + u * sU + (1 << R["span_bits"])].astype(np.complex128))
    return out


def amplitudes(graph, data, bits, slice_begin=0, slice_end=None, dtype=np.complex128, shuffle_seed=None):
    """graph: qxb200 executor.Graph (only its host-side queries are used)."""
    dims = graph.slice_dims
    if slice_end is None:
        slice_end = graph.n_slices
    total = None
    cache = {}
    for n_free, vals in decompose(dims, slice_begin, slice_end):
        if n_free not in cache:
            cache[n_free] = graph.describe(n_free)
        part = run_block(cache[n_free], data, bits, vals, dtype, shuffle_seed)
        total = part if total is None else total + part
    if total is None:
        total = np.zeros(bits.shape[0], dtype=np.complex128)
    return total.astype(dtype)


def amplitudes_subspace(graph, data, bits, fixed_vars, fixed_vals, dtype=np.complex128, shuffle_seed=None):
    k = len(graph.slice_dims)
    mask = (1 << k) - 1
    vals = [0] * k
    for v, x in zip(fixed_vars, fixed_vals):
        mask &= ~(1 << v)
        vals[v] = x
    return run_block(graph.describe_mask(mask), data, bits, vals, dtype, shuffle_seed).astype(dtype)
