"""Host-side handle on a compiled contraction program in ``libqxb200.so``.

``Graph`` is the Python face of ``qxb_graph`` (include/qxb200.h): built either from
``.qx`` text (``execute`` path, /root/reference/bin/qxrun.jl:83-87) or from a
``ComputeGraph`` object (``single_amplitude`` path,
/root/reference/src/simulation.jl:86-91 through build_compute_graph).
All compute runs on the GPU in the library; nothing here computes amplitudes.
"""
from __future__ import annotations

import ctypes as C
import json
from typing import Dict, Iterable, List, Optional, Sequence

import numpy as np

from . import _lib
from ._lib import QXB_C32, QXB_C64, Options, Stats, check

_DT = {"c32": QXB_C32, "c64": QXB_C64, "complex64": QXB_C32, "complex128": QXB_C64,
       "ComplexF32": QXB_C32, "ComplexF64": QXB_C64}
_CHARS = {"0": 0, "1": 1, "+": 2, "-": 3}


def bits_from_strings(bitstrings: Sequence[str], n_outputs: int) -> np.ndarray:
    """char i <-> qubit i <-> ``output`` index i (docs/src/basics.md:55)."""
    out = np.zeros((len(bitstrings), max(n_outputs, 1)), dtype=np.uint8)
    for r, s in enumerate(bitstrings):
        if len(s) < n_outputs:
            raise ValueError(f"bitstring {s!r} shorter than the {n_outputs} outputs of the program")
        for c in range(n_outputs):
            try:
                out[r, c] = _CHARS[s[c]]
            except KeyError:
                raise ValueError(f"bad bitstring character {s[c]!r}") from None
    return out[:, :n_outputs] if n_outputs else out[:, :0]


class Graph:
    def __init__(self, dtype: str = "c64"):
        self._lib = _lib.load()
        self.dtype = _DT[dtype]
        self.np_dtype = np.complex64 if self.dtype == QXB_C32 else np.complex128
        h = C.c_void_p()
        check(self._lib.qxb_graph_create(C.byref(h), self.dtype))
        self._h = h
        self.compiled = False

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            self._lib.qxb_graph_destroy(h)
            self._h = None

    # -- construction ------------------------------------------------------
    @classmethod
    def from_dsl(cls, text: str, data: Dict[str, np.ndarray], dtype: str = "c64", replan: float = 0.0,
                 replan_n_amp: int = 1024) -> "Graph":
        """``replan`` > 0: re-derive the contraction order for batched execution first
        (qxb_graph_replan, exact; ``replan`` = number of candidate orders); ``replan`` < 0: decide
        from the cost model whether the search can pay for itself."""
        g = cls(dtype)
        g.replan_info = None
        g.text = text
        b = text.encode()
        check(g._lib.qxb_graph_parse_dsl(g._h, b, len(b)))
        g.set_data(data)
        if replan and replan > 0:
            g.replan(max(1, int(round(replan))), replan_n_amp)
        elif replan and replan < 0:
            # auto: re-plan only when the modelled run time (bytes / ~5 TB/s) can pay for the search
            # (~20 ms per candidate order)
            t_run = g.cost_bytes(replan_n_amp) / 5e12
            if t_run > 0.02:
                g.replan(int(min(128, max(8, t_run / 0.02))), replan_n_amp)
        return g

    @classmethod
    def from_compute_graph(cls, cg, dtype: str = "c64") -> "Graph":
        """Walk the tree post-order and issue one qxb_graph_* call per command --
        exactly what the Julia shim does with a ``ComputeGraph``."""
        from .compute_graph import (LoadCommand, OutputCommand, ViewCommand, ContractCommand, SaveCommand)
        g = cls(dtype)
        L = g._lib
        arr = lambda xs: (C.c_int64 * max(len(xs), 1))(*xs)
        for op in cg.commands():
            if isinstance(op, LoadCommand):
                check(L.qxb_graph_load(g._h, op.name.encode(), op.label.encode(), arr(op.dims), len(op.dims)))
            elif isinstance(op, OutputCommand):
                check(L.qxb_graph_output(g._h, op.name.encode(), op.idx, op.dim))
            elif isinstance(op, ViewCommand):
                check(L.qxb_graph_view(g._h, op.name.encode(), op.target.encode(), op.slice_sym.encode(),
                                       op.bond_index, op.bond_dim))
            elif isinstance(op, ContractCommand):
                check(L.qxb_graph_ncon(g._h, op.output_name.encode(), arr(op.output_idxs), len(op.output_idxs),
                                       op.left_name.encode(), arr(op.left_idxs), len(op.left_idxs),
                                       op.right_name.encode(), arr(op.right_idxs), len(op.right_idxs)))
            elif isinstance(op, SaveCommand):
                check(L.qxb_graph_save(g._h, op.label.encode(), op.name.encode()))
            else:
                raise TypeError(op)
        g.set_data(cg.tensors)
        return g

    def set_data(self, data: Dict[str, np.ndarray]) -> None:
        for label, a in data.items():
            a = np.asarray(a)
            flat = np.ascontiguousarray(a.reshape(-1, order="F").astype(np.complex128))
            dims = (C.c_int64 * max(a.ndim, 1))(*a.shape)
            check(self._lib.qxb_graph_set_data(self._h, label.encode(), flat.ctypes.data_as(C.c_void_p), dims, a.ndim))

    # -- queries (host logic only) -------------------------------------------
    @property
    def n_outputs(self) -> int:
        n = C.c_int()
        check(self._lib.qxb_graph_num_outputs(self._h, C.byref(n)))
        return n.value

    @property
    def root_dims(self) -> List[int]:
        """Shape of the saved tensor in Julia order: ``[]`` for a closed network (qxb_graph_root_dims)."""
        r = C.c_int()
        check(self._lib.qxb_graph_root_dims(self._h, C.byref(r), None))
        d = (C.c_int64 * max(r.value, 1))()
        check(self._lib.qxb_graph_root_dims(self._h, C.byref(r), d))
        return list(d[:r.value])

    def _out_buffer(self, n: int):
        """(flat buffer for the library, view to hand back): ``[n]`` amplitudes for a closed network, ``[n, *root_dims]``
        (each tensor column-major, as Julia holds it) for an open one."""
        dims = self.root_dims
        per = int(np.prod(dims, dtype=np.int64)) if dims else 1
        flat = np.zeros(n * per, dtype=self.np_dtype)
        if not dims:
            return flat, flat
        return flat, flat.reshape([n] + dims[::-1]).transpose([0] + list(range(len(dims), 0, -1)))

    @property
    def slice_dims(self) -> List[int]:
        k = C.c_int()
        check(self._lib.qxb_graph_num_slice_vars(self._h, C.byref(k), None))
        d = (C.c_int64 * max(k.value, 1))()
        check(self._lib.qxb_graph_num_slice_vars(self._h, C.byref(k), d))
        return list(d[:k.value])

    @property
    def n_slices(self) -> int:
        n = C.c_int64()
        check(self._lib.qxb_graph_num_slices(self._h, C.byref(n)))
        return n.value

    def slice_values(self, s: int) -> List[int]:
        k = len(self.slice_dims)
        v = (C.c_int64 * max(k, 1))()
        check(self._lib.qxb_slice_values(self._h, s, v))
        return list(v[:k])

    def describe(self, n_free: int = -1) -> dict:
        need = check(self._lib.qxb_graph_describe(self._h, n_free, None, 0))
        buf = C.create_string_buffer(need)
        check(self._lib.qxb_graph_describe(self._h, n_free, buf, need))
        return json.loads(buf.value.decode())

    def replan(self, candidates: int = 24, n_amp: int = 1024, n_free: int = -1, budget_bytes: int = 0,
               seed: int = 0) -> dict:
        """In-library batch-aware re-planning (qxb_graph_replan_ex); call before compile().
        ``n_free``: slice variables batched per block (-1 all, -2 chosen against ``budget_bytes``)."""
        a, b, sec, nf = C.c_double(), C.c_double(), C.c_double(), C.c_int()
        check(self._lib.qxb_graph_replan_ex(self._h, candidates, n_amp, n_free, budget_bytes, seed, C.byref(nf),
                                            C.byref(sec), C.byref(a), C.byref(b)))
        self.replan_info = {"replanned": np.isfinite(sec.value), "given_bytes": a.value, "bytes": b.value,
                            "candidates": candidates, "n_amp_model": n_amp, "n_free": nf.value,
                            "model_seconds_per_block": sec.value}
        self.text = self.program_text()
        return self.replan_info

    def program_text(self) -> str:
        need = check(self._lib.qxb_graph_program_text(self._h, None, 0))
        buf = C.create_string_buffer(need)
        check(self._lib.qxb_graph_program_text(self._h, buf, need))
        return buf.value.decode()

    def describe_mask(self, free_mask: int) -> dict:
        need = check(self._lib.qxb_graph_describe_mask(self._h, free_mask, None, 0))
        buf = C.create_string_buffer(need)
        check(self._lib.qxb_graph_describe_mask(self._h, free_mask, buf, need))
        return json.loads(buf.value.decode())

    def cost_bytes(self, n_amp: int, fixed_vars: Sequence[int] = ()) -> float:
        """Algorithmic bytes of one step (cost model; host logic only)."""
        mask = (1 << len(self.slice_dims)) - 1
        for v in fixed_vars:
            mask &= ~(1 << v)
        out = C.c_double()
        check(self._lib.qxb_graph_cost_bytes(self._h, mask, n_amp, C.byref(out)))
        return out.value

    def choose_partition(self, n_amp: int, n_parts: int, part: int):
        """How rank ``part`` of ``n_parts`` shares a step: by bitstrings (the reference's top
        level of parallelism, docs/src/users_guide.md:12-13) or by fixing slice variables
        (its lower level, :14-20) -- whichever leaves less work per rank under the cost model.
        -> ("amps", (begin, end)) | ("slices", (fixed_vars, fixed_vals)) | ("ranges", (s0, s1))"""
        from .dist import partition_range
        if n_parts <= 1:
            return "amps", (0, n_amp)
        a0, a1 = partition_range(n_amp, n_parts, part)
        per_rank_amps = self.cost_bytes(max(1, (n_amp + n_parts - 1) // n_parts))
        assign = self.partition_assignment(n_parts, part)
        if assign is None:
            if n_amp >= n_parts:
                return "amps", (a0, a1)
            return "ranges", partition_range(self.n_slices, n_parts, part)
        per_rank_slices = self.cost_bytes(n_amp, assign[0])
        if per_rank_amps <= per_rank_slices and n_amp >= n_parts:
            return "amps", (a0, a1)
        return "slices", assign

    def partition_vars(self, n_parts: int) -> List[int]:
        """0-based slice variables to fix when sharding over ``n_parts`` ranks ([] = use ranges)."""
        n = C.c_int()
        out = (C.c_int32 * 64)()
        check(self._lib.qxb_partition_vars(self._h, n_parts, out, C.byref(n)))
        return list(out[:n.value])

    def partition_assignment(self, n_parts: int, part: int):
        """(fixed_vars, fixed_vals) of rank ``part``, or None when ranges must be used."""
        vs = self.partition_vars(n_parts)
        if n_parts > 1 and not vs:
            return None
        dims = self.slice_dims
        vals, r = [], part
        for v in vs:
            vals.append(r % dims[v])
            r //= dims[v]
        return vs, vals

    # -- compute -------------------------------------------------------------
    @staticmethod
    def _options(hbm_budget_bytes: int = 0, amp_batch: int = 0, profile: bool = False,
                 cuda_graph: bool = True, sum_at_root: bool = False, smem_stage: bool = True,
                 gemm: bool = True, gemm_mode: int = 0, row_programs=True, min_lob: int = 0,
                 kc_regs_multi: int = 0, kc_regs_one: int = 0, smem_tma: bool = True, row_min_tt_bits: int = 0,
                 row_tile_regs: int = 0, row_ctas_per_sm: int = 0, row_chunk_max_amps: int = 0, ring: bool = True, chain: bool = True, row_dmma: bool = False,
                 streaming=True, row_bank_opt: bool = True, chain_side: int = 0) -> Options:
        """``qxb_options`` (include/qxb200.h); 0 / default = the library's own choice.
        ``row_programs``: True = auto (block phase always, chunk phase for small calls), False = never,
        "all" = both phases whenever the row fits shared memory, "block" = block phase only.
        ``streaming``: True = auto (huge x tiny nodes on ``bigsmall_kernel``), False = never, "tma" / "ffma2" = its
        opt-in variants.  ``row_bank_opt``: bank-aware lane bits / arena layouts in row programs.  ``chain_side``: depth of
        the side branches a fused chain takes in (0 = none)."""
        rp = {True: 0, False: 1, "auto": 0, "never": 1, "all": 2, "block": 3}[row_programs]
        return Options(hbm_budget_bytes, amp_batch, 1 if profile else 0, 0 if cuda_graph else 1,
                       1 if sum_at_root else 0, 0 if smem_stage else 1, 0 if gemm else 1, gemm_mode,
                       rp, min_lob, kc_regs_multi, kc_regs_one, 0 if smem_tma else 2,
                       row_min_tt_bits, row_tile_regs, row_ctas_per_sm, 0 if ring else 1, 0 if chain else 1, 2 if row_dmma else 0, row_chunk_max_amps,
                       {True: 0, False: 1, "tma": 2, "ffma2": 3}[streaming], 0 if row_bank_opt else 1, chain_side)

    def configure(self, **kw) -> "Graph":
        o = self._options(**kw)
        check(self._lib.qxb_graph_configure(self._h, C.byref(o)))
        return self

    def compile(self, **kw) -> "Graph":
        """Lower, fold constants, build the launch plans.  Keyword arguments = the fields of ``qxb_options``
        (see ``_options``): e.g. ``profile=True``, ``cuda_graph=False``, ``row_programs=False``, ``min_lob=8``."""
        o = self._options(**kw)
        check(self._lib.qxb_graph_compile(self._h, C.byref(o)))
        self.compiled = True
        return self

    def amplitudes(self, bitstrings, slice_begin: int = 0, slice_end: Optional[int] = None) -> np.ndarray:
        """Host-buffer entry: H2D of the bitstrings, compute, D2H of the amplitudes."""
        if isinstance(bitstrings, np.ndarray) and bitstrings.dtype == np.uint8:
            bits = np.ascontiguousarray(bitstrings)
        else:
            bits = np.ascontiguousarray(bits_from_strings(list(bitstrings), self.n_outputs))
        n = bits.shape[0]
        if slice_end is None:
            slice_end = self.n_slices
        out, view = self._out_buffer(n)
        check(self._lib.qxb_amplitudes(self._h, bits.ctypes.data_as(C.c_void_p), n, slice_begin, slice_end,
                                       out.ctypes.data_as(C.c_void_p)))
        return view

    def amplitudes_device(self, d_bits_ptr: int, n_amp: int, d_out_ptr: int, slice_begin: int = 0,
                          slice_end: Optional[int] = None) -> None:
        """Device-pointer entry (asynchronous on the library's stream)."""
        if slice_end is None:
            slice_end = self.n_slices
        check(self._lib.qxb_amplitudes_device(self._h, C.c_void_p(d_bits_ptr), n_amp, slice_begin, slice_end,
                                              C.c_void_p(d_out_ptr)))

    def amplitudes_subspace(self, bitstrings, fixed_vars: Sequence[int], fixed_vals: Sequence[int]) -> np.ndarray:
        if isinstance(bitstrings, np.ndarray) and bitstrings.dtype == np.uint8:
            bits = np.ascontiguousarray(bitstrings)
        else:
            bits = np.ascontiguousarray(bits_from_strings(list(bitstrings), self.n_outputs))
        n = bits.shape[0]
        out, view = self._out_buffer(n)
        fv = (C.c_int32 * max(len(fixed_vars), 1))(*fixed_vars)
        fx = (C.c_int64 * max(len(fixed_vals), 1))(*fixed_vals)
        check(self._lib.qxb_amplitudes_subspace(self._h, bits.ctypes.data_as(C.c_void_p), n, fv, fx, len(fixed_vars),
                                                out.ctypes.data_as(C.c_void_p), 0))
        return view

    def amplitudes_subspace_device(self, d_bits_ptr: int, n_amp: int, d_out_ptr: int, fixed_vars, fixed_vals) -> None:
        fv = (C.c_int32 * max(len(fixed_vars), 1))(*fixed_vars)
        fx = (C.c_int64 * max(len(fixed_vals), 1))(*fixed_vals)
        check(self._lib.qxb_amplitudes_subspace(self._h, C.c_void_p(d_bits_ptr), n_amp, fv, fx, len(fixed_vars),
                                                C.c_void_p(d_out_ptr), 1))

    def stats(self) -> dict:
        s = Stats()
        check(self._lib.qxb_last_stats(self._h, C.byref(s)))
        return {k: getattr(s, k) for k, _ in Stats._fields_}

    def profile_dump(self, path: str) -> dict:
        check(self._lib.qxb_profile_dump(self._h, path.encode()))
        with open(path) as f:
            return json.load(f)


class MultiGraph:
    """Several GPUs of one node behind the C ABI (``qxb_multi_*``, include/qxb200.h): one compiled replica of an
    UNCOMPILED ``Graph`` per device, driven by this one process.  ``amplitudes`` splits the bitstrings across groups of
    ``sub_comm_size`` devices and the slice range inside a group (0 = auto), like ``-m`` / ``-s`` of bin/qxrun.jl."""

    def __init__(self, graph: "Graph", n_devices: int = 0, device_ids: Optional[Sequence[int]] = None, **compile_kw):
        self._lib = graph._lib
        self._h = C.c_void_p()
        self.n_outputs, self.np_dtype, self.n_slices = graph.n_outputs, graph.np_dtype, graph.n_slices
        o = Graph._options(**compile_kw)
        ids = (C.c_int * len(device_ids))(*device_ids) if device_ids else None
        check(self._lib.qxb_multi_create(C.byref(self._h), graph._h, len(device_ids) if device_ids else n_devices, ids, C.byref(o)))
        n = C.c_int()
        check(self._lib.qxb_multi_num_devices(self._h, C.byref(n)))
        self.n_devices = n.value

    def amplitudes(self, bitstrings, slice_begin: int = 0, slice_end: Optional[int] = None, sub_comm_size: int = 0) -> np.ndarray:
        if isinstance(bitstrings, np.ndarray) and bitstrings.dtype == np.uint8:
            bits = np.ascontiguousarray(bitstrings)
        else:
            bits = np.ascontiguousarray(bits_from_strings(list(bitstrings), self.n_outputs))
        n = bits.shape[0]
        out = np.zeros(n, dtype=self.np_dtype)
        check(self._lib.qxb_multi_amplitudes(self._h, bits.ctypes.data_as(C.c_void_p), n, slice_begin,
                                             self.n_slices if slice_end is None else slice_end, sub_comm_size,
                                             out.ctypes.data_as(C.c_void_p)))
        return out

    def __del__(self):
        if getattr(self, "_h", None) and self._h.value:
            self._lib.qxb_multi_destroy(self._h)
            self._h = C.c_void_p()


def autotune(candidates, build, probe, reduce_times=None, rel_tol: float = 1e-9):
    """Measured choice among EXACT alternatives of the same program (different contraction trees from the re-planner,
    different kernel-selection knobs): every candidate is built, run on the same probe input and timed; a candidate
    whose result differs from the first one's by more than ``rel_tol`` (relative to the largest amplitude), or that
    raises, is discarded.  The model picks the candidates, the GPU picks the winner.

    ``candidates``: ``[(tag, plan_text, options)]`` -- ``options`` = keyword arguments of ``Graph.compile`` (the fields
    of ``qxb_options``: nothing is left behind in ``os.environ``); the first candidate is the baseline.
    ``build(plan_text, **options) -> graph``; ``probe(graph) -> (milliseconds, result ndarray)``;
    ``reduce_times(list) -> list`` combines the times over the ranks of a multi-GPU job (max), so that every rank takes
    the same decision.  Returns ``(index, report)``; index 0 when nothing could be measured."""
    times, report, ref = [], [], None
    for tag, text, kw in candidates:
        ms, note = float("inf"), "ok"
        try:
            g = build(text, **kw)
            t, res = probe(g)
            res = np.asarray(res)
            del g
            if ref is None:
                if len(times) == 0:
                    ref, ms = res, float(t)
                else:
                    note = "no baseline result to compare with"
            else:
                scale = max(float(np.max(np.abs(ref))), 1e-300)
                diff = float(np.max(np.abs(res - ref))) / scale
                if diff <= rel_tol:
                    ms = float(t)
                else:
                    note = f"discarded: differs from the baseline by {diff:.2e}"
        except Exception as e:                       # a candidate must never take the run down
            note = f"failed: {e!r}"[:200]
        times.append(ms)
        report.append({"tag": tag, "ms": None if ms == float("inf") else ms, "note": note})
    if reduce_times is not None:
        times = [float(t) for t in reduce_times(list(times))]
        for r, t in zip(report, times):
            r["ms_max_over_ranks"] = None if t == float("inf") else t
    best = min(range(len(times)), key=lambda i: (times[i], i)) if times and min(times) < float("inf") else 0
    return best, report


def init(device: int = 0) -> None:
    check(_lib.load().qxb_init(device))


def set_stream(stream_ptr: int) -> None:
    check(_lib.load().qxb_set_stream(C.c_void_p(stream_ptr)))


def fma_peak(dtype: str = "c64") -> float:
    """Measured FMA-pipe peak of the current device in TFLOP/s (``qxb_fma_peak``): FFMA for c32, DFMA for c64,
    packed FFMA2 (fma.rn.f32x2) for ``"c32x2"``."""
    v = C.c_double()
    check(_lib.load().qxb_fma_peak({"c32": QXB_C32, "c64": QXB_C64, "c32x2": 2}[dtype], C.byref(v)))
    return v.value


def synchronize() -> None:
    check(_lib.load().qxb_device_synchronize())


def expand_open_indices(reduced: np.ndarray, root_indices: Sequence[int], wires: Sequence[int]) -> np.ndarray:
    """Saved tensor (one mode per distinct index id, hyper-edges merged) -> the full tensor over the open ``wires``:
    wires that share an id are forced equal (the diagonal structure the reference keeps implicit in its hyper-index
    groups and QXTns materialises in the result of ``contract_tn!``).  Pure index bookkeeping on the host."""
    reduced = np.asarray(reduced)
    if not wires:
        return reduced
    mode = {ix: m for m, ix in enumerate(root_indices)}
    missing = [w for w in wires if w not in mode]
    if missing:
        raise ValueError(f"open wires {missing} are not modes of the saved tensor")
    full = np.zeros([reduced.shape[mode[w]] for w in wires], dtype=reduced.dtype)
    idx = np.indices(reduced.shape)
    full[tuple(idx[mode[w]] for w in wires)] = reduced
    return full


def contract_tn(tnc, plan, dtype: str = "c64") -> np.ndarray:
    """``QXTns.contract_tn!(tnc, plan)`` (call sites /root/reference/src/simulation.jl:89 and
    /root/reference/test/test_contraction_planning.jl:58,111,126,142,158): contract the whole network along ``plan`` and
    return the data of the remaining tensor -- a 1-element array for a closed network; for a network built with
    ``no_output=True`` the tensor over the open output wires, qubit k on axis k (``out.reshape(-1, order="F")`` is the
    reference's ``reshape(output, prod(size(output)))``), followed by the open input wires when ``no_input=True``.
    The contraction itself runs on the GPU (tensor-valued ``save``, ``qxb_graph_root_dims``); the caller's network is
    not modified."""
    from .compute_graph import build_compute_graph
    cg = build_compute_graph(tnc, plan)
    g = Graph.from_compute_graph(cg, dtype).compile()
    out = g.amplitudes(["0" * g.n_outputs])          # output tensors, if the network has any, are the |0> projectors
    if not g.root_dims:
        return np.asarray(out[:1])
    open_ids = set(cg.root_indices)
    wires = [w for w in tnc.wire if w in open_ids] + [w for w in tnc.wire_first if w in open_ids and w not in tnc.wire]
    return expand_open_indices(out[0], cg.root_indices, wires)


def amplitudes_for_network(tnc, plan, bitstrings: Sequence[str], dtype: str = "c64") -> np.ndarray:
    """``contract_tn!`` for a batch of bitstrings: build the (unsliced) compute graph
    once, then one library call (simulation.jl:86-91, compute_graph.jl:15-98)."""
    from .compute_graph import build_compute_graph
    cg = build_compute_graph(tnc, plan)
    g = Graph.from_compute_graph(cg, dtype).compile()
    return g.amplitudes(list(bitstrings))
