"""``.jld2`` data files through libqxb200's native reader / writer (``csrc/qxb_jld2.cpp``).

The reference keeps leaf tensors in a JLD2 file, one dataset per data label holding the N-d
ComplexF64 array (/root/reference/src/compute_graph/tensor_cache.jl:90-106; read back with
``load`` in /root/reference/test/test_compute_graph.jl:19-22), and ``bin/qxrun.jl -o`` names a
``.jld2`` results file (/root/reference/docs/src/distributed.md:30-33).  This module is the thin
``ctypes`` view of the ``qxb_jld2_*`` entry points; no Python HDF5 package is involved.
"""
from __future__ import annotations

import ctypes as C
from collections import OrderedDict
from typing import Dict, Mapping

import numpy as np

from ._lib import Params, check, load

# element kinds of include/qxb200.h
C64, C32, F64, F32, INT, STRING, OTHER = range(7)
_NP = {C64: np.complex128, C32: np.complex64, F64: np.float64, F32: np.float32}


def load_jld2(path: str, as_c64: bool = False, info: dict | None = None) -> "OrderedDict[str, np.ndarray]":
    """``label -> array`` for every decodable dataset of the file (Julia's ``load(path)``).

    Arrays come back in Julia's column-major order (Fortran-contiguous, same shape as in Julia).
    ``as_c64=True`` converts every numeric dataset to ComplexF64 -- what ``qxb_graph_set_data``
    takes.  ``info`` (optional dict) receives ``checksum_failures`` and the names of skipped datasets."""
    lib = load()
    h = C.c_void_p()
    check(lib.qxb_jld2_open(path.encode(), C.byref(h)))
    try:
        n, bad = C.c_int(), C.c_int()
        check(lib.qxb_jld2_count(h, C.byref(n), C.byref(bad)))
        out: "OrderedDict[str, np.ndarray]" = OrderedDict()
        skipped = []
        for i in range(n.value):
            name, kind, esize, rank = C.c_char_p(), C.c_int(), C.c_int(), C.c_int()
            dims = (C.c_int64 * 32)()
            check(lib.qxb_jld2_info(h, i, C.byref(name), C.byref(kind), C.byref(esize), C.byref(rank), dims))
            shape = tuple(dims[k] for k in range(rank.value))
            key = name.value.decode("utf-8", "replace")
            if kind.value == OTHER:
                skipped.append(key)
                continue
            if as_c64 and kind.value != STRING:
                dt = np.dtype(np.complex128)
            elif kind.value in _NP:
                dt = np.dtype(_NP[kind.value])
            elif kind.value == INT:
                dt = np.dtype(f"<i{esize.value}")
            else:
                dt = np.dtype(f"S{esize.value}")
            count = int(np.prod(shape, dtype=np.int64)) if shape else 1
            flat = np.empty(count, dtype=dt)
            if count:
                check(lib.qxb_jld2_read(h, i, flat.ctypes.data_as(C.c_void_p), 1 if (as_c64 and kind.value != STRING) else 0))
            out[key] = flat.reshape(shape, order="F") if shape else flat.reshape(())
        if info is not None:
            info["checksum_failures"] = bad.value
            info["skipped"] = skipped
        return out
    finally:
        lib.qxb_jld2_close(h)


def save_jld2(path: str, arrays: Mapping[str, np.ndarray], commit_types: bool = True) -> None:
    """Write ``label -> array`` as a JLD2-layout file (``save_cache``, tensor_cache.jl:100-106).

    complex128 / complex64 / float64 / float32 / int64 / fixed-length bytes (``S<n>``) arrays.
    ``commit_types=True`` mirrors JLD2.jl (complex datatype committed under ``_types/``)."""
    lib = load()
    names, kinds, sizes, ranks, dimarrs, bufs = [], [], [], [], [], []
    for k, v in arrays.items():
        a = np.asarray(v)
        if a.dtype == np.complex128:
            kind = C64
        elif a.dtype == np.complex64:
            kind = C32
        elif a.dtype == np.float64:
            kind = F64
        elif a.dtype == np.float32:
            kind = F32
        elif a.dtype.kind == "i":
            a, kind = a.astype("<i8"), INT
        elif a.dtype.kind == "S":
            kind = STRING
        elif a.dtype.kind == "U":
            a, kind = np.char.encode(a, "ascii"), STRING
        else:
            raise TypeError(f"{k}: dtype {a.dtype} cannot be written")
        shape = a.shape                      # asfortranarray promotes 0-d to 1-d: keep the Julia scalar a scalar
        a = np.asfortranarray(a)
        names.append(k.encode())
        kinds.append(kind)
        sizes.append(a.dtype.itemsize)
        ranks.append(len(shape))
        dimarrs.append((C.c_int64 * max(1, len(shape)))(*shape))
        bufs.append(a)
    n = len(names)
    c_names = (C.c_char_p * n)(*names)
    c_kinds = (C.c_int * n)(*kinds)
    c_sizes = (C.c_int * n)(*sizes)
    c_ranks = (C.c_int * n)(*ranks)
    c_dims = (C.POINTER(C.c_int64) * n)(*[C.cast(d, C.POINTER(C.c_int64)) for d in dimarrs])
    c_data = (C.c_void_p * n)(*[b.ctypes.data for b in bufs])
    check(lib.qxb_jld2_write(path.encode(), n, c_names, c_kinds, c_sizes, c_ranks, c_dims, c_data, int(commit_types)))


def read_params(path: str) -> dict:
    """The parameter file through the library's YAML-subset reader (``qxb_params_read``):
    ``{"method", "num_qubits", "num_samples", "seed", "M", "fix_M", "bitstrings"}`` (outputs.jl:47-78)."""
    lib = load()
    p = Params()
    check(lib.qxb_params_read(path.encode(), C.byref(p), None, 0))
    need = p.n_bitstrings * (p.num_qubits + 1)
    buf = C.create_string_buffer(max(1, need))
    check(lib.qxb_params_read(path.encode(), C.byref(p), buf, need))
    w = p.num_qubits + 1
    bs = [buf.raw[i * w:(i + 1) * w - 1].decode() for i in range(p.n_bitstrings)]
    return {"method": ("List", "Uniform", "Rejection")[p.method], "num_qubits": p.num_qubits,
            "num_samples": p.num_samples, "seed": p.seed if p.has_seed else None, "M": p.M,
            "fix_M": bool(p.fix_M), "bitstrings": bs}


def load_data_file(path: str) -> Dict[str, np.ndarray]:
    """Leaf tensors of a triple: ``.jld2`` (native reader) or the harness' ``.npz``."""
    if path.endswith(".npz"):
        return dict(np.load(path))
    return dict(load_jld2(path, as_c64=True))
