"""ctypes binding of ``libqxb200.so`` (C ABI: ``include/qxb200.h``).

The library is built in-tree by ``csrc/Makefile`` (``__graft_entry__.build()``).
There is no fallback: if the shared object is missing this module raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libqxb200.so")

QXB_C32, QXB_C64 = 0, 1
ERR_NAMES = {0: "OK", -1: "ERR_ARG", -2: "ERR_STATE", -3: "ERR_CUDA", -4: "ERR_UNSUPP", -5: "ERR_MEM"}

# every symbol include/qxb200.h declares (tests check that all of them are exported)
SYMBOLS = [
    "qxb_version", "qxb_last_error", "qxb_init", "qxb_shutdown", "qxb_set_stream", "qxb_device_synchronize", "qxb_fma_peak",
    "qxb_graph_create", "qxb_graph_destroy", "qxb_graph_load", "qxb_graph_output", "qxb_graph_view",
    "qxb_graph_ncon", "qxb_graph_save", "qxb_graph_parse_dsl", "qxb_graph_set_data",
    "qxb_graph_num_outputs", "qxb_graph_root_dims", "qxb_graph_num_slice_vars", "qxb_graph_num_slices", "qxb_slice_values",
    "qxb_graph_describe", "qxb_graph_replan", "qxb_graph_replan_ex", "qxb_graph_program_text", "qxb_graph_configure", "qxb_graph_compile", "qxb_amplitudes", "qxb_amplitudes_device",
    "qxb_amplitudes_subspace", "qxb_partition_vars", "qxb_graph_describe_mask", "qxb_graph_cost_bytes",
    "qxb_last_stats", "qxb_profile_dump", "qxb_debug_mma_smem_bit",
    "qxb_jld2_open", "qxb_jld2_close", "qxb_jld2_count", "qxb_jld2_info", "qxb_jld2_read", "qxb_jld2_write",
    "qxb_graph_load_jld2", "qxb_params_read", "qxb_execute_files", "qxb_debug_lookup3", "qxb_debug_templates",
    "qxb_debug_rowprog", "qxb_debug_fused_plan", "qxb_debug_tc5_smem_bit",
    "qxb_multi_create", "qxb_multi_destroy", "qxb_multi_num_devices", "qxb_multi_amplitudes", "qxb_execute_files_multi",
]


class QxbError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libqxb200 {ERR_NAMES.get(code, code)}: {msg}")
        self.code = code


class Options(C.Structure):
    _fields_ = [("hbm_budget_bytes", C.c_int64), ("amp_batch", C.c_int64),
                ("profile", C.c_int32), ("no_cuda_graph", C.c_int32),
                ("sum_at_root", C.c_int32), ("no_smem_stage", C.c_int32),
                ("no_gemm", C.c_int32), ("gemm_mode", C.c_int32),
                ("row_programs", C.c_int32), ("min_lob", C.c_int32), ("kc_regs_multi", C.c_int32),
                ("kc_regs_one", C.c_int32), ("smem_tma", C.c_int32), ("row_min_tt_bits", C.c_int32),
                ("row_tile_regs", C.c_int32), ("row_ctas_per_sm", C.c_int32), ("ring", C.c_int32),
                ("chain", C.c_int32), ("row_dmma", C.c_int32), ("row_chunk_max_amps", C.c_int32),
                ("streaming", C.c_int32), ("row_bank_opt", C.c_int32), ("chain_side", C.c_int32)]


class Params(C.Structure):
    """qxb_params (include/qxb200.h): the parameter file of a triple, outputs.jl:47-78."""
    _fields_ = [("method", C.c_int32), ("has_seed", C.c_int32), ("seed", C.c_int64), ("num_qubits", C.c_int64),
                ("num_samples", C.c_int64), ("M", C.c_double), ("fix_M", C.c_int32), ("reserved", C.c_int32),
                ("n_bitstrings", C.c_int64)]


class Stats(C.Structure):
    _fields_ = [("kernel_launches", C.c_int64), ("contract_launches", C.c_int64),
                ("flops", C.c_double), ("bytes", C.c_double), ("workspace_bytes", C.c_int64),
                ("amp_batch", C.c_int64), ("n_blocks", C.c_int64)]


_lib = None


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    p, i32, i64, cp = C.c_void_p, C.c_int, C.c_int64, C.c_char_p
    pi64 = C.POINTER(C.c_int64)
    sig = {
        "qxb_version": (i32, []),
        "qxb_last_error": (cp, []),
        "qxb_init": (i32, [i32]),
        "qxb_shutdown": (i32, []),
        "qxb_set_stream": (i32, [p]),
        "qxb_device_synchronize": (i32, []),
        "qxb_fma_peak": (i32, [i32, C.POINTER(C.c_double)]),
        "qxb_graph_create": (i32, [C.POINTER(p), i32]),
        "qxb_graph_destroy": (None, [p]),
        "qxb_graph_load": (i32, [p, cp, cp, pi64, i32]),
        "qxb_graph_output": (i32, [p, cp, i64, i64]),
        "qxb_graph_view": (i32, [p, cp, cp, cp, i64, i64]),
        "qxb_graph_ncon": (i32, [p, cp, pi64, i32, cp, pi64, i32, cp, pi64, i32]),
        "qxb_graph_save": (i32, [p, cp, cp]),
        "qxb_graph_parse_dsl": (i32, [p, cp, C.c_size_t]),
        "qxb_graph_set_data": (i32, [p, cp, p, pi64, i32]),
        "qxb_graph_num_outputs": (i32, [p, C.POINTER(i32)]),
        "qxb_graph_root_dims": (i32, [p, C.POINTER(i32), pi64]),
        "qxb_graph_num_slice_vars": (i32, [p, C.POINTER(i32), pi64]),
        "qxb_graph_num_slices": (i32, [p, pi64]),
        "qxb_slice_values": (i32, [p, i64, pi64]),
        "qxb_graph_describe": (i64, [p, i32, cp, i64]),
        "qxb_graph_replan": (i32, [p, i32, i64, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
        "qxb_graph_replan_ex": (i32, [p, i32, i64, i32, i64, C.c_uint64, C.POINTER(C.c_int), C.POINTER(C.c_double),
                                      C.POINTER(C.c_double), C.POINTER(C.c_double)]),
        "qxb_graph_program_text": (i64, [p, cp, i64]),
        "qxb_graph_configure": (i32, [p, C.POINTER(Options)]),
        "qxb_graph_compile": (i32, [p, C.POINTER(Options)]),
        "qxb_amplitudes": (i32, [p, p, i64, i64, i64, p]),
        "qxb_amplitudes_device": (i32, [p, p, i64, i64, i64, p]),
        "qxb_amplitudes_subspace": (i32, [p, p, i64, C.POINTER(C.c_int32), pi64, i32, p, i32]),
        "qxb_partition_vars": (i32, [p, i32, C.POINTER(C.c_int32), C.POINTER(i32)]),
        "qxb_graph_describe_mask": (i64, [p, C.c_uint64, cp, i64]),
        "qxb_graph_cost_bytes": (i32, [p, C.c_uint64, i64, C.POINTER(C.c_double)]),
        "qxb_last_stats": (i32, [p, C.POINTER(Stats)]),
        "qxb_profile_dump": (i32, [p, cp]),
        "qxb_debug_mma_smem_bit": (i32, [i32, i32, i32, i32]),
        "qxb_debug_tc5_smem_bit": (i32, [i32, i32]),
        "qxb_jld2_open": (i32, [cp, C.POINTER(p)]),
        "qxb_jld2_close": (None, [p]),
        "qxb_jld2_count": (i32, [p, C.POINTER(i32), C.POINTER(i32)]),
        "qxb_jld2_info": (i32, [p, i32, C.POINTER(cp), C.POINTER(i32), C.POINTER(i32), C.POINTER(i32), pi64]),
        "qxb_jld2_read": (i32, [p, i32, p, i32]),
        "qxb_jld2_write": (i32, [cp, i32, C.POINTER(cp), C.POINTER(i32), C.POINTER(i32), C.POINTER(i32),
                                 C.POINTER(pi64), C.POINTER(p), i32]),
        "qxb_graph_load_jld2": (i32, [p, cp, C.POINTER(i32)]),
        "qxb_params_read": (i32, [cp, C.POINTER(Params), p, i64]),
        "qxb_debug_templates": (i64, [p, i32, p, i64]),
        "qxb_debug_rowprog": (i64, [p, C.c_uint64, i32, p, i64]),
        "qxb_debug_fused_plan": (i64, [p, C.c_uint64, p, i64]),
        "qxb_debug_lookup3": (C.c_uint32, [p, C.c_size_t, C.c_uint32]),
        "qxb_execute_files": (i32, [cp, cp, cp, cp, i32, i64, i64, i32, pi64, C.POINTER(C.c_double)]),
        "qxb_execute_files_multi": (i32, [cp, cp, cp, cp, i32, i64, i64, i32, i32, i32, pi64, C.POINTER(C.c_double)]),
        "qxb_multi_create": (i32, [C.POINTER(p), p, i32, C.POINTER(C.c_int), C.POINTER(Options)]),
        "qxb_multi_destroy": (None, [p]),
        "qxb_multi_num_devices": (i32, [p, C.POINTER(i32)]),
        "qxb_multi_amplitudes": (i32, [p, p, i64, i64, i64, i32, p]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc):
    if rc < 0:
        raise QxbError(rc, load().qxb_last_error().decode("utf-8", "replace"))
    return rc
