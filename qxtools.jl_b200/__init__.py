"""qxb200 -- B200-native executor for the QXTools contraction hot path.

The directory is ``qxtools.jl_b200/`` (not an importable name); import it as
``qxb200`` through the loader module ``qxb200.py`` at the repo root.

Host-side mirror of the reference's Julia API for the path (same names and
argument meaning as ``/root/reference/src``); the arithmetic is in
``csrc/`` -> ``libqxb200.so`` (C ABI declared in ``include/qxb200.h``).
"""
from .circuits import (Circuit, Gate, create_test_circuit, create_ghz_circuit, create_qft_circuit,
                       create_rqc_circuit, create_sycamore_like_circuit, gate_matrix, gate_qubits)
from .tn import TensorNetworkCircuit, convert_to_tnc
from .planning import (convert_to_graph, convert_to_line_graph, contraction_scheme,
                       flow_cutter_contraction_plan, min_fill_contraction_plan, min_fill,
                       order_to_contraction_plan)
from .compute_graph import (build_compute_graph, TensorCache, save_cache, ComputeGraph, write_dsl,
                            LoadCommand, OutputCommand, ViewCommand, ContractCommand, SaveCommand)
from .simulation import (single_amplitude, run_simulation, generate_simulation_files,
                         generate_parameter_file, generate_dsl_files, amplitudes_all,
                         amplitudes_uniform, output_params_dict)



def contract_tn(tnc, plan, dtype: str = "c64"):
    """``QXTns.contract_tn!`` on the GPU, open networks included (see ``executor.contract_tn``)."""
    from .executor import contract_tn as _impl
    return _impl(tnc, plan, dtype)


__version__ = "0.1.0"
