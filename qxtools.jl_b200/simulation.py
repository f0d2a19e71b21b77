"""Simulation API of the hot path -- same names and argument meaning as
``/root/reference/src/simulation.jl`` and ``src/outputs.jl``.

* ``single_amplitude(tnc, plan, amplitude)``       simulation.jl:86-91
* ``run_simulation(circ; num_amplitudes, seed)``    simulation.jl:102-133
* ``generate_simulation_files(circ, prefix, nslice; ...)``  simulation.jl:52-77
* ``amplitudes_all`` / ``amplitudes_uniform``          simulation.jl:14-28
* ``output_params_dict`` / ``generate_parameter_file`` outputs.jl:47-78, simulation.jl:149-155

The arithmetic (``contract_tn!`` in the reference) is done by ``libqxb200.so`` on the
GPU through :mod:`executor`; there is NO CPU fallback -- a missing library or GPU
raises.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Dict, Iterable, List, Optional, Sequence

import numpy as np
import yaml

from .circuits import Circuit
from .compute_graph import build_compute_graph, save_cache, TensorCache, write_dsl
from .planning import contraction_scheme, flow_cutter_contraction_plan
from .tn import TensorNetworkCircuit, convert_to_tnc


def amplitudes_all(qubits: int):
    """All 2^n bitstrings in lexicographic (big-endian text) order (simulation.jl:14-16)."""
    return (format(x, f"0{qubits}b") for x in range(2 ** qubits))


def amplitudes_uniform(qubits: int, seed: Optional[int], number_amplitudes: int):
    """Uniform random bitstrings WITH replacement (simulation.jl:24-28).  The
    reference draws from Julia's MersenneTwister; here numpy's PCG64 -- List
    parameter files carry explicit strings, so the executor never depends on it."""
    rng = np.random.default_rng(seed)
    return ("".join("01"[b] for b in rng.integers(0, 2, qubits)) for _ in range(number_amplitudes))


def output_params_dict(num_qubits: int, num_outputs: int = 10, output_method: str = "List",
                       seed: Optional[int] = None, M: float = 0.0001, fix_M: bool = False,
                       bitstrings: Optional[Sequence[str]] = None) -> "OrderedDict":
    """outputs.jl:47-78."""
    out = OrderedDict()
    out["method"] = output_method
    p = OrderedDict()
    if output_method == "Rejection":
        p["num_qubits"] = num_qubits
        p["M"] = M
        p["fix_M"] = fix_M
        p["seed"] = seed
        p["num_samples"] = num_outputs
    elif output_method == "List":
        if bitstrings is None:
            bitstrings = list(amplitudes_uniform(num_qubits, seed, num_outputs))
        else:
            num_outputs = len(bitstrings)
        p["num_samples"] = num_outputs
        p["bitstrings"] = list(bitstrings)
    elif output_method == "Uniform":
        p["num_qubits"] = num_qubits
        p["num_samples"] = num_outputs
        p["seed"] = seed
    else:
        raise ValueError(f'Output method "{output_method}" not supported')
    out["params"] = p
    return out


def _plain(o):
    if isinstance(o, dict):
        return {k: _plain(v) for k, v in o.items()}
    if isinstance(o, (list, tuple)):
        return [_plain(v) for v in o]
    return o


def generate_parameter_file(filename_prefix: str, output_parameters) -> None:
    """simulation.jl:149-155."""
    with open(f"{filename_prefix}.yml", "w") as f:
        yaml.safe_dump({"output": _plain(output_parameters)}, f, sort_keys=False)


def generate_dsl_files(compute_graph, prefix: str, force: bool = True, metadata=None, npz: bool = False) -> None:
    """QXContexts.generate_dsl_files (call site simulation.jl:73): ``<prefix>.qx`` and
    the tensor data file ``<prefix>.jld2`` (native writer, ``csrc/qxb_jld2.cpp``); ``npz=True`` also
    keeps the ``.npz`` copy the benchmark caches use."""
    import os
    from .jld2 import save_jld2
    if not force and (os.path.exists(prefix + ".qx") or os.path.exists(prefix + ".jld2")):
        raise FileExistsError(prefix)
    with open(prefix + ".qx", "w") as f:
        f.write(write_dsl(compute_graph, metadata))
    save_jld2(prefix + ".jld2", {k: np.asarray(v, dtype=np.complex128) for k, v in compute_graph.tensors.items()})
    if npz:
        np.savez(prefix + ".npz", **{k: np.asfortranarray(v) for k, v in compute_graph.tensors.items()})


def generate_simulation_files(circ: Circuit, output_prefix: str = "simulation_input",
                              number_bonds_to_slice: int = 2, decompose: bool = True,
                              seed: Optional[int] = None, output_args=None, npz: bool = False, **kwargs) -> None:
    """simulation.jl:52-77: writes ``<prefix>.qx``, ``<prefix>.jld2`` and ``<prefix>.yml``."""
    tnc = convert_to_tnc(circ, decompose=decompose)
    fc_seed = -1 if seed is None else seed
    bond_groups, plan, metadata = contraction_scheme(tnc, number_bonds_to_slice, seed=fc_seed, **kwargs)
    cg = build_compute_graph(tnc, plan, bond_groups)
    generate_dsl_files(cg, output_prefix, force=True, metadata=metadata, npz=npz)
    if output_args is None:
        output_args = output_params_dict(tnc.qubits)
    generate_parameter_file(output_prefix, output_args)


def single_amplitude(tnc: TensorNetworkCircuit, plan, amplitude: Optional[str] = None,
                     dtype: str = "c64"):
    """simulation.jl:86-91: the caller's ``tnc`` is not mutated; returns a scalar."""
    from .executor import amplitudes_for_network
    if amplitude is None:
        amplitude = "0" * tnc.qubits
    return complex(amplitudes_for_network(tnc, plan, [amplitude], dtype=dtype)[0])


def run_simulation(circ: Circuit, num_amplitudes: Optional[int] = None, seed: Optional[int] = None,
                   dtype: str = "c64") -> "OrderedDict[str, complex]":
    """simulation.jl:102-133.  The reference re-contracts once per bitstring
    (:129-131); here all unique bitstrings go down in one batched call."""
    from .executor import amplitudes_for_network
    tnc = convert_to_tnc(circ)
    plan = flow_cutter_contraction_plan(tnc, time=0, hypergraph=True)
    if num_amplitudes is None and tnc.qubits > 30:
        num_amplitudes = 1000                                   # :118-120
    if num_amplitudes is None:
        amps = amplitudes_all(tnc.qubits)
    else:
        amps = amplitudes_uniform(tnc.qubits, seed, num_amplitudes)
    uniq = list(OrderedDict.fromkeys(amps))                     # unique(), :129
    vals = amplitudes_for_network(tnc, plan, uniq, dtype=dtype)
    return OrderedDict((b, complex(v)) for b, v in zip(uniq, vals))
