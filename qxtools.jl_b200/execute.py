"""``execute``: run a simulation file triple -- the entry point behind
``bin/qxrun.jl`` (/root/reference/bin/qxrun.jl:66-97), i.e. QXContexts.execute
(call site :83-87) with the same keyword set.

    execute(dsl_file, input_file=None, param_file=None, output_file=None;
            use_mpi=False, sub_comm_size=1, use_gpu=True,
            max_amplitudes=None, max_slices=None, timings=False)

* ``param_file`` defaults to the DSL name with ``.yml``, ``input_file`` to the DSL
  name with ``.jld2`` (qxrun.jl:21,25), read by the library's native JLD2 reader
  (``csrc/qxb_jld2.cpp``); the harness' older ``.npz`` data files are used when no
  ``.jld2`` sits next to the DSL file.  ``output_file``: ``.jld2`` (datasets
  ``bitstrings`` and ``amplitudes``) unless the name ends in ``.npz``.
* ``max_amplitudes`` / ``max_slices`` keep the FIRST N of the parameter file /
  slice space (qxrun.jl:32-39).
* ``use_mpi`` -> torch.distributed (one process per GPU, launched by torchrun);
  ``sub_comm_size`` ranks share the slices of a bitstring group (qxrun.jl:40-46).
* ``use_gpu=False`` is an error: this executor has no CPU path.
"""
from __future__ import annotations

import argparse
import os
import time
from collections import OrderedDict
from typing import Optional

import numpy as np
import yaml

from .dist import Distribution
from .jld2 import load_data_file, save_jld2
from .simulation import amplitudes_uniform


def write_results(output_file: str, bitstrings, amplitudes, **extra) -> None:
    """The results file ``bin/qxrun.jl -o`` names (docs/src/distributed.md:30-33): JLD2 with the
    datasets ``bitstrings`` (fixed-length strings) and ``amplitudes``; ``.npz`` on request."""
    if output_file.endswith(".npz"):
        np.savez(output_file, bitstrings=np.array(bitstrings), amplitudes=np.asarray(amplitudes), **extra)
        return
    n_q = max([len(b) for b in bitstrings], default=1)
    arrays = {"bitstrings": np.array([b.encode() for b in bitstrings], dtype=f"S{max(1, n_q)}"),
              "amplitudes": np.asarray(amplitudes)}
    arrays.update({k: np.asarray(v, dtype=np.float64) for k, v in extra.items()})
    save_jld2(output_file, arrays, commit_types=False)


def _bitstrings_from_params(params: dict, max_amplitudes: Optional[int], param_file: Optional[str] = None):
    """List / Uniform give the bitstrings up front; Rejection returns None (sampled on the fly).
    Uniform with a parameter FILE draws through the library (``qxb_params_read``: its documented splitmix64 stream), so
    that ``python -m qxb200.execute``, ``qxrun`` and ``QXB200.execute`` give the same bitstrings for the same file and seed."""
    out = params["output"]
    method, p = out["method"], out["params"]
    if method == "List":                                  # outputs.jl:63-68
        bs = list(p["bitstrings"])
    elif method == "Uniform" and param_file is not None:  # outputs.jl:69-72, simulation.jl:24-28
        from .jld2 import read_params
        bs = list(read_params(param_file)["bitstrings"])
    elif method == "Uniform":
        bs = list(amplitudes_uniform(int(p["num_qubits"]), p.get("seed"), int(p["num_samples"])))
    elif method == "Rejection":                           # outputs.jl:57-62
        return None
    else:
        raise ValueError(f"output method {method!r} not supported")
    if max_amplitudes is not None:
        bs = bs[:max_amplitudes]
    return bs


def execute(dsl_file: str, input_file: Optional[str] = None, param_file: Optional[str] = None,
            output_file: Optional[str] = None, use_mpi: bool = False, sub_comm_size: int = 1,
            use_gpu: bool = True, max_amplitudes: Optional[int] = None, max_slices: Optional[int] = None,
            timings: bool = False, dtype: str = "c32", replan: int = -1, autotune: bool = False):
    """Returns ``OrderedDict{bitstring => amplitude}`` on rank 0 (None elsewhere).
    ``autotune=True`` (single process, List / Uniform): measured choice among the re-planner's trees and the
    register-tile knob on the first bitstrings before the run (``qxb200.tuning.tune``)."""
    if not use_gpu:
        raise RuntimeError("qxb200 has no CPU path: use_gpu must be True")
    from .executor import Graph, init
    stem = os.path.splitext(dsl_file)[0]
    param_file = param_file or stem + ".yml"
    if input_file is None:
        input_file = stem + ".jld2" if os.path.exists(stem + ".jld2") or not os.path.exists(stem + ".npz") else stem + ".npz"
    t = OrderedDict()
    t0 = time.perf_counter()
    text = open(dsl_file).read()
    data = load_data_file(input_file)
    params = yaml.safe_load(open(param_file))
    t["Parse input files"] = time.perf_counter() - t0

    t0 = time.perf_counter()
    rank, world = 0, 1
    td = None
    if use_mpi:
        import torch
        import torch.distributed as td
        if not td.is_initialized():
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            os.environ.setdefault("MASTER_PORT", "29533")
            td.init_process_group("nccl" if torch.cuda.is_available() else "gloo")
        rank, world = td.get_rank(), td.get_world_size()
    init(int(os.environ.get("LOCAL_RANK", "0")))
    bitstrings = _bitstrings_from_params(params, max_amplitudes, param_file)
    n_model = len(bitstrings) if bitstrings is not None else 1024
    tune_report = None
    if autotune and not use_mpi and bitstrings:
        from .executor import bits_from_strings
        from .tuning import tune
        g0 = Graph.from_dsl(text, data, dtype)
        probe = bits_from_strings(bitstrings[:min(len(bitstrings), 32768)], g0.n_outputs)
        del g0
        g, tune_report = tune(text, data, dtype, np.ascontiguousarray(probe), replan_candidates=32 if replan < 0 else max(1, replan))
        g.compile(**tune_report.get("options", {}))
    else:
        g = Graph.from_dsl(text, data, dtype, replan=replan, replan_n_amp=max(1, n_model)).compile()
    if g.root_dims:
        raise ValueError("the program saves a tensor (open network); execute() writes one amplitude per bitstring -- "
                         "use Graph.amplitudes / contract_tn for open networks")
    t["Create Context"] = time.perf_counter() - t0

    t0 = time.perf_counter()
    n_slices = g.n_slices if max_slices is None else min(max_slices, g.n_slices)
    if bitstrings is None:
        # Rejection sampling: every candidate batch is one amplitudes call on this rank's GPU
        # (rank 0 only; the sampler is sequential in its acceptance bound M)
        from .samplers import rejection_sample
        p = params["output"]["params"]
        n_samp = int(p["num_samples"]) if max_amplitudes is None else min(int(p["num_samples"]), max_amplitudes)
        t["Create sampler"] = time.perf_counter() - t0
        t0 = time.perf_counter()
        results = None
        if rank == 0:
            bs, amps, info = rejection_sample(lambda bits: g.amplitudes(bits, 0, n_slices), int(p["num_qubits"]), n_samp,
                                              M=float(p.get("M", 0.0001)), fix_M=bool(p.get("fix_M", False)),
                                              seed=p.get("seed"))
            t["Simulation"] = time.perf_counter() - t0
            results = OrderedDict(zip(bs, amps))
            if output_file:
                write_results(output_file, bs, np.array(amps), M=info["M"], drawn=info["drawn"])
            if timings:
                for k, v in t.items():
                    print(f"  {k:<20s} {v * 1e3:10.3f} ms")
        return results
    d = Distribution(len(bitstrings), n_slices, world, rank, sub_comm_size if use_mpi else 1)
    t["Create sampler"] = time.perf_counter() - t0

    t0 = time.perf_counter()
    mine = bitstrings[d.amp_begin:d.amp_end]
    # inside a sub-communicator the slice space is split by FIXING the slice variables the cost
    # model picks (they carry the work); contiguous ranges when -n truncates the slice space
    assign = None
    if d.sub_comm_size > 1 and n_slices == g.n_slices:
        assign = g.partition_assignment(d.sub_comm_size, d.rank_in_group)
    if not mine:
        part = np.zeros(0, g.np_dtype)
    elif assign is not None:
        part = g.amplitudes_subspace(mine, assign[0], assign[1])
    else:
        part = g.amplitudes(mine, d.slice_begin, d.slice_end)
    full = np.zeros(len(bitstrings), dtype=g.np_dtype)
    full[d.amp_begin:d.amp_end] = part
    if use_mpi and world > 1:
        import torch
        dev = torch.device("cuda", torch.cuda.current_device()) if td.get_backend() == "nccl" else torch.device("cpu")
        buf = torch.from_numpy(full).to(dev)
        td.all_reduce(torch.view_as_real(buf))       # slices summed inside a group, groups are disjoint in bitstrings
        full = buf.cpu().numpy()
    t["Simulation"] = time.perf_counter() - t0

    results = None
    if rank == 0:
        t0 = time.perf_counter()
        results = OrderedDict((b, complex(a)) for b, a in zip(bitstrings, full))
        if output_file:
            write_results(output_file, bitstrings, full)
        t["Write results"] = time.perf_counter() - t0
        if timings:
            for k, v in t.items():
                print(f"  {k:<20s} {v * 1e3:10.3f} ms")
    return results


def main(argv=None):
    """CLI with the flags of bin/qxrun.jl:15-57."""
    ap = argparse.ArgumentParser("qxrun (qxb200)")
    ap.add_argument("--dsl", "-d", required=True, help="DSL file path")
    ap.add_argument("--parameter-file", "-p", default=None)
    ap.add_argument("--input-file", "-i", default=None)
    ap.add_argument("--output-file", "-o", required=True)
    ap.add_argument("--number-amplitudes", "-a", type=int, default=None)
    ap.add_argument("--number-slices", "-n", type=int, default=None)
    ap.add_argument("--sub-comm-size", "-s", type=int, default=1)
    ap.add_argument("--mpi", "-m", action="store_true")
    ap.add_argument("--gpu", "-g", action="store_true", help="accepted for compatibility; the GPU is always used")
    ap.add_argument("--timings", "-t", action="store_true")
    ap.add_argument("--blas-threads", "-b", type=int, default=8, help="ignored (no BLAS on the path)")
    ap.add_argument("--dtype", default="c32", choices=["c32", "c64"])
    ap.add_argument("--autotune", action="store_true", help="measured choice among re-planned trees / register-tile knobs first")
    a = ap.parse_args(argv)
    kw = dict(use_mpi=a.mpi, sub_comm_size=a.sub_comm_size, use_gpu=True, max_amplitudes=a.number_amplitudes,
              max_slices=a.number_slices, timings=a.timings, dtype=a.dtype, autotune=a.autotune)
    res = execute(a.dsl, a.input_file, a.parameter_file, a.output_file, **kw)
    if a.timings:                 # qxrun.jl:89-96 runs twice so the second table excludes warm-up
        res = execute(a.dsl, a.input_file, a.parameter_file, a.output_file, **kw)
    return res


if __name__ == "__main__":
    main()
