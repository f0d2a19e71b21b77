"""Writes the benchmark file triples (<name>.qx / .jld2 / .yml, plus the .npz copy of the data) that bench.py loads.

Seeded and deterministic: circuits from qxb200.circuits (numpy PCG64, seed 42), plan from
contraction_scheme(time=0) = deterministic min-fill + greedy tree-trimming slicing -- the
reference's own fallback path when FlowCutter returns nothing
(/root/reference/src/contraction_planning.jl:102-106,229-241).  Run from the repo root:

    python workloads/make_workloads.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench                      # noqa: E402
import qxb200 as q                # noqa: E402

for name, w in bench.WORKLOADS.items():
    circ = (q.create_qft_circuit(w["qft"]) if "qft" in w else
            q.create_sycamore_like_circuit(w["sycamore"], seed=w["seed"]) if "sycamore" in w else
            q.create_rqc_circuit(w["rows"], w["cols"], w["depth"], w["seed"]))
    n_q = circ.num_qubits
    prefix = os.path.join(ROOT, "workloads", name)
    q.generate_simulation_files(circ, prefix, w["n_slice"], seed=w["seed"], time=0,
                                output_args=q.output_params_dict(n_q, 16, seed=2020), npz=True)
    print(name, os.path.getsize(prefix + ".qx"), "bytes of .qx")
