"""Writes workloads/sycamore53_d12_c32_s16.{qx,jld2,npz,yml}: BASELINE.json configs[4], a Sycamore-like 53-qubit
12-cycle fSim circuit, sliced by the LIBRARY's GPU-aware slicing (qxb_graph_replan_ex, n_free = -3) so that
the largest tensor of one slice is 2^31 ComplexF32 elements (17 GB) -- "largest intermediate held in HBM".

(workloads/sycamore53_d12_c32_s2048.* is the same circuit as sliced and planned by the first version of the tree
search -- greedy + deterministic subtree reconfiguration only: 2^51 flops unsliced, 11 indices sliced; it is the
file the measured numbers in profiles/r1p_summary.md refer to.  With the stochastic reconfiguration sweeps and
the recursive-bisection seeds the search reaches 2^44.4 flops unsliced and needs 4 sliced indices.)

The reference cannot produce this file (circuits.jl builds CZ grids only, contraction_scheme slices by
treewidth alone, contraction_planning.jl:219-299).  Seeded and deterministic.  Run from the repo root:

    python workloads/make_sycamore12.py
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import qxb200 as q                                # noqa: E402
from qxb200.executor import Graph                 # noqa: E402
from qxb200.simulation import output_params_dict, _plain  # noqa: E402

NAME = "sycamore53_d12_c32_s16"
BUDGET = int(96e9)           # 3 x 8 B x 2^32 would be 103 GB: the slicer stops at 2^31-element tensors

t = time.time()
circ = q.create_sycamore_like_circuit(12, seed=1)
tnc = q.convert_to_tnc(circ)
plan = q.min_fill_contraction_plan(tnc)           # any valid order: the library re-plans it
cg = q.build_compute_graph(tnc, plan, None)
txt, data = cg.dsl(), dict(cg.tensors)
g = Graph.from_dsl(txt, data, "c32")
info = g.replan(128, 1, n_free=-3, budget_bytes=BUDGET)
assert info["replanned"], info
g2 = Graph.from_dsl(g.text, data, "c32")
assert g2.n_slices == 16, g2.n_slices
print(f"{NAME}: {g2.n_slices} slices {g2.slice_dims}, modelled {info['model_seconds_per_block']:.3f} s per slice, "
      f"{time.time() - t:.0f} s to plan")
prefix = os.path.join(ROOT, "workloads", NAME)
header = ("# version: 0.4.0\n"
          f"# Sycamore-like 53 qubits, 12 cycles, seed 1; sliced and planned by qxb_graph_replan_ex(candidates=128, n_amp=1, "
          f"n_free=-3, budget={BUDGET})\n")
body = "\n".join(ln for ln in g.text.splitlines() if not ln.startswith("#"))
with open(prefix + ".qx", "w") as f:
    f.write(header + body + "\n")
np.savez(prefix + ".npz", **data)
from qxb200.jld2 import save_jld2  # noqa: E402
save_jld2(prefix + ".jld2", {k: np.asarray(v, dtype=np.complex128) for k, v in data.items()})
import yaml                                       # noqa: E402
with open(prefix + ".yml", "w") as f:
    yaml.safe_dump({"output": _plain(output_params_dict(53, 16, seed=2020))}, f, sort_keys=False)
