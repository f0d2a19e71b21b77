"""One step of a workload through the row-program path, for `ncu -k regex:rowprog_kernel` (the first rowprog launch is the
single-CTA block program, the second the persistent chunk kernel).  Numbers printed under a profiler are not measurements."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from qxb200.executor import Graph, init
init(0)
wl = os.environ.get("PROBE_WORKLOAD", "rqc_7x7_d20_c64_s4096")
n_amp = int(os.environ.get("PROBE_AMPS", "131072"))
txt, data, w = bench.build_workload(wl)
kw = {}
for k in ("row_min_tt_bits", "row_tile_regs", "row_ctas_per_sm"):
    if os.environ.get("NCU_" + k.upper()):
        kw[k] = int(os.environ["NCU_" + k.upper()])
g = Graph.from_dsl(txt, data, w["dtype"], replan=128, replan_n_amp=131072).compile(cuda_graph=False, **kw)
out = g.amplitudes(bench.synth_bits(n_amp, w["rows"] * w["cols"]))
print("step done", np.sum(out), g.stats())
