"""One step of the headline workload through the library's DEFAULT path (per-op chunk phase with the fused chain),
serial launches (no CUDA graph) so that ncu sees every kernel.  Prints the plan hash the capture belongs to.
Numbers printed under a profiler are not measurements."""
import hashlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from qxb200.executor import Graph, init
init(0)
wl = os.environ.get("PROBE_WORKLOAD", "rqc_7x7_d20_c64_s4096")
n_amp = int(os.environ.get("PROBE_AMPS", "32768"))
txt, data, w = bench.build_workload(wl)
g = Graph.from_dsl(txt, data, w["dtype"], replan=128, replan_n_amp=131072)
print("plan_sha1", hashlib.sha1(g.text.encode()).hexdigest()[:16], flush=True)
g.compile(cuda_graph=False)
bits = bench.synth_bits(n_amp, w["rows"] * w["cols"])
for _ in range(int(os.environ.get("NCU_CALLS", "2"))):
    out = g.amplitudes(bits)
print("step done", np.sum(out), g.stats())
