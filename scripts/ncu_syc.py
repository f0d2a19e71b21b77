"""One slice of the Sycamore-53 depth-12 workload (BASELINE.json configs[4]) with serial launches, for ncu.
Numbers printed under a profiler are not measurements."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from qxb200.executor import Graph, init
init(0)
txt, data, w = bench.build_workload(os.environ.get("PROBE_WORKLOAD", "sycamore53_d12_c32_s16"))
g = Graph.from_dsl(txt, data, "c32").compile(hbm_budget_bytes=int(150e9), cuda_graph=False)
bits = bench.synth_bits(1, 53)
out = g.amplitudes(bits, 0, 1)
print("slice done", out)
