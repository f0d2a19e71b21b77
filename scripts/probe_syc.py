"""GPU probe: Sycamore-like 53q depth 7 (GEMM-shaped nodes) with the streaming kernel."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from qxb200.executor import Graph, init
from oracle import qx_oracle as orc
init(0)
txt, data, w = bench.build_workload("sycamore53_d7_c32")
bits = bench.synth_bits(64, 53)
for dt, gemm in (("c32", True), ("c32", False), ("c64", True), ("c64", False)):
    g = Graph.from_dsl(txt, data, dt, replan=64, replan_n_amp=64).compile(gemm=gemm)
    out = g.amplitudes(bits)
    best = 1e9
    for _ in range(3):
        t = time.time(); out = g.amplitudes(bits); best = min(best, time.time() - t)
    st = g.stats()
    print(dt, "gemm" if gemm else "stream", f"{best*1e3:.1f} ms per 64 bitstrings -> {64/best:.0f} amp/s; {st['flops']/best/1e12:.2f} TFLOP/s, {st['bytes']/best/1e9:.0f} GB/s, ws {st['workspace_bytes']/1e9:.1f} GB, mean p*2^n {np.mean(np.abs(out)**2)*2.0**53:.3f}", flush=True)
    if dt == "c64":
        ref64 = out
    else:
        out32 = out
print("c32 vs c64 rel err", np.max(np.abs(out32 - ref64)) / np.max(np.abs(ref64)))
t = time.time()
bs = ["".join("01"[b] for b in bits[0])]
cmds = orc.parse_dsl(txt)
ref = orc.amplitudes(cmds, data, bs, slice_begin=3, slice_end=4)
print("oracle one slice", time.time() - t, "s")
g = Graph.from_dsl(txt, data, "c64", replan=64, replan_n_amp=64).compile()
got = g.amplitudes(bs, 3, 4)
print("slice parity rel err", abs(got[0] - ref[0]) / abs(ref[0]))
