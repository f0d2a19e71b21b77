"""Run ONE step of a workload with serial kernel launches so that `ncu --profile-from-start off` captures exactly the
contractions named in QXB_NCU_OPS (the library brackets them with cudaProfilerStart/Stop).  Two modes:
  NCU_CASE=rqc   headline workload (RQC 7x7 d20, re-planned), PROBE_AMPS bitstrings (default 32768)
  NCU_CASE=gemm  one GEMM-shaped node M = N = K = 2^9 per bitstring, 16 bitstrings, NCU_DTYPE c32|c64 (op name: c)
Numbers printed under a profiler are not measurements."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from qxb200.executor import Graph, init
init(0)
case = os.environ.get("NCU_CASE", "rqc")
if case == "rqc":
    n_amp = int(os.environ.get("PROBE_AMPS", "32768"))
    txt, data, w = bench.build_workload("rqc_7x7_d20_c64_s4096")
    g = Graph.from_dsl(txt, data, w["dtype"], replan=128, replan_n_amp=131072).compile(cuda_graph=False)
    out = g.amplitudes(bench.synth_bits(n_amp, 49))
    print("rqc step done", np.sum(out))
else:
    dtype = os.environ.get("NCU_DTYPE", "c32")
    rng = np.random.default_rng(5)
    nm = nn = nk = 9
    M = list(range(1, nm + 1)); N = list(range(nm + 1, nm + nn + 1)); K = list(range(nm + nn + 1, nm + nn + nk + 1))
    o_lab = nm + nn + nk + 1
    la = list(rng.permutation(M + K)); lb = list(rng.permutation(N + K)) + [o_lab]
    lb2 = [x for x in lb if x != o_lab]
    lc = list(rng.permutation(M + N))
    mk = lambda n: (rng.normal(size=(2,) * n) + 1j * rng.normal(size=(2,) * n)) / 8
    j = lambda l: ",".join(str(int(i)) for i in l)
    txt = ("# version: 0.4.0\n"
           f"load a dA {j([2] * len(la))}\nload b dB {j([2] * len(lb))}\nload v dV {j([2] * len(lc))}\noutput o1 1 2\n"
           f"ncon b2 {j(lb2)} b {j(lb)} o1 {o_lab}\nncon c {j(lc)} a {j(la)} b2 {j(lb2)}\n"
           f"ncon z 0 c {j(lc)} v {j(lc)}\nsave output z\n")
    g = Graph.from_dsl(txt, {"dA": mk(len(la)), "dB": mk(len(lb)), "dV": mk(len(lc))}, dtype).compile(cuda_graph=False)
    out = g.amplitudes(["0", "1", "+", "-"] * 4)
    print("gemm step done", dtype, np.sum(out))
