"""GPU probe: ONE GEMM-shaped node (M = N = 2^PM, K = 2^PK per bitstring, shuffled mode orders, PROBE_AMPS bitstrings)
through the SIMT kernel (gemm_mode 1), the mma.sync tensor-core kernels (4) and the auto choice (2: tcgen05 / TMEM for
ComplexF32, DMMA for ComplexF64).  Per-node TFLOP/s (complex-equivalent: 8 flop per complex MAC) and the error
against ComplexF64 SIMT."""
import json, os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from qxb200.executor import Graph, init
init(0)
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
nm = nn = int(os.environ.get("PROBE_PM", "9"))
nk = int(os.environ.get("PROBE_PK", "9"))
n_amp = int(os.environ.get("PROBE_AMPS", "16"))
rng = np.random.default_rng(5)
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
data = {"dA": mk(len(la)), "dB": mk(len(lb)), "dV": mk(len(lc))}
bs = (["0", "1", "+", "-"] * ((n_amp + 3) // 4))[:n_amp]
res, ref = {}, None
for dt, mode in (("c64", 1), ("c64", 2), ("c32", 1), ("c32", 4), ("c32", 2)):
    g = Graph.from_dsl(txt, data, dt).compile(gemm_mode=mode, profile=True, row_programs=False)
    g.amplitudes(bs)
    out = g.amplitudes(bs)
    prof = g.profile_dump(os.path.join(tempfile.mkdtemp(), "p.json"))
    o = [o for v in prof["variants"] for o in v["ops"] if o["name"] == "c"][0]
    if ref is None:
        ref = out
    err = float(np.max(np.abs(out - ref)) / np.max(np.abs(ref)))
    tf = o["flops"] / o["ms"] / 1e9
    res[f"{dt}_m{mode}"] = {"kernel": o["kernel"], "ms": o["ms"], "tflops": tf, "rel_err_vs_c64_simt": err}
    print(f"{dt} gemm_mode {mode}: {o['kernel']:10s} {o['ms']:.3f} ms  {tf:7.1f} TFLOP/s complex-equivalent   err vs c64 SIMT {err:.2e}", flush=True)
json.dump({"M_bits": nm, "N_bits": nn, "K_bits": nk, "n_amp": n_amp, "results": res},
          open(os.path.join(ROOT, "gpurun_out", f"probe_gemm_node_m{nm}_k{nk}.json"), "w"), indent=1)
