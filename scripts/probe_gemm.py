"""GPU probe: SIMT FMA GEMM (gemm_mode 1) against the tensor-core GEMM (gemm_mode 2: DMMA / 3xTF32)
on the Sycamore-like 53q depth 7 workload (GEMM-shaped nodes).  Prints whole-step times and the
per-node TFLOP/s of the GEMM nodes; writes gpurun_out/probe_gemm.json."""
import sys, time, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
from qxb200.executor import Graph, init
init(0)
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
txt, data, w = bench.build_workload("sycamore53_d7_c32")
n_amp = int(os.environ.get("PROBE_AMPS", "64"))
bits = bench.synth_bits(n_amp, 53)
res = {}
outs = {}
DTYPES = os.environ.get("PROBE_DTYPES", "c32,c64").split(",")
MODES = [int(m) for m in os.environ.get("PROBE_MODES", "1,2").split(",")]
NOPROF = os.environ.get("PROBE_NOPROF", "") != ""          # under ncu: one warm-up step + one step, nothing else
for dt in DTYPES:
    for mode in MODES:
        g = Graph.from_dsl(txt, data, dt, replan=64, replan_n_amp=64).compile(gemm_mode=mode)
        out = g.amplitudes(bits)
        best = 1e9
        for _ in range(1 if NOPROF else 4):
            torch.cuda.synchronize(); t = time.time(); out = g.amplitudes(bits); best = min(best, time.time() - t)
        st = g.stats()
        outs[(dt, mode)] = out
        if NOPROF:
            print(dt, "mode", mode, f"{best*1e3:.2f} ms (under a profiler: not a measurement)", flush=True)
            continue
        gp = Graph.from_dsl(txt, data, dt, replan=64, replan_n_amp=64).compile(gemm_mode=mode, profile=True)
        gp.amplitudes(bits); gp.amplitudes(bits)
        prof = gp.profile_dump(os.path.join(ROOT, "gpurun_out", f"probe_gemm_{dt}_m{mode}.json"))
        ops = [o for v in prof["variants"] for o in v["ops"]]
        ops.sort(key=lambda o: -o["flops"])
        top = [{"name": o["name"], "nC": o["nC"], "nK": o["nK"], "m": o["m_bits"], "n": o["n_bits"],
                "ms": o["ms"] / max(o["launches"], 1), "tflops": o["flops"] / max(o["launches"], 1) / (o["ms"] / max(o["launches"], 1)) / 1e9}
               for o in ops[:10]]
        res[f"{dt}_m{mode}"] = {"ms_per_step": best * 1e3, "amp_per_s": n_amp / best, "tflops": st["flops"] / best / 1e12, "top": top}
        print(dt, "mode", mode, f"{best*1e3:.2f} ms per {n_amp} bitstrings -> {n_amp/best:.0f} amp/s; {st['flops']/best/1e12:.2f} TFLOP/s whole step", flush=True)
        for o in top[:8]:
            print("    ", o["name"], f"nC {o['nC']} nK {o['nK']} m {o['m']} n {o['n']}: {o['ms']:.3f} ms {o['tflops']:.1f} TFLOP/s", flush=True)
if NOPROF or ("c64", 1) not in outs:
    sys.exit(0)
ref = outs[("c64", 1)]
sc = np.max(np.abs(ref))
for k, v in outs.items():
    e = float(np.max(np.abs(v - ref)) / sc)
    res[f"{k[0]}_m{k[1]}"]["rel_err_vs_c64_simt"] = e
    print(k, "rel err vs c64 SIMT", e, flush=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "probe_gemm.json"), "w"), indent=1)
