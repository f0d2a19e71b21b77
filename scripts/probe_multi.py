"""Several GPUs behind the C ABI (qxb_multi_*), one process: (1) BASELINE.json configs[4] -- Sycamore-like 53 qubits,
12 cycles, ComplexF32, 16 slices with a 2^30-element largest tensor -- one amplitude at a time, the slices split over
N = 1, 2, 4, 8 devices; (2) the headline workload, 131072 bitstrings per call split over the devices, through host buffers."""
import gc, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
from qxb200.executor import Graph, MultiGraph, init
init(0)
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ndev = torch.cuda.device_count()
res = {"devices": ndev, "sycamore53_d12_c32_s16": {}, "rqc_7x7_d20_c64_s4096": {}}
txt, data, w = bench.build_workload("sycamore53_d12_c32_s16")
bits = bench.synth_bits(1, 53)
ref = None
for N in (1, 2, 4, 8):
    if N > ndev:
        break
    m = MultiGraph(Graph.from_dsl(txt, data, "c32"), n_devices=N, hbm_budget_bytes=int(150e9))
    a = m.amplitudes(bits)                      # first call: arenas, constant folding, step graphs on every device
    reps = int(os.environ.get("PROBE_MULTI_REPS", "2"))
    t0 = time.perf_counter()
    for _ in range(reps):
        a = m.amplitudes(bits)
    dt = (time.perf_counter() - t0) / reps
    if ref is None:
        ref = a
    err = float(np.max(np.abs(a - ref)) / np.max(np.abs(ref)))
    res["sycamore53_d12_c32_s16"][N] = {"s_per_amplitude": dt, "amplitudes_per_s": 1.0 / dt, "rel_diff_vs_1gpu": err, "value": [float(a[0].real), float(a[0].imag)]}
    print(f"sycamore53 d12 c32, 16 slices over {N} GPU(s): {dt:.3f} s per amplitude -> {1.0 / dt:.3f} amplitudes/s, diff vs 1 GPU {err:.1e}", flush=True)
    del m; gc.collect()
    for d in range(ndev):
        with torch.cuda.device(d):
            torch.cuda.empty_cache()
if os.environ.get("PROBE_MULTI_ONLY") == "syc":
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "probe_multi.json"), "w"), indent=1)
    sys.exit(0)
txt, data, w = bench.build_workload("rqc_7x7_d20_c64_s4096")
n_amp = 131072
bits = bench.synth_bits(n_amp, 49)
g = Graph.from_dsl(txt, data, "c64", replan=128, replan_n_amp=n_amp)
plan = g.text
ref = None
for N in (1, 2, 4, 8):
    if N > ndev:
        break
    m = MultiGraph(Graph.from_dsl(plan, data, "c64"), n_devices=N)
    for _ in range(2):
        a = m.amplitudes(bits)
    reps = 5
    t0 = time.perf_counter()
    for _ in range(reps):
        a = m.amplitudes(bits)
    dt = (time.perf_counter() - t0) / reps
    if ref is None:
        ref = a
    err = float(np.max(np.abs(a - ref)) / np.max(np.abs(ref)))
    res["rqc_7x7_d20_c64_s4096"][N] = {"ms_per_call": dt * 1e3, "amplitudes_per_s": n_amp / dt, "rel_diff_vs_1gpu": err}
    print(f"rqc 7x7 d20 c64, {n_amp} bitstrings over {N} GPU(s) (host buffers, one process): {dt * 1e3:.2f} ms -> {n_amp / dt:.4g} amplitudes/s, diff vs 1 GPU {err:.1e}", flush=True)
    del m; gc.collect()
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "probe_multi.json"), "w"), indent=1)
