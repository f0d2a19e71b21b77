"""GPU probe: step time of a workload against the number of bitstrings per call, for the three execution modes --
per-op kernels only (row_programs=False), block phase as one program + per-op chunk kernels ("block"), and both
phases as row programs ("all").  Finds the crossover the auto mode uses (qxb_options.row_chunk_max_amps) and gives
the latency-regime numbers (the reference's default is 10 amplitudes per run, src/outputs.jl:48)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
from qxb200.executor import Graph, init
init(0)
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
wl = os.environ.get("PROBE_WORKLOAD", "rqc_7x7_d20_c64_s4096")
sizes = [int(x) for x in os.environ.get("PROBE_SIZES", "1,10,128,1024,4096,8192,16384,32768,131072").split(",")]
txt, data, w = bench.build_workload(wl)
nq = w["rows"] * w["cols"]
nmax = max(sizes)
bits = torch.from_numpy(bench.synth_bits(nmax, nq)).cuda()
cdt = torch.complex64 if w["dtype"] == "c32" else torch.complex128
out = torch.zeros(nmax, dtype=cdt, device="cuda")
plan_txt = Graph.from_dsl(txt, data, w["dtype"], replan=128, replan_n_amp=131072).text
res = {}
ref = {}
for mode in (False, "block", "all"):
    g = Graph.from_dsl(plan_txt, data, w["dtype"]).compile(row_programs=mode)
    S = g.n_slices
    for n in sizes:
        for _ in range(3):
            g.amplitudes_device(bits.data_ptr(), n, out.data_ptr(), 0, S)
        torch.cuda.synchronize()
        reps = 20 if n <= 8192 else 5
        t0 = time.perf_counter()
        for _ in range(reps):
            g.amplitudes_device(bits.data_ptr(), n, out.data_ptr(), 0, S)
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) * 1e3 / reps
        r = out[:n].cpu().numpy().copy()
        if mode is False:
            ref[n] = r
        err = float(np.max(np.abs(r - ref[n])) / np.max(np.abs(ref[n])))
        res.setdefault(str(mode), {})[n] = {"ms": ms, "us_per_amp": ms * 1e3 / n, "launches": g.stats()["kernel_launches"], "rel_diff": err}
        print(f"[{mode}] n_amp {n:7d}: {ms:8.3f} ms  {ms * 1e3 / n:10.3f} us/amp  launches {g.stats()['kernel_launches']:4d}  diff {err:.1e}", flush=True)
    del g
json.dump({"workload": wl, "results": res}, open(os.path.join(ROOT, "gpurun_out", f"probe_batch_{wl}.json"), "w"), indent=1)
