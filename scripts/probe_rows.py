"""GPU probe: the headline workload (RQC 7x7 d20, 2^12 slices batched) through the row-program path under several
settings of its knobs, against the per-op kernels.  Prints ms per step, amplitudes/s, the difference of the amplitudes
against the per-op result, and the profile of the fused launches."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
from qxb200.executor import Graph, init
init(0)
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
wl = os.environ.get("PROBE_WORKLOAD", "rqc_7x7_d20_c64_s4096")
n_amp = int(os.environ.get("PROBE_AMPS", "131072"))
txt, data, w = bench.build_workload(wl)
nq = w["rows"] * w["cols"]
bits = torch.from_numpy(bench.synth_bits(n_amp, nq)).cuda()
cdt = torch.complex64 if w["dtype"] == "c32" else torch.complex128
out = torch.zeros(n_amp, dtype=cdt, device="cuda")
g0 = Graph.from_dsl(txt, data, w["dtype"], replan=128, replan_n_amp=n_amp)
plan_txt = g0.text
del g0
CONFIGS = [("per_op", dict(row_programs=False)), ("rows", {}), ("rows_tt8", dict(row_min_tt_bits=8)), ("rows_tt6", dict(row_min_tt_bits=6)),
           ("rows_cta1", dict(row_ctas_per_sm=1)), ("rows_regs64", dict(row_tile_regs=64)), ("rows_tt8_regs64", dict(row_min_tt_bits=8, row_tile_regs=64))]
only = os.environ.get("PROBE_ONLY")
if only:
    CONFIGS = [c for c in CONFIGS if c[0] in only.split(",")]
results, ref = {}, None
for tag, kw in CONFIGS:
    g = Graph.from_dsl(plan_txt, data, w["dtype"]).compile(**kw)
    S = g.n_slices
    for _ in range(3):
        g.amplitudes_device(bits.data_ptr(), n_amp, out.data_ptr(), 0, S)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    reps = 5
    for _ in range(reps):
        g.amplitudes_device(bits.data_ptr(), n_amp, out.data_ptr(), 0, S)
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) * 1e3 / reps
    st = g.stats()
    res = out.cpu().numpy()
    if ref is None:
        ref = res
    err = float(np.max(np.abs(res - ref)) / np.max(np.abs(ref)))
    del g
    gp = Graph.from_dsl(plan_txt, data, w["dtype"]).compile(profile=True, **kw)
    for _ in range(2):
        gp.amplitudes_device(bits.data_ptr(), n_amp, out.data_ptr(), 0, S)
    prof = gp.profile_dump(os.path.join(ROOT, "gpurun_out", f"op_profile_rows_{tag}.json"))
    del gp
    fusedops = [o for v in prof["variants"] for o in v["ops"] if o["name"].startswith("ROWPROG")]
    print(f"[{tag}] {ms:.3f} ms per {n_amp} bitstrings -> {n_amp / ms * 1e3:.3e} amp/s, launches {st['kernel_launches']}, "
          f"{st['flops'] / ms / 1e9:.2f} TFLOP/s, {st['bytes'] / ms / 1e6:.0f} GB/s algorithmic, max rel diff vs per_op {err:.2e}", flush=True)
    for o in fusedops:
        print(f"    {o['name']}: {o['ms']:.3f} ms, {o['fused_ops']} ops, {o['levels']} levels, {o['units']} units, arena {o['arena_bytes']} B, "
              f"{o['flops'] / max(o['ms'], 1e-9) / 1e9:.2f} TFLOP/s", flush=True)
    results[tag] = {"kw": kw, "ms_per_step": ms, "amp_per_s": n_amp / ms * 1e3, "tflops": st["flops"] / ms / 1e9, "rel_diff_vs_per_op": err,
                    "fused": fusedops}
json.dump({"workload": wl, "n_amp": n_amp, "results": results}, open(os.path.join(ROOT, "gpurun_out", f"probe_rows_{wl}.json"), "w"), indent=1)
