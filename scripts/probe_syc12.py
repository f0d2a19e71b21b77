"""GPU probe for BASELINE.json configs[4]: Sycamore-like 53 qubits, 12 cycles, ComplexF32, 2048 slices with a
2^31-element (17 GB) largest tensor per slice.  Runs a few slices, checks ComplexF32 against ComplexF64 on the
same slices, prints the per-node profile of one slice and projects the whole-amplitude time.
Writes gpurun_out/probe_syc12.json."""
import gc, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
from qxb200.executor import Graph, init
init(0)
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NAME = os.environ.get("PROBE_WORKLOAD", "sycamore53_d12_c32_s16")
txt, data, w = bench.build_workload(NAME)
bits = bench.synth_bits(1, 53)
n_sl = min(int(os.environ.get("PROBE_SLICES", "4")), 16)
res = {"workload": NAME}
t = time.time()
g = Graph.from_dsl(txt, data, "c32").compile(hbm_budget_bytes=int(150e9))
S = g.n_slices
res["n_slices"] = S
out = g.amplitudes(bits, 0, 1)                     # compiles the variant, folds constants
torch.cuda.synchronize()
res["compile_plus_first_slice_s"] = time.time() - t
t = time.time(); a32 = g.amplitudes(bits, 0, n_sl); torch.cuda.synchronize(); dt = time.time() - t
st = g.stats()
res.update({"slices_timed": n_sl, "s_per_slice": dt / n_sl, "tflops": st["flops"] / dt / 1e12, "gb_per_s": st["bytes"] / dt / 1e9,
            "workspace_gb": st["workspace_bytes"] / 1e9, "n_blocks": st["n_blocks"], "kernel_launches": st["kernel_launches"],
            "projected_s_per_amplitude_1gpu": dt / n_sl * S, "projected_amp_per_s_8gpu": 8.0 / (dt / n_sl * S)})
print(f"c32: {n_sl} slices in {dt:.3f} s -> {dt / n_sl * 1e3:.1f} ms per slice, {st['flops'] / dt / 1e12:.1f} TFLOP/s, "
      f"{st['bytes'] / dt / 1e9:.0f} GB/s, workspace {st['workspace_bytes'] / 1e9:.1f} GB, blocks {st['n_blocks']}; "
      f"whole amplitude ({S} slices): {dt / n_sl * S:.0f} s on one GPU", flush=True)
first32 = g.amplitudes(bits, 0, 2)
del g; gc.collect(); torch.cuda.empty_cache()
if not os.environ.get("PROBE_NO_PROFILE"):
    gp = Graph.from_dsl(txt, data, "c32").compile(hbm_budget_bytes=int(150e9), profile=True)
    gp.amplitudes(bits, 0, 1); gp.amplitudes(bits, 0, 1)
    prof = gp.profile_dump(os.path.join(ROOT, "gpurun_out", "op_profile_syc12.json"))
    ops = [o for v in prof["variants"] for o in v["ops"] if o["launches"]]
    ops.sort(key=lambda o: -o["ms"])
    tot = sum(o["ms"] / o["launches"] for o in ops)
    res["top_ops"] = []
    for o in ops[:12]:
        ms = o["ms"] / o["launches"]; fl = o["flops"] / o["launches"]; by = o["bytes"] / o["launches"]
        res["top_ops"].append({"name": o["name"], "nC": o["nC"], "nK": o["nK"], "m": o["m_bits"], "n": o["n_bits"], "ms": ms,
                               "tflops": fl / ms / 1e9, "gb_s": by / ms / 1e6, "share": ms / tot, "kernel": o.get("kernel", "")})
        print(f"    {o['name']} [{o.get('kernel', '')}]: nC {o['nC']} nK {o['nK']} m {o['m_bits']} n {o['n_bits']}: {ms:.2f} ms ({ms / tot:.0%}) "
              f"{fl / ms / 1e9:.1f} TFLOP/s {by / ms / 1e6:.0f} GB/s", flush=True)
    del gp; gc.collect(); torch.cuda.empty_cache()
if not os.environ.get("PROBE_NO_C64"):
    g64 = Graph.from_dsl(txt, data, "c64").compile(hbm_budget_bytes=int(165e9))
    t = time.time(); first64 = g64.amplitudes(bits, 0, 2); torch.cuda.synchronize()
    res["c64_2_slices_s"] = time.time() - t
    err = float(np.max(np.abs(first32 - first64)) / np.max(np.abs(first64)))
    res["c32_vs_c64_rel_err_2_slices"] = err
    print("c32 vs c64 on slices [0, 2): rel err", err, "values", first64, flush=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "probe_syc12.json"), "w"), indent=1)
