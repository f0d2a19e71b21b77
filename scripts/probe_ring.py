"""GPU probe: headline workload, per-op chunk path with and without the TMA ring kernel (and its tile knob);
prints step time, the dominant ops with the kernel that ran them, and the difference against ring=False."""
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
plan_txt = Graph.from_dsl(txt, data, w["dtype"], replan=128, replan_n_amp=n_amp).text
CONFIGS = [("nochain", dict(chain=False), {}), ("default", {}, {}), ("chain_dmma", dict(row_dmma=True), {}),
           ("chain_tt8", {}, {"QXB_CHAIN_MIN_TT": "8"})]
only = os.environ.get("PROBE_ONLY")
if only:
    CONFIGS = [c for c in CONFIGS if c[0] in only.split(",")]
ref, results = None, {}
for tag, kw, env in CONFIGS:
    for k in ("QXB_RING_MIN_TT", "QXB_RING_MIN_ROW_BYTES", "QXB_CHAIN_MIN_TT", "QXB_CHAIN_MIN_MACS"):
        os.environ.pop(k, None)
    os.environ.update(env)
    g = Graph.from_dsl(plan_txt, data, w["dtype"]).compile(**kw)
    S = g.n_slices
    for _ in range(3):
        g.amplitudes_device(bits.data_ptr(), n_amp, out.data_ptr(), 0, S)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        g.amplitudes_device(bits.data_ptr(), n_amp, out.data_ptr(), 0, S)
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) * 1e3 / 5
    st = g.stats()
    res = out.cpu().numpy()
    if ref is None:
        ref = res
    err = float(np.max(np.abs(res - ref)) / np.max(np.abs(ref)))
    del g
    gp = Graph.from_dsl(plan_txt, data, w["dtype"]).compile(profile=True, **kw)
    for _ in range(2):
        gp.amplitudes_device(bits.data_ptr(), n_amp, out.data_ptr(), 0, S)
    prof = gp.profile_dump(os.path.join(ROOT, "gpurun_out", f"op_profile_ring_{tag}.json"))
    del gp
    ops = [o for v in prof["variants"] for o in v["ops"] if o["launches"]]
    ops.sort(key=lambda o: -o["flops"])
    tf, acc, dom = sum(o["flops"] for o in ops), 0.0, []
    for o in ops:
        dom.append(o); acc += o["flops"]
        if acc >= 0.8 * tf:
            break
    dby, dms = sum(o["bytes"] for o in dom), sum(o["ms"] for o in dom)
    print(f"[{tag}] {ms:.3f} ms per {n_amp} -> {n_amp / ms * 1e3:.3e} amp/s, {st['bytes'] / ms / 1e6:.0f} GB/s alg; dominant {dby / dms / 1e6:.0f} GB/s "
          f"({dby / dms / 1e6 / 6451.5:.3f} of HBM peak); diff vs first {err:.1e}", flush=True)
    print("    " + "  ".join(f"{o['name']}[{o.get('nC', '')},{o.get('nK', '')}]{o.get('kernel', '')} {o['ms']:.2f}ms {o['bytes'] / o['ms'] / 1e6:.0f}GB/s {o['flops'] / o['ms'] / 1e9:.1f}TF" for o in dom), flush=True)
    for o in ops:
        if o["name"] == "ROWPROG_CHAIN":
            print(f"    chain: fused {o['fused']} levels {o['levels']} units {o['units']} arena {o['arena_bytes']} B, io {o['io_bytes_per_row']:.0f} B/row, "
                  f"{o['ms']:.3f} ms, {o['flops'] / o['ms'] / 1e9:.2f} TFLOP/s, algorithmic {o['bytes'] / o['ms'] / 1e6:.0f} GB/s", flush=True)
            rows_cta0 = 2 * ((n_amp + 295) // 296)
            print("    cycles per level and row (CTA 0):", [round(c / rows_cta0) for c in o.get("level_cycles_cta0", [])], flush=True)
    nring = sum(1 for o in ops if o.get("kernel") == "ring")
    print(f"    ring nodes: {nring} of {len(ops)}; sum of op ms {sum(o['ms'] for o in ops):.2f}", flush=True)
    results[tag] = {"ms": ms, "dominant_gbs": dby / dms / 1e6, "rel_diff": err, "ring_nodes": nring}
json.dump(results, open(os.path.join(ROOT, "gpurun_out", f"probe_ring_{wl}.json"), "w"), indent=1)
