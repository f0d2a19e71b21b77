"""GPU probe: whole-step time and per-node profile of the headline workload (RQC 7x7 d20) under several settings of
the kernel-selection knobs (QXB_MINB, QXB_KC_REGS_*, QXB_SMEM_*; read by the library when a graph is lowered / its nodes are built).
Writes gpurun_out/op_profile_7x7_<tag>.json per setting and gpurun_out/probe_variants.json."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
from qxb200.executor import Graph, init
init(0)
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
n_amp = int(os.environ.get("PROBE_AMPS", "131072"))
txt, data, w = bench.build_workload("rqc_7x7_d20_c64_s4096")
bits = torch.from_numpy(bench.synth_bits(n_amp, 49)).cuda()
out = torch.zeros(n_amp, dtype=torch.complex128, device="cuda")
g0 = Graph.from_dsl(txt, data, w["dtype"], replan=128, replan_n_amp=n_amp)
plan_txt = g0.text
del g0
CONFIGS = [("default", {}),
           ("minb3", {"QXB_MINB": "3"}),
           ("minb3_kc128", {"QXB_MINB": "3", "QXB_KC_REGS_MULTI": "128", "QXB_KC_REGS_ONE": "96"}),
           ("kc128", {"QXB_KC_REGS_MULTI": "128", "QXB_KC_REGS_ONE": "96"})]
if os.environ.get("PROBE_CONFIGS") == "smem":
    CONFIGS = [("default", {}),
               ("stage_more", {"QXB_SMEM_RATIO": "0.3", "QXB_SMEM_MINHB": "0"}),
               ("stage_more_shared", {"QXB_SMEM_RATIO": "0.3", "QXB_SMEM_MINHB": "0", "QXB_SMEM_SHARED": "1"}),
               ("shared_only", {"QXB_SMEM_SHARED": "1"})]
if os.environ.get("PROBE_CONFIGS") == "lob":
    # register tiles for nodes with <= 2^8..2^10 elements per bitstring: the CTA's 256 threads span several bitstring rows
    CONFIGS = [("default", {}), ("lob7", {"QXB_MIN_LOB": "7"}), ("lob6", {"QXB_MIN_LOB": "6"}), ("lob5", {"QXB_MIN_LOB": "5"})]
if os.environ.get("PROBE_CONFIGS") == "tma":
    # operand rows through 1-D TMA bulk copies into a ring of shared-memory stages (contract_tma_kernel, never run before r2)
    CONFIGS = [("default", {}), ("tma", {"QXB_SMEM_TMA": "1"}), ("tma_all", {"QXB_SMEM_TMA": "1", "QXB_SMEM_TMA_RATIO": "0"}),
               ("tma_lob6", {"QXB_SMEM_TMA": "1", "QXB_MIN_LOB": "6"})]
KNOBS = ("QXB_SMEM_TMA", "QXB_SMEM_TMA_RATIO", "QXB_MIN_LOB", "QXB_SMEM_RATIO", "QXB_SMEM_MINHB", "QXB_SMEM_SHARED", "QXB_MINB", "QXB_KC_REGS_MULTI", "QXB_KC_REGS_ONE")
results, ref = {}, None
for tag, env in CONFIGS:
    for k in KNOBS:
        os.environ.pop(k, None)
    os.environ.update(env)
    g = Graph.from_dsl(plan_txt, data, w["dtype"]).compile()
    S = g.n_slices
    for _ in range(3):
        g.amplitudes_device(bits.data_ptr(), n_amp, out.data_ptr(), 0, S)
    torch.cuda.synchronize()
    t0 = time.perf_counter()               # the library launches on its own stream: time with a device-wide sync
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
    gp = Graph.from_dsl(plan_txt, data, w["dtype"]).compile(profile=True)
    for _ in range(2):
        gp.amplitudes_device(bits.data_ptr(), n_amp, out.data_ptr(), 0, S)
    prof = gp.profile_dump(os.path.join(ROOT, "gpurun_out", f"op_profile_7x7_{tag}.json"))
    del gp
    ops = [o for v in prof["variants"] for o in v["ops"] if o["launches"]]
    ops.sort(key=lambda o: -o["ms"])
    print(f"[{tag}] {ms:.3f} ms per {n_amp} bitstrings -> {n_amp / ms * 1e3:.3e} amp/s, {st['bytes'] / ms / 1e6:.0f} GB/s algorithmic, "
          f"max rel diff vs default {err:.2e}", flush=True)
    print("   top: " + "  ".join(f"{o['name']} {o['ms'] / o['launches']:.2f}" for o in ops[:12]), flush=True)
    results[tag] = {"env": env, "ms_per_step": ms, "amp_per_s": n_amp / ms * 1e3, "gb_s": st["bytes"] / ms / 1e6, "rel_diff_vs_default": err}
json.dump({"n_amp": n_amp, "results": results}, open(os.path.join(ROOT, "gpurun_out", "probe_variants.json"), "w"), indent=1)
