#!/usr/bin/env python
"""bench.py -- amplitudes/s of the QXTools contraction hot path on B200.

One "step" = one pass of the hot path over one batch of synthetic input:
`--amps` output bitstrings x ALL slices of the sliced RQC program, summed to
amplitudes (what QXContexts.execute's "Simulation" section does,
/root/reference/bin/qxrun.jl:83-87; reference timings docs/src/distributed.md:79-103).

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference ...                     # the CPU restatement (oracle/) on host cores

value : whole-job amplitudes/s, bitstrings already resident in HBM, timed with CUDA
        events on the launching stream (max over ranks)
e2e   : the same through the host-buffer C-ABI call qxb_amplitudes (pinned host
        bitstrings -> H2D -> contraction -> D2H of the amplitudes inside the timed region)
N > 1 : one process per GPU (torchrun).  Default partition: every rank takes a contiguous share of the
        BITSTRINGS over all slices (the dependency-tracked lowering makes a rank's work proportional to its
        bitstrings; fixing slice variables per rank instead -- `--partition slices`, what north_star names --
        leaves the slice-independent part of every node replicated and measured 0.44 efficiency at N = 8,
        profiles/r1_scaling.md); the shards meet in ONE NCCL all-gather of [n_amp / N] complex numbers per rank
        -> "scaling": "strong".  The gather of step i runs on NCCL's stream while step i + 1 computes into the
        other of two shard buffers; the pipeline is drained before the closing event of the timed region.
The plan and every kernel knob are the library's defaults (deterministic: the re-planner is seeded); `--autotune`
adds the measured choice among exact alternatives of round 1 (qxb200/tuning.py).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np

WORKLOADS = {
    # BASELINE.json configs[3]: RQC 7x7 depth 20, ComplexF64, 2^12 slices (the multi-GPU config)
    "rqc_7x7_d20_c64_s4096": dict(rows=7, cols=7, depth=20, n_slice=12, dtype="c64", seed=42),
    # BASELINE.json configs[2]: RQC 6x6 depth 16, ComplexF32, 2^6 slices
    "rqc_6x6_d16_c32_s64": dict(rows=6, cols=6, depth=16, n_slice=6, dtype="c32", seed=42),
    "rqc_4x4_d12_c64_s16": dict(rows=4, cols=4, depth=12, n_slice=4, dtype="c64", seed=42),
    # BASELINE.json configs[1]: QFT on 20 qubits, 1024 bitstrings, no slicing
    "qft_20_unsliced": dict(qft=20, rows=20, cols=1, n_slice=0, dtype="c64", seed=42),
    # towards configs[4]: Sycamore-like 53 qubits (fSim, extent-4 bonds), depth 7, ComplexF32, 8 sliced bonds.
    # GEMM-shaped nodes (K up to 512): compute-bound, unlike the RQC grids.  Depth 12 needs a stronger planner.
    "sycamore53_d7_c32": dict(sycamore=7, rows=53, cols=1, n_slice=8, dtype="c32", seed=1),
    # BASELINE.json configs[4]: Sycamore-like 53 qubits, 12 cycles, ComplexF32; sliced by the library's GPU-aware
    # slicing so that the largest tensor of a slice is 2^31 elements (17 GB in HBM); file written by
    # workloads/make_sycamore12.py.  A parity / capability case (scripts/probe_syc12.py), not a bench line.
    "sycamore53_d12_c32_s2048": dict(sycamore=12, rows=53, cols=1, n_slice=0, dtype="c32", seed=1),
    # same circuit, sliced and planned by the current tree search (stochastic reconfiguration + bisection seeds):
    # 16 slices, 2^38 MACs per slice -- modelled 0.06 s per slice; not yet measured (profiles/r1p_summary.md)
    "sycamore53_d12_c32_s16": dict(sycamore=12, rows=53, cols=1, n_slice=0, dtype="c32", seed=1),
}
DEFAULT_WORKLOAD = "rqc_7x7_d20_c64_s4096"


def build_workload(name):
    """Seeded synthetic file triple (the reference planner/FlowCutter cannot run here)."""
    import qxb200 as q
    w = WORKLOADS[name]
    cache = os.path.join(ROOT, "workloads", name)
    if os.path.exists(cache + ".qx") and (os.path.exists(cache + ".jld2") or os.path.exists(cache + ".npz")):
        from qxb200.jld2 import load_data_file
        txt = open(cache + ".qx").read()
        # the committed triple: .jld2 through the library's native reader (.npz = pre-JLD2 copies of the same arrays)
        data = load_data_file(cache + (".jld2" if os.path.exists(cache + ".jld2") else ".npz"))
    else:
        circ = (q.create_qft_circuit(w["qft"]) if "qft" in w else
                q.create_sycamore_like_circuit(w["sycamore"], seed=w["seed"]) if "sycamore" in w else
                q.create_rqc_circuit(w["rows"], w["cols"], w["depth"], w["seed"]))
        tnc = q.convert_to_tnc(circ)
        bg, plan, meta = q.contraction_scheme(tnc, w["n_slice"], time=0, seed=w["seed"])
        cg = q.build_compute_graph(tnc, plan, bg)
        txt, data = cg.dsl(meta), dict(cg.tensors)
    return txt, data, w


def synth_bits(n_amp, n_qubits, seed=2020):
    return np.random.default_rng(seed).integers(0, 2, (n_amp, n_qubits)).astype(np.uint8)


# ------------------------------------------------------------------ CPU arms
def _cpu_task(args):
    txt, data, bitstring, s0, s1, dtype = args
    from oracle import qx_oracle as orc
    cmds = _cpu_task.cache.get(hash(txt))
    if cmds is None:
        cmds = orc.parse_dsl(txt)
        _cpu_task.cache[hash(txt)] = cmds
    return orc.amplitude(cmds, data, bitstring, np.complex64 if dtype == "c32" else np.complex128, s0, s1)


_cpu_task.cache = {}


def cpu_sample(txt, data, w, n_bitstrings, n_slices_sample, cores):
    """Time the oracle (the CPU restatement of the reference loop) on a bounded
    sample: n_bitstrings x the first n_slices_sample slices, spread over `cores`
    processes; extrapolate linearly in the slice count (slices are independent,
    equal-cost units) to amplitudes/s for the full slice space."""
    import multiprocessing as mp
    from oracle import qx_oracle as orc
    os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
    n_q = w["rows"] * w["cols"]
    bits = synth_bits(n_bitstrings, n_q)
    bss = ["".join("01"[b] for b in row) for row in bits]
    total_slices = 1
    for _, d in orc.slice_dims(orc.parse_dsl(txt)):
        total_slices *= d
    n_slices_sample = min(n_slices_sample, total_slices)
    per = max(1, n_slices_sample // max(1, cores // max(1, n_bitstrings)))
    tasks = []
    for b in bss:
        for s0 in range(0, n_slices_sample, per):
            tasks.append((txt, data, b, s0, min(s0 + per, n_slices_sample), w["dtype"]))
    t0 = time.perf_counter()
    if cores > 1:
        with mp.get_context("fork").Pool(cores) as pool:
            pool.map(_cpu_task, tasks, chunksize=1)
    else:
        for t in tasks:
            _cpu_task(t)
    dt = time.perf_counter() - t0
    inst_per_s = n_bitstrings * n_slices_sample / dt
    amps_per_s = inst_per_s / total_slices
    sample = (f"{n_bitstrings} bitstrings x first {n_slices_sample} of {total_slices} slices "
              f"({n_bitstrings * n_slices_sample} contractions) in {dt:.1f}s, numpy {np.__version__} oracle, "
              f"{cores} processes x 1 BLAS thread; extrapolated linearly to all slices")
    return amps_per_s, dt, sample


def load_workload_numpy(name):
    """The committed file triple with numpy only (.qx text + the .npz copy of the .jld2 arrays): the reference arm must
    not map libqxb200.so."""
    cache = os.path.join(ROOT, "workloads", name)
    txt = open(cache + ".qx").read()
    with np.load(cache + ".npz") as z:
        data = {k: z[k] for k in z.files}
    return txt, data, WORKLOADS[name]


def run_reference(args):
    """--impl reference: the CPU implementation of the path on the host cores.  The
    real reference (Julia QXContexts) cannot be installed here (no Julia, no network;
    DESIGN.md), so this is the oracle port.  numpy only: no repo .so is loaded."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    txt, data, w = load_workload_numpy(args.workload)
    cores = os.cpu_count() or 1
    vals, times = [], []
    sample = ""
    n_bs = max(1, min(4, cores))
    for i in range(args.warmup + args.steps):
        v, dt, sample = cpu_sample(txt, data, w, n_bs, args.ref_slices, cores)
        if i >= args.warmup:
            vals.append(v)
            times.append(dt)
    value = float(np.mean(vals))
    line = {
        "impl": "reference", "metric": "RQC amplitudes/sec", "value": value, "unit": "amplitudes/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * float(np.mean(times)), "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64" if w["dtype"] == "c64" else "f32", "data": "synthetic",
        "config": {"workload": args.workload, "n_amp": n_bs, "sample_slices": args.ref_slices, "same_config": False,
                   "extrapolation": "each step contracts n_amp bitstrings x the first sample_slices slices of the file's own "
                                    "contraction order (one slice of one bitstring per contraction, as QXContexts does); "
                                    "amplitudes/s = contractions/s / total slices (slices are independent, equal-cost units)"},
        "cpu_baseline": {"value": value, "unit": "amplitudes/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "amplitudes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.rows.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for nme, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(max(mx)) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def tune_plan(args, txt, data, w, g, dev, stream, world, dist, torch):
    """--autotune: -> (graph to run (uncompiled), its plan text, winning compile options, report).  Candidates: (planner
    model setting) x (min_lob); the probe is the first min(amps, 32768) bitstrings over the full slice space, 2 warm-up +
    3 timed replays, CUDA events on the stream the library launches on; results must agree with the baseline (default
    plan, default knobs) to 1e-9 (ComplexF64) / 1e-4 (ComplexF32) of the largest amplitude.  The knobs travel as
    qxb_options (compile keyword arguments): nothing is left in os.environ."""
    from qxb200.executor import Graph, autotune
    n_q = w["rows"] * w["cols"]
    n_probe = int(min(args.amps, 32768))
    from qxb200.tuning import candidates_of, plan_candidates
    plans = plan_candidates(txt, data, w["dtype"], args.replan_candidates, args.amps, first=g)
    cands = candidates_of(plans)
    bits = torch.from_numpy(synth_bits(n_probe, n_q)).to(dev)
    cdt = torch.complex64 if w["dtype"] == "c32" else torch.complex128
    out = torch.zeros(n_probe, dtype=cdt, device=dev)

    def build(text, **kw):
        return Graph.from_dsl(text, data, w["dtype"]).compile(amp_batch=args.amp_batch, cuda_graph=not args.no_graph, **kw)

    def probe(gc):
        S = gc.n_slices
        for _ in range(2):
            gc.amplitudes_device(bits.data_ptr(), n_probe, out.data_ptr(), 0, S)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(3):
            gc.amplitudes_device(bits.data_ptr(), n_probe, out.data_ptr(), 0, S)
        e1.record(stream)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / 3.0, out.cpu().numpy().copy()

    def reduce_times(ts):
        if world == 1:
            return ts
        t = torch.tensor([1e30 if x == float("inf") else x for x in ts], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float("inf") if x >= 1e29 else x for x in t.tolist()]

    tol = 1e-4 if w["dtype"] == "c32" else 1e-9
    best, report = autotune(cands, build, probe, reduce_times, rel_tol=tol)
    tag, text, kw = cands[best]
    del bits, out
    info = next(i for t, _, i in plans if tag.startswith(t + "/"))
    chosen = Graph.from_dsl(text, data, w["dtype"])
    chosen.replan_info = info
    return chosen, text, kw, {"chosen": tag, "options": kw, "probe_bitstrings": n_probe, "candidates": report}


# ------------------------------------------------------------------ GPU arm
def run_gpu(args):
    import torch
    import torch.distributed as dist
    from qxb200.executor import Graph, init, set_stream

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    init(local)
    dev = torch.device("cuda", local)
    # a real (non-null) stream shared by torch (events, NCCL) and the library's launches
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    set_stream(stream.cuda_stream)

    txt, data, w = build_workload(args.workload)
    n_q = w["rows"] * w["cols"]
    # batch-aware re-planning inside the library (qxb_graph_replan: seeded, hence identical on every rank)
    g = Graph.from_dsl(txt, data, w["dtype"], replan=0 if args.no_replan else args.replan_candidates,
                       replan_n_amp=args.amps)
    plan_txt = g.text
    tune_report, knobs = None, {}
    if not args.no_replan and args.autotune:
        # Measured choice among exact alternatives before anything is timed (qxb200.executor.autotune).
        # Any failure in here leaves the run exactly as it was without tuning.
        try:
            g, plan_txt, knobs, tune_report = tune_plan(args, txt, data, w, g, dev, stream, world, dist, torch)
        except Exception as e:                                   # noqa: BLE001
            tune_report, knobs = {"error": repr(e)[:300]}, {}
    import hashlib
    plan_sha1 = hashlib.sha1(plan_txt.encode()).hexdigest()[:16]
    g.compile(amp_batch=args.amp_batch, cuda_graph=not args.no_graph, **knobs)
    S = g.n_slices
    n_amp = args.amps
    # How the ranks share a step: either each rank takes a share of the bitstrings
    # over all slices, or all ranks take all bitstrings and fix different values of the slice
    # variables the cost model picks (qxb_partition_vars) -- whichever leaves less work per rank;
    # either way the partial results meet in ONE NCCL all-reduce of [n_amp] complex numbers.
    s0, s1 = 0, S
    assign, a0, a1 = None, 0, n_amp
    if world > 1:
        mode, how = g.choose_partition(n_amp, world, rank)
        if args.partition == "slices":
            assign = g.partition_assignment(world, rank)
            mode = "slices" if assign is not None else "ranges"
            how = assign if assign is not None else ((S * rank) // world, (S * (rank + 1)) // world)
        elif args.partition == "amps":
            mode, how = "amps", ((n_amp * rank) // world, (n_amp * (rank + 1)) // world)
        if mode == "amps":
            a0, a1 = how
        elif mode == "slices":
            assign = how
        else:
            s0, s1 = how
    else:
        mode = "single"
    bits_h = torch.from_numpy(synth_bits(n_amp, n_q)).pin_memory()
    bits_d = bits_h.to(dev)
    cdt = torch.complex64 if w["dtype"] == "c32" else torch.complex128
    out_d = torch.zeros(n_amp, dtype=cdt, device=dev)
    out_h = torch.zeros(n_amp, dtype=cdt).pin_memory()
    out_h_np = out_h.numpy()
    bits_h_np = bits_h.numpy()

    es = 8 if w["dtype"] == "c32" else 16
    n_mine = a1 - a0

    # bitstring shards of equal size meet in ONE all-gather (each rank owns a disjoint [n_amp / N] range: nothing to
    # sum); slice partitions and ragged shards need the sum -> all-reduce of the zero-padded vector
    gather = world > 1 and mode == "amps" and n_amp % world == 0
    shard_d = torch.zeros(max(n_mine, 1), dtype=cdt, device=dev) if gather else None
    # the all-gather of step i runs on NCCL's stream while step i + 1 computes into the other shard buffer; a buffer is
    # reused only after its gather has completed (stream-side wait), and drain() closes the pipeline inside the timed region
    shard_alt = torch.zeros(max(n_mine, 1), dtype=cdt, device=dev) if gather else None
    pending = [None, None]
    turn = [0]

    def drain():
        for i in (0, 1):
            if pending[i] is not None:
                pending[i].wait()
                pending[i] = None

    def step_device():
        if gather:
            i = turn[0] & 1
            turn[0] += 1
            buf = shard_d if i == 0 else shard_alt
            if pending[i] is not None:
                pending[i].wait()
            g.amplitudes_device(bits_d.data_ptr() + a0 * n_q, n_mine, buf.data_ptr(), 0, S)
            pending[i] = dist.all_gather_into_tensor(torch.view_as_real(out_d), torch.view_as_real(buf), async_op=True)
            return
        if mode == "amps":
            out_d.zero_()
            if n_mine:
                g.amplitudes_device(bits_d.data_ptr() + a0 * n_q, n_mine, out_d.data_ptr() + a0 * es, 0, S)
        elif assign is not None:
            g.amplitudes_subspace_device(bits_d.data_ptr(), n_amp, out_d.data_ptr(), assign[0], assign[1])
        else:
            g.amplitudes_device(bits_d.data_ptr(), n_amp, out_d.data_ptr(), s0, s1)
        if world > 1:
            dist.all_reduce(torch.view_as_real(out_d))

    def step_e2e():
        # the public host-buffer call: pinned bitstrings -> H2D -> contraction -> D2H, synchronised
        from qxb200._lib import check
        import ctypes as C
        if gather:
            # every rank: its shard through the host-buffer call; the shards then meet on the devices (256 KB per rank)
            # and rank 0 reads the whole vector back -- what a job that writes one results file does
            check(g._lib.qxb_amplitudes(g._h, C.c_void_p(bits_h.data_ptr() + a0 * n_q), n_mine, 0, S,
                                        C.c_void_p(out_h.data_ptr() + a0 * es)))
            shard_d.copy_(out_h[a0:a1], non_blocking=True)
            dist.all_gather_into_tensor(torch.view_as_real(out_d), torch.view_as_real(shard_d))
            if rank == 0:
                out_h.copy_(out_d, non_blocking=False)
            return
        if mode == "amps":
            out_h.zero_()
            if n_mine:
                check(g._lib.qxb_amplitudes(g._h, C.c_void_p(bits_h.data_ptr() + a0 * n_q), n_mine, 0, S,
                                            C.c_void_p(out_h.data_ptr() + a0 * es)))
        elif assign is not None:
            fv = (C.c_int32 * max(len(assign[0]), 1))(*assign[0])
            fx = (C.c_int64 * max(len(assign[1]), 1))(*assign[1])
            check(g._lib.qxb_amplitudes_subspace(g._h, C.c_void_p(bits_h.data_ptr()), n_amp, fv, fx, len(assign[0]),
                                                 C.c_void_p(out_h.data_ptr()), 0))
        else:
            check(g._lib.qxb_amplitudes(g._h, C.c_void_p(bits_h.data_ptr()), n_amp, s0, s1,
                                        C.c_void_p(out_h.data_ptr())))
        if world > 1:
            t = torch.view_as_real(out_h.to(dev, non_blocking=True))
            dist.all_reduce(t)
            out_h.copy_(torch.view_as_complex(t), non_blocking=False)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step_device()
    drain()
    barrier()
    ws_bytes = g.stats()["workspace_bytes"]
    flush = ws_bytes < 4 * 126e6            # working set could sit in the 126 MB L2 -> flush between steps
    flush_buf = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device=dev) if flush else None
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    if not flush:
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(stream)
        for _ in range(args.steps):
            step_device()
        drain()                             # the last gathers complete before the closing event
        ev1.record(stream)
        barrier()
        ms = ev0.elapsed_time(ev1)
        l2_note = f"working set {ws_bytes / 1e9:.2f} GB per GPU >> 126 MB L2 (no flush needed)"
    else:
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        for a, b in evs:
            flush_buf.fill_(1)              # evicts L2; not inside the timed event pair
            a.record(stream)
            step_device()
            drain()
            b.record(stream)
        barrier()
        ms = sum(a.elapsed_time(b) for a, b in evs)
        l2_note = (f"working set {ws_bytes / 1e6:.0f} MB could fit the 126 MB L2: a 512 MB buffer is written between "
                   "steps (outside the per-step CUDA-event pairs) to flush it")
    st = g.stats()
    # e2e
    for _ in range(2):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    barrier()
    ms_e2e = 1e3 * (time.perf_counter() - t0)
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        t = torch.tensor([ms, ms_e2e], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e = float(t[0]), float(t[1])
    result = out_d.cpu().numpy()

    if rank == 0:
        ms_step = ms / args.steps
        value = n_amp / (ms_step * 1e-3)
        e2e = n_amp / (ms_e2e / args.steps * 1e-3)
        # sanity: Porter-Thomas / norm check -- mean |amp|^2 * 2^n should be ~1 for an RQC
        norm = float(np.mean(np.abs(result) ** 2) * 2.0 ** n_q)
        replan_info = g.replan_info
        if ws_bytes > 40e9:
            # a working set of this size (BASELINE configs[4]: 86 GB per slice) leaves no room for the profiled replica the
            # roofline pass builds: the timed graph is released first (its numbers are all taken by now)
            import gc
            g = None
            gc.collect()
            torch.cuda.empty_cache()
        roof = roofline(args, plan_txt, data, w, bits_d, out_d, n_amp if mode != "amps" else n_mine, s0, s1, assign,
                        bits_off=a0 * n_q if mode == "amps" else 0, knobs=knobs, plan_sha1=plan_sha1)
        as_given = None
        if world == 1 and g is not None and replan_info and replan_info.get("replanned") and not args.no_as_given:
            # the same step on the contraction order exactly as the file gives it (no re-planning)
            g0 = Graph.from_dsl(txt, data, w["dtype"]).compile(amp_batch=args.amp_batch, **knobs)
            for _ in range(3):
                g0.amplitudes_device(bits_d.data_ptr(), n_amp, out_d.data_ptr(), 0, S)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(5):
                g0.amplitudes_device(bits_d.data_ptr(), n_amp, out_d.data_ptr(), 0, S)
            e1.record(stream)
            torch.cuda.synchronize()
            t_ms = e0.elapsed_time(e1) / 5
            given = out_d.cpu().numpy()
            as_given = {"value": n_amp / (t_ms * 1e-3), "ms_per_step": t_ms, "gb_per_step": g0.stats()["bytes"] / 1e9,
                        "max_rel_diff_vs_replanned": float(np.max(np.abs(given - result)) / np.max(np.abs(given)))}
            del g0
        cpu = None
        if world == 1 and not args.no_cpu:
            cores = os.cpu_count() or 1
            v, dt, sample = cpu_sample(txt, data, w, max(1, min(4, cores)), args.ref_slices, cores)
            cpu = {"value": v, "unit": "amplitudes/s", "cores": cores, "kind": "port", "sample": sample, "same_config": False,
                   "n_amp": max(1, min(4, cores)), "sample_slices": args.ref_slices}
        es = 8 if w["dtype"] == "c32" else 16
        line = {
            "metric": "RQC amplitudes/sec", "value": value, "unit": "amplitudes/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64" if w["dtype"] == "c64" else "f32", "data": "synthetic",
            "config": {"workload": args.workload, "n_amp_per_step": n_amp, "n_slices": S,
                       "n_qubits": n_q, "complex": w["dtype"],
                       "partition": ("single GPU" if world == 1 else
                                     f"bitstrings split over {world} ranks, all slices each; shards meet in one NCCL "
                                     f"{'all-gather (asynchronous: it overlaps the next step, drained inside the timed region)' if gather else 'all-reduce'}" if mode == "amps" else
                                     f"all bitstrings on every rank, slice variables {[v + 1 for v in assign[0]]} fixed per rank; one NCCL all-reduce"
                                     if assign is not None else f"contiguous slice ranges / {world}; one NCCL all-reduce"),
                       "plan": ("re-planned for batched execution (qxb_graph_replan, exact re-association, seeded): "
                                f"{replan_info['given_bytes'] / 1e9:.2f} -> {replan_info['bytes'] / 1e9:.2f} GB per "
                                f"{replan_info['n_amp_model']} bitstrings") if replan_info and replan_info.get("replanned")
                               else "contraction order as given by the file",
                       "plan_sha1": plan_sha1,
                       "kernels": "library defaults (qxb_options all 0): block phase = one row program; chunk phase: the chain of "
                                  "dominant contractions fused into one row-program launch, the other nodes one kernel each "
                                  "(TMA ring kernel on rows >= 48 KB, TMA-staged / streaming contract kernels elsewhere)"
                                  if not knobs else f"autotuned options {knobs}",
                       "as_given_plan": as_given,
                       "autotune": tune_report,
                       "l2": l2_note,
                       "amp_batch": st["amp_batch"], "mean_p_times_2^n": norm},
            "e2e": {"value": e2e, "unit": "amplitudes/s", "h2d_bytes_per_step": int(n_amp * n_q),
                    "d2h_bytes_per_step": int(n_amp * es)},
            "gpu_launches": int(st["kernel_launches"]) * args.steps,
            "roofline": roof, "cpu_baseline": cpu, "clocks": clocks,
            "algorithmic": {"gflop_per_step": st["flops"] / 1e9, "gb_per_step": st["bytes"] / 1e9,
                            "launches_per_step": int(st["kernel_launches"])},
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def roofline(args, txt, data, w, bits_d, out_d, n_amp, s0, s1, assign=None, bits_off=0, knobs=None, plan_sha1=""):
    """Per-op CUDA-event timing of one more step (same stream, same inputs, same plan and knobs) on a profiled clone
    of the graph; the reported kernels are those of the DOMINANT contractions = top ops by FLOPs covering >= 80% of
    the step's FLOPs.  achieved = algorithmic bytes s*(|A|+|B|+|C|) / event time.
    traffic = DRAM bytes per dominant launch from the committed ncu --set full capture, used ONLY when the capture was
    taken on this very plan (profiles/ncu_traffic.json is keyed by workload and plan_sha1)."""
    from qxb200.executor import Graph
    gp = Graph.from_dsl(txt, data, w["dtype"]).compile(amp_batch=args.amp_batch, profile=True, **(knobs or {}))
    for _ in range(2):
        if assign is not None:
            gp.amplitudes_subspace_device(bits_d.data_ptr(), n_amp, out_d.data_ptr(), assign[0], assign[1])
        else:
            gp.amplitudes_device(bits_d.data_ptr() + bits_off, n_amp, out_d.data_ptr(), s0, s1)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    prof = gp.profile_dump(os.path.join(ROOT, "gpurun_out", f"op_profile_{args.workload}.json"))
    ops = [o for v in prof["variants"] for o in v["ops"]]
    tot_fl = sum(o["flops"] for o in ops) or 1.0
    ops.sort(key=lambda o: -o["flops"])
    dom, acc = [], 0.0
    for o in ops:
        dom.append(o); acc += o["flops"]
        if acc >= 0.8 * tot_fl:
            break
    by = sum(o["bytes"] for o in dom); ms = sum(o["ms"] for o in dom); fl = sum(o["flops"] for o in dom)
    launches = sum(o["launches"] for o in dom)
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, src = 6650.0, "fallback (B200_PROFILING.md)"
    achieved = by / (ms * 1e-3) / 1e9 if ms > 0 else 0.0
    traffic, traffic_note = None, None
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tpath):
        t = json.load(open(tpath)).get(args.workload)
        if t and t.get("plan_sha1") == plan_sha1 and not knobs:
            traffic = t["bytes_per_launch_mean_dominant"] * n_amp / float(t["at_amps"])
            traffic_note = f"ncu --set full of this plan ({t['source']}), {t['at_amps']} bitstrings, linear in the batch"
        elif t:
            traffic_note = (f"the committed ncu capture ({t.get('source')}) belongs to plan {t.get('plan_sha1')}, "
                            f"this run timed plan {plan_sha1}: not used")
    all_ms = sum(o["ms"] for o in ops)
    kernels = sorted({o.get("kernel", "") for o in dom})
    dominant = [{"op": o["name"], "kernel": o.get("kernel", ""), "nC": o.get("nC"), "nK": o.get("nK"), "fused": o.get("fused"),
                 "ms": o["ms"], "gbs": o["bytes"] / o["ms"] / 1e6 if o["ms"] else None,
                 "tflops": o["flops"] / o["ms"] / 1e9 if o["ms"] else None} for o in dom]
    hbm = {"achieved": achieved, "peak": peak, "peak_source": src, "unit": "GB/s", "frac": achieved / peak if peak else None,
           "what": "algorithmic bytes s*(|A|+|B|+|C|) of the dominant contractions / their CUDA-event time"}
    chain = next((o for o in dom if o["name"] == "ROWPROG_CHAIN"), None)
    if chain is None:
        return {"bound": "hbm", "kernel": f"{' + '.join(kernels)} (dominant contractions: top ops by FLOPs covering >=80%)",
                "achieved": achieved, "peak": peak, "peak_source": src, "unit": "GB/s",
                "frac": achieved / peak if peak else None, "traffic": traffic, "traffic_note": traffic_note,
                "bytes_per_launch": by / max(launches, 1), "ms_per_launch": ms / max(launches, 1),
                "dominant_ops": len(dom), "dominant_ms": ms, "all_contract_ms": all_ms, "dominant": dominant,
                "dominant_gflops": fl / (ms * 1e-3) / 1e9 if ms > 0 else 0.0,
                "all_ops_achieved": sum(o["bytes"] for o in ops) / (all_ms * 1e-3) / 1e9 if all_ms > 0 else 0.0}
    # The dominant contractions run as ONE fused launch whose intermediates never reach HBM.  The contract's roofline
    # object keeps its definition -- achieved = ALGORITHMIC bytes of the dominant contractions / their device time, against
    # the measured HBM peak -- and `traffic` (ncu DRAM bytes of the same launch) shows that the launch really moves
    # several times LESS than the algorithmic bytes: that is what fusion buys, and why `frac` may approach or pass 1.
    # What bounds the fused launch itself is the FP64 (FP32) FMA pipe: reported under "compute", peak measured live.
    from qxb200.executor import fma_peak
    pipe_peak = fma_peak(w["dtype"])
    tf = chain["flops"] / (chain["ms"] * 1e-3) / 1e12 if chain["ms"] else 0.0
    io = chain.get("io_bytes_per_row", 0.0) * n_amp
    return {"bound": "hbm",
            "kernel": f"rowprog_kernel: the {chain['fused_ops']} dominant contractions {chain['fused']} fused into one launch "
                      "(inputs staged per bitstring row, intermediates in shared memory)",
            "achieved": achieved, "peak": peak, "peak_source": src, "unit": "GB/s", "frac": achieved / peak if peak else None,
            "traffic": traffic, "traffic_note": traffic_note,
            "bytes_per_launch": by / max(launches, 1), "ms_per_launch": ms / max(launches, 1),
            "dram_bytes_model_per_launch": io,
            "note": "fused launch: algorithmic bytes s*(|A|+|B|+|C|) summed over the contractions it covers; its DRAM traffic "
                    "(traffic / dram_bytes_model_per_launch) is the per-row inputs and results only, so the launch is bound by "
                    "the FMA pipe, see compute",
            "compute": {"bound": "fp64-pipe" if w["dtype"] == "c64" else "fp32-pipe", "achieved": tf, "peak": pipe_peak,
                        "peak_source": "measured live (qxb_fma_peak: FMA loop, 2 flops per FMA)", "unit": "TFLOP/s",
                        "frac": tf / pipe_peak if pipe_peak else None,
                        "flops_per_launch": chain["flops"] / max(chain["launches"], 1),
                        "hbm_time_floor_ms": io / (peak * 1e9) * 1e3 if peak else None,
                        "pipe_time_floor_ms": chain["flops"] / max(chain["launches"], 1) / (pipe_peak * 1e12) * 1e3 if pipe_peak else None},
            "dominant_ops": len(dom), "dominant_ms": ms, "all_contract_ms": all_ms, "dominant": dominant,
            "dominant_gflops": fl / (ms * 1e-3) / 1e9 if ms > 0 else 0.0,
            "all_ops_achieved": sum(o["bytes"] for o in ops) / (all_ms * 1e-3) / 1e9 if all_ms > 0 else 0.0}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="qxb200", choices=["qxb200", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--amps", type=int, default=131072, help="bitstrings per step")
    ap.add_argument("--amp-batch", type=int, default=0)
    ap.add_argument("--ref-slices", type=int, default=64, help="slices per bitstring in the CPU sample")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-replan", action="store_true", help="run the contraction order exactly as the file gives it")
    ap.add_argument("--autotune", action="store_true",
                    help="measured choice among planner models / register-tile knobs before the timed region (round 1's default)")
    ap.add_argument("--no-autotune", action="store_true", help="accepted for compatibility (autotune is off unless --autotune)")
    ap.add_argument("--replan-candidates", type=int, default=128, help="orders scored by the re-planner (seeded)")
    ap.add_argument("--no-as-given", action="store_true", help="skip the extra as-given-plan measurement")
    ap.add_argument("--partition", default="auto", choices=["auto", "amps", "slices"])
    ap.add_argument("--no-graph", action="store_true", help="launch kernels directly (no CUDA-graph replay)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
