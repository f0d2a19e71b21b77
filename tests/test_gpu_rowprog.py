"""GPU parity of the row-program path (csrc/qxb_rowprog.cu: one persistent kernel takes every bitstring row through
the whole chunk phase in shared memory; the block phase is a single-CTA program) against the numpy oracle and
against the per-op kernels (row_programs=False) on the same inputs.  Tolerances are BASELINE.json's."""
import numpy as np
import pytest

from qxb200.executor import Graph
from oracle import qx_oracle as orc
from cases import kat0, rqc_case, circuit_case, random_program, rel_err
import qxb200 as q

pytestmark = pytest.mark.gpu
TOL = {"c64": 1e-10, "c32": 1e-5}


def fused(g):
    """True when the last call ran as row programs: a handful of launches per slice block instead of one per ncon."""
    st = g.stats()
    return st["kernel_launches"] <= 2 + 2 * st["n_blocks"]


def fused_block(g):
    """True when the block phase of the last (single-block) call ran as ONE program: launches = 1 + the chunk-phase
    contractions + output leaves, root reduction, finalize."""
    ops = g.describe()["ops"]
    n_block = sum(1 for o in ops if o["phase"] == "block")
    n_chunk = sum(1 for o in ops if o["phase"] == "chunk")
    return n_block < 2 or g.stats()["kernel_launches"] <= n_chunk + 4


@pytest.mark.parametrize("dtype", ["c64", "c32"])
@pytest.mark.parametrize("shape", [(3, 3, 8, 2), (3, 4, 10, 3), (4, 4, 12, 4)])
def test_rows_vs_oracle_and_per_op(gpu, dtype, shape):
    r, c, d, ns = shape
    txt, data, bs = rqc_case(r, c, d, ns)
    bs = bs + ["+" * (r * c), "-" * (r * c)]
    cmds = orc.parse_dsl(txt)
    ref = orc.amplitudes(cmds, data, bs)
    g = Graph.from_dsl(txt, data, dtype).compile()
    got = g.amplitudes(bs)
    assert fused(g), g.stats()
    assert rel_err(got, ref, r * c) < TOL[dtype]
    g0 = Graph.from_dsl(txt, data, dtype).compile(row_programs=False)
    got0 = g0.amplitudes(bs)
    assert not fused(g0)
    assert rel_err(got0, ref, r * c) < TOL[dtype]
    S = g.n_slices
    for b, e in ((1, S - 1), (S // 2, S), (3, 4)):
        ref2 = orc.amplitudes(cmds, data, bs, slice_begin=b, slice_end=e)
        assert rel_err(g.amplitudes(bs, b, e), ref2, r * c) < TOL[dtype]


def test_kat0_rows(gpu):
    txt, data = kat0()
    g = Graph.from_dsl(txt, data, "c64").compile()
    cmds = orc.parse_dsl(txt)
    bs = ["00", "11", "01", "10"]
    for b in range(4):
        for e in range(b, 5):
            ref = orc.amplitudes(cmds, data, bs, slice_begin=b, slice_end=e)
            assert rel_err(g.amplitudes(bs, b, e), ref, 2) < 1e-10
    assert fused(g)


@pytest.mark.parametrize("dtype", ["c64", "c32"])
def test_replanned_rows(gpu, dtype):
    txt, data, bs = rqc_case(4, 4, 12, 4, n_amp=16)
    ref = orc.amplitudes(orc.parse_dsl(txt), data, bs)
    g = Graph.from_dsl(txt, data, dtype, replan=16, replan_n_amp=1024).compile()
    assert rel_err(g.amplitudes(bs), ref, 16) < TOL[dtype]
    assert fused(g)


@pytest.mark.parametrize("seed", range(6))
def test_random_programs_rows(gpu, seed):
    txt, data, bs = random_program(seed)
    cmds = orc.parse_dsl(txt)
    ref = orc.amplitudes(cmds, data, bs)
    g = Graph.from_dsl(txt, data, "c64").compile()
    scale = max(1.0, float(np.max(np.abs(ref))))
    assert np.max(np.abs(g.amplitudes(bs) - ref)) < 1e-10 * scale
    S = g.n_slices
    if S > 2:
        ref = orc.amplitudes(cmds, data, bs, slice_begin=1, slice_end=S - 1)
        assert np.max(np.abs(g.amplitudes(bs, 1, S - 1) - ref)) < 1e-10 * scale


def test_many_rows_per_cta(gpu):
    """More bitstrings than resident CTAs: every CTA walks several rows and rewrites its arena."""
    txt, data, _ = rqc_case(3, 4, 10, 3)
    n = 148 * 2 * 3 + 5
    bits = np.random.default_rng(3).integers(0, 2, (n, 12)).astype(np.uint8)
    g = Graph.from_dsl(txt, data, "c64").compile(row_programs="all")
    got = g.amplitudes(bits)
    assert fused(g)
    ga = Graph.from_dsl(txt, data, "c64").compile()          # auto: above row_chunk_max_amps the chunk phase runs per op
    got_auto = ga.amplitudes(bits)
    assert not fused(ga) and fused_block(ga)
    g0 = Graph.from_dsl(txt, data, "c64").compile(row_programs=False)
    ref = g0.amplitudes(bits)
    assert np.max(np.abs(got_auto - ref)) < 1e-12 * np.max(np.abs(ref))
    assert np.max(np.abs(got - ref)) < 1e-12 * np.max(np.abs(ref))
    idx = [0, 1, 295, 296, 297, n - 1]
    bs = ["".join("01"[b] for b in bits[i]) for i in idx]
    assert rel_err(got[idx], orc.amplitudes(orc.parse_dsl(txt), data, bs), 12) < 1e-10
    assert abs(float(np.sum(np.abs(got) ** 2)) * 2.0 ** 12 / n - 1) < 0.2        # Porter-Thomas mean


@pytest.mark.parametrize("kw", [dict(cuda_graph=False), dict(profile=True), dict(row_min_tt_bits=8), dict(row_min_tt_bits=6),
                                dict(row_tile_regs=64), dict(row_ctas_per_sm=1), dict(row_dmma=True), dict(row_bank_opt=False)])
def test_row_options_agree(gpu, kw):
    txt, data, bs = rqc_case(4, 4, 12, 4, n_amp=32)
    ref = Graph.from_dsl(txt, data, "c64", replan=16, replan_n_amp=1024).compile(row_programs=False).amplitudes(bs)
    g = Graph.from_dsl(txt, data, "c64", replan=16, replan_n_amp=1024).compile(**kw)
    assert np.max(np.abs(g.amplitudes(bs) - ref)) < 1e-12 * np.max(np.abs(ref))
    assert fused(g)


@pytest.mark.parametrize("workload,dtype,tol", [("rqc_7x7_d20_c64_s4096", "c64", 1e-11), ("rqc_6x6_d16_c32_s64", "c32", 1e-5),
                                                ("rqc_6x6_d16_c32_s64", "c64", 1e-11)])
def test_bench_workloads_rows_equal_per_op(gpu, workload, dtype, tol):
    """Full size: the plan the bench runs, all slices, 300 bitstrings; rows vs per-op kernels vs (short range) oracle."""
    import bench
    txt, data, w = bench.build_workload(workload)
    nq = w["rows"] * w["cols"]
    bits = bench.synth_bits(300, nq)
    g = Graph.from_dsl(txt, data, dtype, replan=32, replan_n_amp=131072)
    plan = g.text
    g.compile()
    got = g.amplitudes(bits)
    assert fused(g), g.stats()
    ref = Graph.from_dsl(plan, data, "c64").compile(row_programs=False).amplitudes(bits)
    assert np.max(np.abs(got - ref)) < tol * max(np.max(np.abs(ref)), 2.0 ** (-nq / 2))
    bs = ["".join("01"[b] for b in row) for row in bits[:2]]
    o = orc.amplitudes(orc.parse_dsl(plan), data, bs, slice_begin=5, slice_end=8)
    assert rel_err(g.amplitudes(bits[:2], 5, 8), o, nq) < max(tol, 1e-10)
    assert 0.3 < float(np.mean(np.abs(got) ** 2)) * 2.0 ** nq < 3.0          # Porter-Thomas scale (300 samples)


def test_ghz_qft_rows(gpu):
    for circ, ns in ((q.create_ghz_circuit(5), 0), (q.create_qft_circuit(8), 0), (q.create_ghz_circuit(4), 2)):
        txt, data, bs = circuit_case(circ, n_slice=ns, n_amp=8)
        ref = orc.amplitudes(orc.parse_dsl(txt), data, bs)
        g = Graph.from_dsl(txt, data, "c64").compile()
        assert rel_err(g.amplitudes(bs), ref, circ.num_qubits) < 1e-10


@pytest.mark.parametrize("workload,dtype,tol", [("rqc_7x7_d20_c64_s4096", "c64", 1e-12), ("rqc_7x7_d20_c64_s4096", "c32", 1e-5),
                                                ("rqc_6x6_d16_c32_s64", "c64", 1e-12)])
def test_ring_kernel_equals_streaming_kernels(gpu, workload, dtype, tol):
    """The TMA ring kernel (bulk copies in, bulk store out) on the dominant nodes of the bench plans against the
    streaming kernels (ring=False) on the same plan, enough bitstrings that every CTA walks several rows; a short
    slice range against the oracle."""
    import bench
    txt, data, w = bench.build_workload(workload)
    nq = w["rows"] * w["cols"]
    n = 148 * 5 + 3
    bits = bench.synth_bits(n, nq)
    g = Graph.from_dsl(txt, data, dtype, replan=32, replan_n_amp=131072)
    plan = g.text
    g.compile(row_programs="block", chain=False)         # (the fused chain would take the ring-eligible nodes)
    got = g.amplitudes(bits)
    ref = Graph.from_dsl(plan, data, "c64").compile(row_programs=False, ring=False).amplitudes(bits)
    assert np.max(np.abs(got - ref)) < tol * max(np.max(np.abs(ref)), 2.0 ** (-nq / 2))
    prof = Graph.from_dsl(plan, data, dtype).compile(row_programs="block", chain=False, profile=True)
    prof.amplitudes(bits)
    import json, tempfile, os
    path = os.path.join(tempfile.mkdtemp(), "p.json")
    ops = [o for v in prof.profile_dump(path)["variants"] for o in v["ops"]]
    if dtype == "c64" and workload.startswith("rqc_7x7"):      # rows of >= 48 KB (2^11 ComplexF64 elements in and out): ring-eligible
        assert any(o.get("kernel") == "ring" for o in ops), "no node ran on the ring kernel"
    bs = ["".join("01"[b] for b in row) for row in bits[:2]]
    o = orc.amplitudes(orc.parse_dsl(plan), data, bs, slice_begin=5, slice_end=8)
    big = np.tile(bits[:2], (148, 1))                    # >= 148 rows so that the ring kernel is eligible
    assert rel_err(g.amplitudes(big, 5, 8)[:2], o, nq) < max(tol, 1e-10)


@pytest.mark.parametrize("dtype,tol", [("c64", 1e-12), ("c32", 1e-5)])
def test_fused_chain_equals_per_op(gpu, dtype, tol):
    """The chain of dominant contractions of the headline plan as ONE launch (inputs staged per row from global memory,
    intermediates in shared memory, result back to global) against the same plan with every node as its own kernel."""
    import bench, os, tempfile
    txt, data, w = bench.build_workload("rqc_7x7_d20_c64_s4096")
    n = 148 * 4 + 7
    bits = bench.synth_bits(n, 49)
    g = Graph.from_dsl(txt, data, dtype, replan=32, replan_n_amp=131072)
    plan = g.text
    got = g.compile(row_programs="block").amplitudes(bits)
    ref = Graph.from_dsl(plan, data, "c64").compile(row_programs=False).amplitudes(bits)
    assert np.max(np.abs(got - ref)) < tol * max(np.max(np.abs(ref)), 2.0 ** (-49 / 2))
    prof = Graph.from_dsl(plan, data, dtype).compile(row_programs="block", profile=True)
    got_p = prof.amplitudes(bits)                         # serial launches (no CUDA graph): same numbers
    assert np.max(np.abs(got_p - got)) <= 1e-15 + 1e-6 * (dtype == "c32")
    ops = [o for v in prof.profile_dump(os.path.join(tempfile.mkdtemp(), "p.json"))["variants"] for o in v["ops"]]
    chain = [o for o in ops if o["name"] == "ROWPROG_CHAIN"]
    assert chain and chain[0]["fused_ops"] >= 2, "the dominant chain was not fused"
    off = Graph.from_dsl(plan, data, dtype).compile(row_programs="block", chain=False).amplitudes(bits)
    assert np.max(np.abs(off - ref)) < tol * max(np.max(np.abs(ref)), 2.0 ** (-49 / 2))
    # the same chain with its big nodes on the FP64 tensor pipe (8 x 8 DMMA tiles; ComplexF64 only, a no-op for c32)
    tc = Graph.from_dsl(plan, data, dtype).compile(row_programs="block", row_dmma=True).amplitudes(bits)
    assert np.max(np.abs(tc - ref)) < tol * max(np.max(np.abs(ref)), 2.0 ** (-49 / 2))
    # lanes as lowered (no bank-aware lane bits, no re-laid-out intermediates), and the chain with side branches in its levels
    for kw in (dict(row_bank_opt=False), dict(chain_side=1), dict(chain_side=2, row_bank_opt=False)):
        alt = Graph.from_dsl(plan, data, dtype).compile(row_programs="block", **kw).amplitudes(bits)
        assert np.max(np.abs(alt - ref)) < tol * max(np.max(np.abs(ref)), 2.0 ** (-49 / 2)), kw
    # a slice sub-range (blocks with fixed variables: other variants, other chains) against the oracle
    bs = ["".join("01"[b] for b in row) for row in bits[:2]]
    o = orc.amplitudes(orc.parse_dsl(plan), data, bs, slice_begin=5, slice_end=8)
    big = np.tile(bits[:2], (160, 1))
    assert rel_err(g.amplitudes(big, 5, 8)[:2], o, 49) < max(tol, 1e-10)


def test_hoisted_block_phase_is_refreshed(gpu):
    """The block phase of a one-block step lives outside the captured graph and is re-run only when the block arena holds
    another block's results.  Alternate full-range calls (all slices batched: no fixed value), single slices (every variable
    fixed, a different value each time), a repeated slice and a multi-block range on ONE graph, replaying each captured step:
    every call must match the oracle for ITS slices."""
    txt, data, bs = rqc_case(4, 4, 12, 3, n_amp=8)
    cmds = orc.parse_dsl(txt)
    g = Graph.from_dsl(txt, data, "c64").compile()
    S = g.n_slices
    assert S == 8
    calls = [(0, S), (3, 4), (0, S), (5, 6), (3, 4), (3, 4), (1, 7), (5, 6), (0, S), (0, S), (2, 3)]
    want = {r: orc.amplitudes(cmds, data, bs, slice_begin=r[0], slice_end=r[1]) for r in set(calls)}
    for r in calls:
        got = g.amplitudes(bs, r[0], r[1])
        assert rel_err(got, want[r], 16) < 1e-10, r
