"""Parity at BASELINE.json's full sizes.  The oracle needs ~35 ms per (bitstring, slice)
contraction, so single slices of the full-size programs are checked against it
directly; whole amplitudes (4096 slices x many bitstrings) are checked through
size-independent properties: partitions of the slice space sum to the whole,
ComplexF32 agrees with ComplexF64, QFT moduli are exactly 2^{-n/2}."""
import numpy as np
import pytest

import qxb200 as q
from qxb200.executor import Graph
from oracle import qx_oracle as orc
from cases import rel_err
import bench

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def rqc77(gpu):
    txt, data, w = bench.build_workload("rqc_7x7_d20_c64_s4096")
    return txt, data, Graph.from_dsl(txt, data, "c64").compile()


def test_rqc_7x7_single_slices_vs_oracle(rqc77):
    txt, data, g = rqc77
    cmds = orc.parse_dsl(txt)
    bits = bench.synth_bits(3, 49)
    bss = ["".join("01"[b] for b in row) for row in bits]
    assert g.n_slices == 4096 and g.n_outputs == 49
    for s in (0, 1, 2049, 4095):
        ref = orc.amplitudes(cmds, data, bss, slice_begin=s, slice_end=s + 1)
        assert rel_err(g.amplitudes(bss, s, s + 1), ref, 49) < 1e-10
    # a short unaligned range (several aligned blocks)
    ref = orc.amplitudes(cmds, data, bss[:1], slice_begin=1021, slice_end=1030)
    assert rel_err(g.amplitudes(bss[:1], 1021, 1030), ref, 49) < 1e-10


def test_rqc_7x7_partitions_sum_to_whole(rqc77):
    txt, data, g = rqc77
    bits = bench.synth_bits(256, 49)
    full = g.amplitudes(bits)
    # contiguous halves / quarters of the linear slice id
    parts = sum(g.amplitudes(bits, b, b + 1024) for b in range(0, 4096, 1024))
    assert rel_err(parts, full, 49) < 1e-10
    # the multi-GPU partition: fixed slice variables chosen by the cost model
    tot = np.zeros_like(full)
    for r in range(8):
        fv, fx = g.partition_assignment(8, r)
        tot += g.amplitudes_subspace(bits, fv, fx)
    assert rel_err(tot, full, 49) < 1e-10
    # Porter-Thomas: mean probability of random bitstrings of a deep RQC is ~2^-n
    assert 0.7 < np.mean(np.abs(full) ** 2) * 2.0 ** 49 < 1.4


def test_rqc_7x7_c32_matches_c64(rqc77):
    txt, data, g = rqc77
    bits = bench.synth_bits(64, 49)
    g32 = Graph.from_dsl(txt, data, "c32").compile()
    assert rel_err(g32.amplitudes(bits), g.amplitudes(bits), 49) < 1e-5


def test_rqc_7x7_batching_is_invisible(rqc77):
    """Same amplitudes whether 300 bitstrings go down in one batch or in chunks of 37."""
    txt, data, g = rqc77
    bits = bench.synth_bits(300, 49)
    gb = Graph.from_dsl(txt, data, "c64").compile(amp_batch=37)
    assert rel_err(gb.amplitudes(bits), g.amplitudes(bits), 49) < 1e-12


@pytest.mark.parametrize("dtype,tol", [("c64", 1e-10), ("c32", 1e-5)])
def test_qft20_moduli(gpu, dtype, tol):
    """BASELINE.json configs[1]: QFT-20, 1024 bitstrings, no slicing: every amplitude has modulus 2^-10."""
    txt, data, w = bench.build_workload("qft_20_unsliced")
    g = Graph.from_dsl(txt, data, dtype).compile()
    bits = bench.synth_bits(1024, 20)
    out = g.amplitudes(bits)
    assert np.max(np.abs(np.abs(out) - 2.0 ** -10)) / 2.0 ** -10 < tol
    cmds = orc.parse_dsl(txt)
    bss = ["".join("01"[b] for b in row) for row in bits[:3]]
    assert rel_err(out[:3], orc.amplitudes(cmds, data, bss), 20) < tol


def test_rqc_7x7_replanned_matches_given_plan(rqc77):
    """The re-planned program (what bench.py times by default) gives the amplitudes of the
    program as written, on the full-size workload, and single slices still match the oracle."""
    txt, data, g = rqc77
    bits = bench.synth_bits(128, 49)
    gr = Graph.from_dsl(txt, data, "c64", replan=12).compile()
    assert gr.replan_info["replanned"] and gr.replan_info["bytes"] < 0.5 * gr.replan_info["given_bytes"]
    assert rel_err(gr.amplitudes(bits), g.amplitudes(bits), 49) < 1e-10
    cmds = orc.parse_dsl(txt)
    bss = ["".join("01"[b] for b in row) for row in bits[:2]]
    for s in (7, 4000):
        assert rel_err(gr.amplitudes(bss, s, s + 1), orc.amplitudes(cmds, data, bss, slice_begin=s, slice_end=s + 1), 49) < 1e-10


def test_smem_staged_kernel_matches(rqc77):
    """Broadcast-type nodes of the re-planned program run through the shared-memory-staged
    kernel; switching it off must not change the amplitudes."""
    txt, data, g = rqc77
    bits = bench.synth_bits(512, 49)
    a = Graph.from_dsl(txt, data, "c64", replan=12).compile()
    b = Graph.from_dsl(txt, data, "c64", replan=12).compile(smem_stage=False)
    assert rel_err(a.amplitudes(bits), b.amplitudes(bits), 49) < 1e-12
    a32 = Graph.from_dsl(txt, data, "c32", replan=12).compile()
    assert rel_err(a32.amplitudes(bits), b.amplitudes(bits), 49) < 1e-5


def test_gemm_kernel_on_sycamore_like(gpu):
    """GEMM-shaped nodes (Sycamore-like 53 qubits, depth 7: K up to 2^9) run through the tiled GEMM
    kernel; it must agree with the streaming kernel and with the oracle on single slices."""
    txt, data, w = bench.build_workload("sycamore53_d7_c32")
    bits = bench.synth_bits(6, 53)
    a = Graph.from_dsl(txt, data, "c64", replan=32, replan_n_amp=64).compile()
    b = Graph.from_dsl(txt, data, "c64", replan=32, replan_n_amp=64).compile(gemm=False)
    ga, gb = a.amplitudes(bits), b.amplitudes(bits)
    assert rel_err(ga, gb, 53) < 1e-12
    a32 = Graph.from_dsl(txt, data, "c32", replan=32, replan_n_amp=64).compile()
    assert rel_err(a32.amplitudes(bits), gb, 53) < 1e-5
    cmds = orc.parse_dsl(txt)
    bss = ["".join("01"[x] for x in row) for row in bits[:2]]
    for s in (0, 4095):
        ref = orc.amplitudes(cmds, data, bss, slice_begin=s, slice_end=s + 1)
        assert rel_err(a.amplitudes(bss, s, s + 1), ref, 53) < 1e-10
    # the program as given (no re-planning) through the GEMM path as well
    c = Graph.from_dsl(txt, data, "c64").compile()
    assert rel_err(c.amplitudes(bits[:2]), gb[:2], 53) < 1e-10


def test_tensor_core_gemm_on_sycamore_like(gpu):
    """Same workload through the tensor-core GEMM kernels (gemm_mode 2): DMMA for ComplexF64 must agree
    with the SIMT kernel to fp64 rounding, 3xTF32 for ComplexF32 within the ComplexF32 tolerance."""
    txt, data, w = bench.build_workload("sycamore53_d7_c32")
    bits = bench.synth_bits(6, 53)
    ref = Graph.from_dsl(txt, data, "c64", replan=32, replan_n_amp=64).compile(gemm_mode=1).amplitudes(bits)
    t64 = Graph.from_dsl(txt, data, "c64", replan=32, replan_n_amp=64).compile(gemm_mode=2)
    assert rel_err(t64.amplitudes(bits), ref, 53) < 1e-12
    t32 = Graph.from_dsl(txt, data, "c32", replan=32, replan_n_amp=64).compile(gemm_mode=2)      # tcgen05 where 2^7 x 2^6 tiles fit
    assert rel_err(t32.amplitudes(bits), ref, 53) < 1e-5
    m32 = Graph.from_dsl(txt, data, "c32", replan=32, replan_n_amp=64).compile(gemm_mode=4)      # mma.sync 3xTF32 only
    assert rel_err(m32.amplitudes(bits), ref, 53) < 1e-5
    cmds = orc.parse_dsl(txt)
    bss = ["".join("01"[x] for x in row) for row in bits[:1]]
    refs = orc.amplitudes(cmds, data, bss, slice_begin=17, slice_end=18)
    assert rel_err(t64.amplitudes(bss, 17, 18), refs, 53) < 1e-10
    assert rel_err(t32.amplitudes(bss, 17, 18), refs, 53) < 1e-5


def _oracle_range(args):
    txt, data, bitstring, s0, s1 = args
    from oracle import qx_oracle as o
    return o.amplitude(o.parse_dsl(txt), data, bitstring, np.complex128, s0, s1)


def test_rqc_7x7_one_whole_amplitude_vs_oracle(gpu):
    """ONE complete amplitude of the full-size program -- all 4096 slices -- against the oracle (~2.4 minutes of numpy on
    one core, spread over the host cores), through every execution mode of the library: the default of the bench (the
    re-planned tree, fused chain, ring / TMA kernels; 2 x 148 copies of the bitstring so that the batched kernels are the
    ones that run), the per-op kernels on the program as given, and the row program (3 launches per call)."""
    import multiprocessing as mp
    import os
    txt, data, w = bench.build_workload("rqc_7x7_d20_c64_s4096")
    bits = bench.synth_bits(1, 49, seed=77)
    bs = "".join("01"[b] for b in bits[0])
    cores = max(1, min(os.cpu_count() or 1, 32))
    S = 4096
    step = (S + cores - 1) // cores
    tasks = [(txt, data, bs, s0, min(s0 + step, S)) for s0 in range(0, S, step)]
    with mp.get_context("fork").Pool(cores) as pool:
        ref = sum(pool.map(_oracle_range, tasks, chunksize=1))
    scale = max(abs(ref), 2.0 ** (-49 / 2))
    g = Graph.from_dsl(txt, data, "c64", replan=128, replan_n_amp=131072).compile()
    many = np.tile(bits, (2 * 148 + 5, 1))
    got = g.amplitudes(many)
    assert np.max(np.abs(got - ref)) / scale < 1e-10
    assert abs(g.amplitudes(bits)[0] - ref) / scale < 1e-10                         # one bitstring: the row program
    as_given = Graph.from_dsl(txt, data, "c64").compile(row_programs=False, chain=False, ring=False)
    assert abs(as_given.amplitudes(bits)[0] - ref) / scale < 1e-10
    g32 = Graph.from_dsl(txt, data, "c32", replan=128, replan_n_amp=131072).compile()
    assert np.max(np.abs(g32.amplitudes(many) - ref)) / scale < 1e-5


@pytest.mark.parametrize("case", ["rqc_unsliced", "fsim"])
def test_library_sliced_program_on_gpu(gpu, case):
    """GPU-aware slicing (qxb_graph_replan_ex, n_free = -3): the library ADDS slice variables until no tensor exceeds the
    budget; the sliced program it writes must produce the original amplitudes on the GPU -- all slices, sub-ranges,
    ComplexF32 -- and equal the oracle run on the unsliced program (CPU counterpart: tests/test_replan.py)."""
    from cases import circuit_case
    if case == "rqc_unsliced":
        txt, data, bs = circuit_case(q.create_rqc_circuit(4, 4, 14, 7), n_slice=0, n_amp=6)
        n_q, lim = 16, 4
    else:
        txt, data, bs = circuit_case(q.create_sycamore_like_circuit(8, seed=3, n_qubits=18), n_slice=0, n_amp=6)
        n_q, lim = 18, 5
    ref = orc.amplitudes(orc.parse_dsl(txt), data, bs)
    g = Graph.from_dsl(txt, data, "c64")
    info = g.replan(8, 1, n_free=-3, budget_bytes=3 * 16 * 2 ** lim)
    assert info["replanned"]
    sliced = g.text
    g2 = Graph.from_dsl(sliced, data, "c64").compile()
    assert g2.n_slices > 1
    assert rel_err(g2.amplitudes(bs), ref, n_q) < 1e-10
    S = g2.n_slices
    parts = sum(g2.amplitudes(bs, b, min(b + 5, S)) for b in range(0, S, 5))
    assert rel_err(parts, ref, n_q) < 1e-10
    assert rel_err(Graph.from_dsl(sliced, data, "c32").compile().amplitudes(bs), ref, n_q) < 1e-5
    assert rel_err(Graph.from_dsl(sliced, data, "c64").compile(row_programs=False).amplitudes(bs), ref, n_q) < 1e-10
