"""GPU parity: the CUDA path through the C ABI vs the numpy oracle on the same
seeded inputs.  Tolerances are BASELINE.json's: 1e-10 relative (ComplexF64),
1e-5 (ComplexF32)."""
import numpy as np
import pytest

import qxb200 as q
from qxb200.executor import Graph
from oracle import qx_oracle as orc
from cases import kat0, rqc_case, circuit_case, rel_err

pytestmark = pytest.mark.gpu
TOL = {"c64": 1e-10, "c32": 1e-5}


@pytest.mark.parametrize("dtype", ["c64", "c32"])
def test_kat0_docs_example(gpu, dtype):
    txt, data = kat0()
    g = Graph.from_dsl(txt, data, dtype).compile()
    out = g.amplitudes(["00", "11", "01", "10"])
    want = np.array([1 / np.sqrt(2), 1 / np.sqrt(2), 0, 0])
    assert rel_err(out, want, 2) < TOL[dtype]
    # every single slice and every sub-range agrees with the oracle
    cmds = orc.parse_dsl(txt)
    for b in range(4):
        for e in range(b, 5):
            ref = orc.amplitudes(cmds, data, ["00", "11", "01", "10"], slice_begin=b, slice_end=e)
            assert rel_err(g.amplitudes(["00", "11", "01", "10"], b, e), ref, 2) < TOL[dtype]


@pytest.mark.parametrize("dtype", ["c64", "c32"])
def test_ghz3_readme(gpu, dtype):
    tnc = q.convert_to_tnc(q.create_ghz_circuit(3))
    plan = q.flow_cutter_contraction_plan(tnc, time=0)
    got = [q.single_amplitude(tnc, plan, b, dtype=dtype) for b in ("000", "111", "100")]
    assert rel_err(got, [1 / np.sqrt(2), 1 / np.sqrt(2), 0], 3) < TOL[dtype]


def test_ghz5_run_simulation(gpu):
    """test/test_simulation.jl:16-26 of the reference."""
    res = q.run_simulation(q.create_ghz_circuit(5))
    assert len(res) == 32
    assert abs(res["00000"] - 1 / np.sqrt(2)) < 1e-12 and abs(res["11111"] - 1 / np.sqrt(2)) < 1e-12
    assert abs(sum(abs(v) ** 2 for v in res.values()) - 1) < 1e-12


@pytest.mark.parametrize("dtype", ["c64", "c32"])
@pytest.mark.parametrize("shape", [(3, 3, 8, 2), (3, 4, 10, 3), (4, 4, 12, 4)])
def test_rqc_sliced_vs_oracle(gpu, dtype, shape):
    r, c, d, ns = shape
    txt, data, bs = rqc_case(r, c, d, ns)
    cmds = orc.parse_dsl(txt)
    ref = orc.amplitudes(cmds, data, bs)
    g = Graph.from_dsl(txt, data, dtype).compile()
    assert rel_err(g.amplitudes(bs), ref, r * c) < TOL[dtype]
    S = g.n_slices
    ref2 = orc.amplitudes(cmds, data, bs, slice_begin=1, slice_end=S - 1)
    assert rel_err(g.amplitudes(bs, 1, S - 1), ref2, r * c) < TOL[dtype]


def test_plus_minus_outputs(gpu):
    txt, data, _ = circuit_case(q.create_ghz_circuit(4))
    bs = ["+-01", "1++0", "----", "0000"]
    ref = orc.amplitudes(orc.parse_dsl(txt), data, bs)
    g = Graph.from_dsl(txt, data, "c64").compile()
    assert rel_err(g.amplitudes(bs), ref, 4) < 1e-10


def test_amp_batching_and_budget(gpu):
    txt, data, bs = rqc_case(3, 4, 10, 3, n_amp=37)
    ref = orc.amplitudes(orc.parse_dsl(txt), data, bs)
    g = Graph.from_dsl(txt, data, "c64").compile(amp_batch=5)
    out = g.amplitudes(bs)
    assert g.stats()["amp_batch"] == 5
    assert rel_err(out, ref, 12) < 1e-10


@pytest.mark.parametrize("dtype", ["c64", "c32"])
def test_golden_vectors(gpu, dtype):
    import json, os
    here = os.path.dirname(os.path.abspath(__file__))
    gold = json.load(open(os.path.join(here, "golden", "golden_rqc.json")))
    for case in gold["cases"]:
        txt = open(os.path.join(here, "golden", case["qx"])).read()
        data = dict(np.load(os.path.join(here, "golden", case["npz"])))
        g = Graph.from_dsl(txt, data, dtype).compile()
        ref = np.array(case["re"]) + 1j * np.array(case["im"])
        assert rel_err(g.amplitudes(case["bitstrings"]), ref, case["n_qubits"]) < TOL[dtype]
        # every single slice of bitstring 0 (slice enumeration must match the oracle's)
        per = np.array(case["slice_re_bs0"]) + 1j * np.array(case["slice_im_bs0"])
        got = np.array([g.amplitudes(case["bitstrings"][:1], s, s + 1)[0] for s in range(0, len(per), 3)])
        assert rel_err(got, per[::3], case["n_qubits"]) < TOL[dtype]


@pytest.mark.parametrize("cuda_graph", [True, False])
@pytest.mark.parametrize("sum_at_root", [False, True])
def test_execution_modes_agree(gpu, cuda_graph, sum_at_root):
    txt, data, bs = rqc_case(4, 4, 12, 4, n_amp=33)
    ref = orc.amplitudes(orc.parse_dsl(txt), data, bs)
    g = Graph.from_dsl(txt, data, "c64").compile(cuda_graph=cuda_graph, sum_at_root=sum_at_root)
    for _ in range(3):                      # replayed graph must stay correct
        assert rel_err(g.amplitudes(bs), ref, 16) < 1e-10
    assert rel_err(g.amplitudes(bs[:7]), ref[:7], 16) < 1e-10


def test_subspace_partition(gpu):
    txt, data, bs = rqc_case(4, 4, 12, 4, n_amp=9)
    ref = orc.amplitudes(orc.parse_dsl(txt), data, bs)
    g = Graph.from_dsl(txt, data, "c64").compile()
    for n_parts in (2, 4, 8):
        tot = np.zeros(len(bs), dtype=np.complex128)
        for part in range(n_parts):
            fv, fx = g.partition_assignment(n_parts, part)
            tot += g.amplitudes_subspace(bs, fv, fx)
        assert rel_err(tot, ref, 16) < 1e-10
    # a fully fixed sub-space is one slice
    dims = g.slice_dims
    s = 5
    one = g.amplitudes_subspace(bs, list(range(len(dims))), g.slice_values(s))
    assert rel_err(one, g.amplitudes(bs, s, s + 1), 16) < 1e-12


def test_non_power_of_two_extents(gpu):
    rng = np.random.default_rng(3)
    A = rng.normal(size=(3, 2)) + 1j * rng.normal(size=(3, 2))
    B = rng.normal(size=(3, 3)) + 1j * rng.normal(size=(3, 3))
    Cc = rng.normal(size=(3, 2)) + 1j * rng.normal(size=(3, 2))
    txt = ("# version: 0.4.0\nload a dA 3,2\nview a_s a v1 1 3\nload b dB 3,3\nview b_s b v1 1 3\n"
           "load c dC 3,2\noutput o1 1 2\noutput o2 2 2\n"
           "ncon ab 1,2,3 a_s 1,2 b_s 1,3\nncon abc 1,2,4 ab 1,2,3 c 3,4\nncon x 1,4 abc 1,2,4 o1 2\n"
           "ncon y 0 x 1,4 o2 4\nsave output y\n")
    data = {"dA": A, "dB": B, "dC": Cc}
    g = Graph.from_dsl(txt, data, "c64").compile()
    bs = ["00", "01", "10", "11"]
    cmds = orc.parse_dsl(txt)
    for (b, e) in [(0, 3), (0, 1), (1, 3), (2, 3)]:
        ref = orc.amplitudes(cmds, data, bs, slice_begin=b, slice_end=e)
        assert np.allclose(g.amplitudes(bs, b, e), ref, atol=1e-12)


def test_large_k_and_qft(gpu):
    """QFT-8 unsliced (controlled-phase hyper-edges) and a deep GHZ: exercises K chunks > 1."""
    txt, data, _ = circuit_case(q.create_qft_circuit(8))
    allb = list(q.amplitudes_all(8))
    g = Graph.from_dsl(txt, data, "c64").compile()
    out = g.amplitudes(allb)
    assert np.allclose(np.abs(out), 2.0 ** -4, atol=1e-12)
    assert rel_err(out[::17], orc.amplitudes(orc.parse_dsl(txt), data, allb[::17]), 8) < 1e-10


def test_execute_file_triple(gpu, tmp_path):
    from qxb200.execute import execute
    prefix = str(tmp_path / "rqc")
    circ = q.create_rqc_circuit(3, 3, 8, 42)
    q.generate_simulation_files(circ, prefix, 3, seed=42, time=0, output_args=q.output_params_dict(9, 6, seed=5))
    from qxb200.jld2 import load_jld2
    res = execute(prefix + ".qx", output_file=prefix + "_out.npz", dtype="c64")
    txt = open(prefix + ".qx").read()
    data = dict(load_jld2(prefix + ".jld2"))
    ref = orc.amplitudes(orc.parse_dsl(txt), data, list(res.keys()))
    assert rel_err(np.array(list(res.values())), ref, 9) < 1e-10
    res2 = execute(prefix + ".qx", max_amplitudes=2, max_slices=3, dtype="c64")
    ref2 = orc.amplitudes(orc.parse_dsl(txt), data, list(res.keys())[:2], slice_begin=0, slice_end=3)
    assert len(res2) == 2 and rel_err(np.array(list(res2.values())), ref2, 9) < 1e-10
    out = np.load(prefix + "_out.npz")
    assert list(out["bitstrings"]) == list(res.keys())


@pytest.mark.parametrize("dtype", ["c64", "c32"])
def test_reduction_shaped_nodes(gpu, dtype):
    """Few outputs, long K (K = 64 and K = 128 with a bitstring axis): the warp-reduce kernel."""
    rng = np.random.default_rng(5)
    A = rng.normal(size=(2,) * 6) + 1j * rng.normal(size=(2,) * 6)
    B = rng.normal(size=(2,) * 7) + 1j * rng.normal(size=(2,) * 7)
    Cc = rng.normal(size=(2,) * 8) + 1j * rng.normal(size=(2,) * 8)
    txt = ("# version: 0.4.0\nload a dA 2,2,2,2,2,2\nload b dB 2,2,2,2,2,2,2\nload c dC 2,2,2,2,2,2,2,2\n"
           "output o1 1 2\noutput o2 2 2\n"
           "ncon t 7 a 1,2,3,4,5,6 b 1,2,3,4,5,6,7\n"            # const phase, K = 64
           "ncon w 1,2,3,4,5,6,7 c 1,2,3,4,5,6,7,8 o1 8\n"       # chunk phase
           "ncon x 7 w 1,2,3,4,5,6,7 a 1,2,3,4,5,6\n"            # chunk phase, K = 64, 2 outputs per bitstring
           "ncon y 7 x 7 t 7\nncon z 0 y 7 o2 7\nsave output z\n")
    data = {"dA": A, "dB": B, "dC": Cc}
    bs = ["00", "01", "10", "11"]
    ref = orc.amplitudes(orc.parse_dsl(txt), data, bs)
    g = Graph.from_dsl(txt, data, dtype).compile()
    got = g.amplitudes(bs)
    assert np.max(np.abs(got - ref)) / np.max(np.abs(ref)) < (1e-12 if dtype == "c64" else 2e-5)


@pytest.mark.parametrize("dtype", ["c64", "c32"])
def test_replanned_graph_on_gpu(gpu, dtype):
    txt, data, bs = rqc_case(4, 5, 14, 5, n_amp=17)
    cmds = orc.parse_dsl(txt)
    ref = orc.amplitudes(cmds, data, bs)
    g = Graph.from_dsl(txt, data, dtype, replan=6, replan_n_amp=64).compile()
    assert g.replan_info["replanned"]
    assert rel_err(g.amplitudes(bs), ref, 20) < TOL[dtype]
    assert rel_err(g.amplitudes(bs, 3, 29), orc.amplitudes(cmds, data, bs, slice_begin=3, slice_end=29), 20) < TOL[dtype]
    tot = sum(g.amplitudes_subspace(bs, *g.partition_assignment(4, r)) for r in range(4))
    assert rel_err(tot, ref, 20) < TOL[dtype]


def test_edge_cases_and_errors(gpu):
    from qxb200._lib import QxbError
    txt, data, bs = rqc_case(3, 4, 10, 3, n_amp=6)
    cmds = orc.parse_dsl(txt)
    g = Graph.from_dsl(txt, data, "c64").compile()
    # empty slice range -> zeros; empty batch -> empty result; a single bitstring
    assert np.all(g.amplitudes(bs, 2, 2) == 0)
    assert g.amplitudes([]).shape == (0,)
    assert rel_err(g.amplitudes(bs[:1]), orc.amplitudes(cmds, data, bs[:1]), 12) < 1e-10
    # out-of-range slices / malformed bitstrings are argument errors, not crashes
    for bad in ((-1, 2), (0, 9), (5, 3)):
        with pytest.raises(QxbError) as e:
            g.amplitudes(bs, *bad)
        assert e.value.code == -1
    with pytest.raises(QxbError):
        g.amplitudes(np.full((2, 12), 7, dtype=np.uint8))
    with pytest.raises(ValueError):
        g.amplitudes(["01"])                     # shorter than the number of outputs
    with pytest.raises(QxbError) as e:
        g.compile()                              # already compiled
    assert e.value.code == -2
    # per-op kernels (the intermediates of a bitstring row live in HBM): a budget that cannot hold one row is a memory
    # error, and a tight one forces batching
    with pytest.raises(QxbError) as e:
        Graph.from_dsl(txt, data, "c64").compile(hbm_budget_bytes=64, row_programs=False).amplitudes(bs)
    assert e.value.code == -5
    gt = Graph.from_dsl(txt, data, "c64").compile(hbm_budget_bytes=40_000, row_programs=False)
    out = gt.amplitudes(bs)
    assert gt.stats()["amp_batch"] < len(bs)
    assert rel_err(out, orc.amplitudes(cmds, data, bs), 12) < 1e-10
    # row programs keep a row's intermediates in shared memory: the same tiny budget is enough
    gr = Graph.from_dsl(txt, data, "c64").compile(hbm_budget_bytes=40_000, row_programs="all")
    assert rel_err(gr.amplitudes(bs), orc.amplitudes(cmds, data, bs), 12) < 1e-10
    # missing leaf data
    with pytest.raises(QxbError) as e:
        Graph.from_dsl(txt, {k: v for k, v in list(data.items())[1:]}, "c64").compile()
    assert e.value.code == -2


def test_execute_rejection_sampler(gpu, tmp_path):
    """qxrun on a Rejection parameter file (outputs.jl:57-62): GHZ-5 yields only its two outcomes."""
    from qxb200.execute import execute
    prefix = str(tmp_path / "ghz5")
    q.generate_simulation_files(q.create_ghz_circuit(5), prefix, 1, time=0,
                                output_args=q.output_params_dict(5, 12, output_method="Rejection", M=16.0,
                                                                 fix_M=True, seed=4))
    res = execute(prefix + ".qx", output_file=prefix + "_out.npz", dtype="c64")
    assert 1 <= len(res) <= 2 and set(res) <= {"00000", "11111"}
    assert all(abs(abs(a) - 1 / np.sqrt(2)) < 1e-12 for a in res.values())
    out = np.load(prefix + "_out.npz")
    assert len(out["bitstrings"]) == 12 and abs(float(out["M"]) - 16.0) < 1e-12


def test_variables_fixed_automatically_to_fit_hbm(gpu):
    """When batching every slice variable needs more workspace than the budget allows, the
    executor fixes variables (loops over their values) instead of failing."""
    txt, data, bs = rqc_case(4, 4, 12, 4, n_amp=5)
    ref = orc.amplitudes(orc.parse_dsl(txt), data, bs)
    g0 = Graph.from_dsl(txt, data, "c64")
    d = g0.describe()
    need = 16 * (d["arena_elems"]["chunk_per_amp"] + d["arena_elems"]["block"])
    for frac in (0.7, 0.4):
        g = Graph.from_dsl(txt, data, "c64").compile(hbm_budget_bytes=int(frac * need))
        out = g.amplitudes(bs)
        st = g.stats()
        assert st["n_blocks"] > 1 and st["amp_batch"] >= 1
        assert rel_err(out, ref, 16) < 1e-10
        assert rel_err(g.amplitudes(bs, 3, 11), orc.amplitudes(orc.parse_dsl(txt), data, bs, slice_begin=3, slice_end=11), 16) < 1e-10


def test_undecomposed_gates(gpu):
    """decompose=false keeps 2-qubit gates as rank-4 tensors (test/test_tn_conversion.jl:10-12)."""
    circ = q.create_rqc_circuit(3, 3, 8, 9)
    txt, data, bs = circuit_case(circ, n_slice=2, decompose=False, n_amp=6)
    ref = orc.amplitudes(orc.parse_dsl(txt), data, bs)
    for replan in (0, 8):
        g = Graph.from_dsl(txt, data, "c64", replan=replan, replan_n_amp=8).compile()
        assert rel_err(g.amplitudes(bs), ref, 9) < 1e-10


def test_network_without_outputs(gpu):
    """A closed network with no `output` statement: one value, whatever bitstrings are passed."""
    rng = np.random.default_rng(0)
    A = rng.normal(size=(2, 2, 2)) + 1j * rng.normal(size=(2, 2, 2))
    B = rng.normal(size=(2, 2)) + 1j * rng.normal(size=(2, 2))
    txt = ("# version: 0.4.0\nload a dA 2,2,2\nview a_s a v1 3 2\nload b dB 2,2\nload c dB 2,2\nview c_s c v1 1 2\n"
           "ncon ab 2,3 a_s 1,2,3 b 1,2\nncon r 0 ab 2,3 c_s 3,2\nsave output r\n")
    data = {"dA": A, "dB": B}
    ref = orc.amplitudes(orc.parse_dsl(txt), data, ["", "", ""])
    g = Graph.from_dsl(txt, data, "c64").compile()
    assert g.n_outputs == 0 and g.n_slices == 2
    out = g.amplitudes(np.zeros((3, 0), dtype=np.uint8))
    assert np.allclose(out, ref, atol=1e-12)
    assert np.allclose(g.amplitudes(np.zeros((3, 0), dtype=np.uint8), 1, 2),
                       orc.amplitudes(orc.parse_dsl(txt), data, ["", "", ""], slice_begin=1, slice_end=2), atol=1e-12)


@pytest.mark.parametrize("dtype", ["c64", "c32"])
def test_fsim_circuit_extent4_bonds(gpu, dtype):
    circ = q.create_sycamore_like_circuit(4, seed=3, n_qubits=12)
    txt, data, bs = circuit_case(circ, n_slice=3, n_amp=9)
    cmds = orc.parse_dsl(txt)
    ref = orc.amplitudes(cmds, data, bs)
    for replan in (0, 8):
        g = Graph.from_dsl(txt, data, dtype, replan=replan, replan_n_amp=16).compile()
        assert rel_err(g.amplitudes(bs), ref, 12) < TOL[dtype]
        S = g.n_slices
        assert rel_err(g.amplitudes(bs, 5, S - 3), orc.amplitudes(cmds, data, bs, slice_begin=5, slice_end=S - 3), 12) < TOL[dtype]


@pytest.mark.parametrize("dtype", ["c64", "c32"])
def test_long_k_block_reduction(gpu, dtype):
    """K = 2^13 with one output per bitstring: the block-per-output reduction kernel."""
    rng = np.random.default_rng(11)
    A = (rng.normal(size=(2,) * 13) + 1j * rng.normal(size=(2,) * 13)) / 8
    B = (rng.normal(size=(2,) * 14) + 1j * rng.normal(size=(2,) * 14)) / 8
    l13 = ",".join(str(i) for i in range(1, 14))
    l14 = ",".join(str(i) for i in range(1, 15))
    txt = ("# version: 0.4.0\nload a dA " + ",".join(["2"] * 13) + "\nload b dB " + ",".join(["2"] * 14) + "\noutput o1 1 2\n"
           f"ncon w {l13} b {l14} o1 14\n"            # chunk phase: 2^13 elements per bitstring
           f"ncon z 0 w {l13} a {l13}\nsave output z\n")   # K = 2^13, nC = 0
    data = {"dA": A, "dB": B}
    bs = ["0", "1", "+", "-"]
    ref = orc.amplitudes(orc.parse_dsl(txt), data, bs)
    g = Graph.from_dsl(txt, data, dtype).compile()
    got = g.amplitudes(bs)
    assert np.max(np.abs(got - ref)) / np.max(np.abs(ref)) < (1e-12 if dtype == "c64" else 5e-5)


@pytest.mark.parametrize("dtype", ["c64", "c32"])
@pytest.mark.parametrize("gemm_mode", [1, 2, 4])
@pytest.mark.parametrize("seed", [0, 1])
def test_gemm_shaped_node_random(gpu, dtype, gemm_mode, seed):
    """One GEMM-shaped node on random data with shuffled mode orders: M = 2^8, N = 2^8, K = 2^6 per
    bitstring, through the SIMT tile kernel (gemm_mode 1), the tensor-core kernels of the auto choice (gemm_mode 2:
    DMMA for ComplexF64, tcgen05 / TMEM 3xTF32 for ComplexF32) and the mma.sync kernels (gemm_mode 4)."""
    rng = np.random.default_rng(100 + seed)
    nm, nn, nk = 8, 8, 6
    M = list(range(1, nm + 1)); N = list(range(nm + 1, nm + nn + 1)); K = list(range(nm + nn + 1, nm + nn + nk + 1))
    o_lab = nm + nn + nk + 1
    la = list(rng.permutation(M + K)); lb = list(rng.permutation(N + K)) + [o_lab]
    lb2 = [x for x in lb if x != o_lab]
    lc = list(rng.permutation(M + N))
    A = (rng.normal(size=(2,) * len(la)) + 1j * rng.normal(size=(2,) * len(la))) / 4
    B = (rng.normal(size=(2,) * len(lb)) + 1j * rng.normal(size=(2,) * len(lb))) / 4
    V = (rng.normal(size=(2,) * len(lc)) + 1j * rng.normal(size=(2,) * len(lc))) / 16
    j = lambda l: ",".join(str(int(i)) for i in l)
    txt = ("# version: 0.4.0\n"
           f"load a dA {j([2] * len(la))}\nload b dB {j([2] * len(lb))}\nload v dV {j([2] * len(lc))}\noutput o1 1 2\n"
           f"ncon b2 {j(lb2)} b {j(lb)} o1 {o_lab}\n"
           f"ncon c {j(lc)} a {j(la)} b2 {j(lb2)}\n"
           f"ncon z 0 c {j(lc)} v {j(lc)}\nsave output z\n")
    data = {"dA": A, "dB": B, "dV": V}
    bs = ["0", "1", "+", "-"]
    ref = orc.amplitudes(orc.parse_dsl(txt), data, bs)
    g = Graph.from_dsl(txt, data, dtype).compile(gemm_mode=gemm_mode)
    got = g.amplitudes(bs)
    d = [o for o in g.describe()["ops"] if o["name"] == "c"][0]
    assert d["m_bits"] == nm and d["n_bits"] == nn and d["nK"] == nk
    assert np.max(np.abs(got - ref)) / np.max(np.abs(ref)) < (1e-12 if dtype == "c64" else 2e-5)
    # which kernel ran the node
    import os, tempfile
    gp = Graph.from_dsl(txt, data, dtype).compile(gemm_mode=gemm_mode, profile=True, row_programs=False)
    gp.amplitudes(bs)
    prof = gp.profile_dump(os.path.join(tempfile.mkdtemp(), "p.json"))
    kern = [o["kernel"] for v in prof["variants"] for o in v["ops"] if o["name"] == "c"][0]
    want = {1: "gemm_simt", 2: "gemm_tc5" if dtype == "c32" else "gemm_tc", 4: "gemm_tc"}[gemm_mode]
    assert kern == want, (kern, want)


@pytest.mark.parametrize("dtype", ["c64", "c32"])
@pytest.mark.parametrize("aligned", [False, True])
@pytest.mark.parametrize("nk,nn,big_first", [(3, 1, True), (4, 4, True), (2, 3, False), (5, 2, False), (4, 5, True), (3, 7, True), (4, 6, False)])
def test_big_times_small_streaming_node(gpu, dtype, nk, nn, big_first, aligned):
    """A 2^(17 + nk)-element operand against a 2^(nk + nn)-element one (per bitstring), shuffled mode orders, either
    operand order: the streaming kernel of csrc/qxb_kred.cu (a thread owns one position of the big operand and all
    min(2^nn, 2^4 or 2^5) outputs, further N bits go to the CTA index) -- the shape of the 24 dominant nodes of a Sycamore-53 depth-12 slice."""
    rng = np.random.default_rng(7 * nk + nn)
    nm = max(17 if nn <= 5 else 14, 20 - nk)                                 # the kernel takes operands of >= 2^20 elements
    ks = list(range(1, nk + 1)); ms = list(range(nk + 1, nk + nm + 1)); ns = list(range(nk + nm + 1, nk + nm + nn + 1))
    o_lab = nk + nm + nn + 1
    la = list(rng.permutation(ks + ms)); lb = list(rng.permutation(ks + ns)) + [o_lab]
    lb2 = [x for x in lb if x != o_lab]
    lc = list(rng.permutation(ms + ns))
    if aligned:
        # the 9 fastest modes of the big operand are also the 9 fastest modes of the result: a CTA's 256 positions are one
        # contiguous run at every k, the kernel brings them in by TMA bulk copies (bigsmall_tma)
        head = ms[:9]
        la = head + [x for x in la if x not in head]
        lc = head + [x for x in lc if x not in head]
    A = (rng.normal(size=(2,) * len(la)) + 1j * rng.normal(size=(2,) * len(la))) / 4
    B = (rng.normal(size=(2,) * len(lb)) + 1j * rng.normal(size=(2,) * len(lb))) / 4
    W = (rng.normal(size=(2,) * len(lc)) + 1j * rng.normal(size=(2,) * len(lc))) / 64
    j = lambda l: ",".join(str(int(i)) for i in l)
    pair = f"a {j(la)} b2 {j(lb2)}" if big_first else f"b2 {j(lb2)} a {j(la)}"
    txt = ("# version: 0.4.0\n"
           f"load a dA {j([2] * len(la))}\nload b dB {j([2] * len(lb))}\nload w dW {j([2] * len(lc))}\noutput o1 1 2\n"
           f"ncon b2 {j(lb2)} b {j(lb)} o1 {o_lab}\n"
           f"ncon c {j(lc)} {pair}\n"
           f"ncon z 0 c {j(lc)} w {j(lc)}\nsave output z\n")
    data = {"dA": A, "dB": B, "dW": W}
    bs = ["-"] if aligned else ["0", "1", "+"]               # the TMA variant takes single-row launches
    ref = orc.amplitudes(orc.parse_dsl(txt), data, bs)
    import os, tempfile
    # aligned layouts also go through the opt-in TMA-staged variant (measured no faster than the register one)
    g = Graph.from_dsl(txt, data, dtype).compile(profile=True, row_programs=False, streaming="tma" if aligned else True)
    got = g.amplitudes(bs)
    assert np.max(np.abs(got - ref)) / np.max(np.abs(ref)) < (1e-12 if dtype == "c64" else 2e-5)
    prof = g.profile_dump(os.path.join(tempfile.mkdtemp(), "p.json"))
    kern = [o["kernel"] for v in prof["variants"] for o in v["ops"] if o["name"] == "c"][0]
    assert kern == ("bigsmall_tma" if aligned else "bigsmall"), kern
    variants = [dict(streaming=False)]                        # the node through the general kernels
    if aligned:
        variants.append(dict())                               # the default (register) kernel on the aligned layout
    elif dtype == "c32":
        variants.append(dict(streaming="ffma2"))              # packed-FFMA2 accumulators (opt-in)
    for kw in variants:
        alt = Graph.from_dsl(txt, data, dtype).compile(row_programs=False, **kw).amplitudes(bs)
        assert np.max(np.abs(alt - ref)) / np.max(np.abs(ref)) < (1e-12 if dtype == "c64" else 2e-5), kw


@pytest.mark.parametrize("dtype", ["c64", "c32"])
@pytest.mark.parametrize("nk,nm,nn", [(14, 4, 4), (13, 3, 4), (12, 4, 2)])
def test_tiled_split_k_reduction(gpu, dtype, nk, nm, nn):
    """2^nm x 2^nn outputs (up to 16 x 16 = 2^8: the template of such a node carries register-tile bits) against
    K = 2^nk, shuffled mode orders: the tiled split-K kernel (csrc/qxb_kred.cu: operand tiles of a K chunk staged in
    shared memory once per CTA) -- the shape of the root-like node that was 39 % of a Sycamore-53 depth-12 slice."""
    rng = np.random.default_rng(31 + nk)
    ks = list(range(1, nk + 1)); ms = list(range(nk + 1, nk + nm + 1)); ns = list(range(nk + nm + 1, nk + nm + nn + 1))
    o_lab = nk + nm + nn + 1
    la = list(rng.permutation(ks + ms)); lb = list(rng.permutation(ks + ns)) + [o_lab]
    lb2 = [x for x in lb if x != o_lab]
    lc = list(rng.permutation(ms + ns))
    A = (rng.normal(size=(2,) * len(la)) + 1j * rng.normal(size=(2,) * len(la))) / 16
    B = (rng.normal(size=(2,) * len(lb)) + 1j * rng.normal(size=(2,) * len(lb))) / 16
    W = rng.normal(size=(2,) * len(lc)) + 1j * rng.normal(size=(2,) * len(lc))
    j = lambda l: ",".join(str(int(i)) for i in l)
    txt = ("# version: 0.4.0\n"
           f"load a dA {j([2] * len(la))}\nload b dB {j([2] * len(lb))}\nload w dW {j([2] * len(lc))}\noutput o1 1 2\n"
           f"ncon b2 {j(lb2)} b {j(lb)} o1 {o_lab}\n"
           f"ncon c {j(lc)} a {j(la)} b2 {j(lb2)}\n"
           f"ncon z 0 c {j(lc)} w {j(lc)}\nsave output z\n")
    data = {"dA": A, "dB": B, "dW": W}
    bs = ["0", "1", "+", "-"]
    ref = orc.amplitudes(orc.parse_dsl(txt), data, bs)
    import os, tempfile
    g = Graph.from_dsl(txt, data, dtype).compile(profile=True, row_programs=False)     # per-op kernels (4 bitstrings would run as a row program)
    got = g.amplitudes(bs)
    prof = g.profile_dump(os.path.join(tempfile.mkdtemp(), "p.json"))
    o = [o for v in prof["variants"] for o in v["ops"] if o["name"] == "c"][0]
    assert o["nK"] == nk and o["nC"] == nm + nn and o["kernel"] == "kreduce_grid", o     # 2 x 2 outputs per thread, cp.async ring
    os.environ["QXB_KRED_GRID"] = "0"                                                      # the one-output-per-thread variant
    try:
        g1 = Graph.from_dsl(txt, data, dtype).compile(profile=True, row_programs=False)
        got1 = g1.amplitudes(bs)
        o1 = [o for v in g1.profile_dump(os.path.join(tempfile.mkdtemp(), "p.json"))["variants"] for o in v["ops"] if o["name"] == "c"][0]
    finally:
        del os.environ["QXB_KRED_GRID"]
    assert o1["kernel"] == "kreduce_tile", o1
    assert np.max(np.abs(got1 - ref)) / np.max(np.abs(ref)) < (1e-12 if dtype == "c64" else 5e-5)
    auto = Graph.from_dsl(txt, data, dtype).compile().amplitudes(bs)                    # whatever the auto mode picks
    assert np.max(np.abs(auto - ref)) / np.max(np.abs(ref)) < (1e-12 if dtype == "c64" else 5e-5)
    assert np.max(np.abs(got - ref)) / np.max(np.abs(ref)) < (1e-12 if dtype == "c64" else 5e-5)


@pytest.mark.parametrize("dtype", ["c64", "c32"])
def test_split_k_reduction(gpu, dtype):
    """K = 2^17 against 2 x 4 outputs: the split-K reduction kernel (several blocks per output, atomicAdd
    into a zeroed C) that the root-like nodes of GEMM-shaped trees go through."""
    rng = np.random.default_rng(12)
    nk = 17
    A = (rng.normal(size=(2,) * (nk + 1)) + 1j * rng.normal(size=(2,) * (nk + 1))) / 16
    B = (rng.normal(size=(2,) * (nk + 2)) + 1j * rng.normal(size=(2,) * (nk + 2))) / 16
    ks = list(range(1, nk + 1))
    la = ks + [nk + 1]                       # one M-only mode
    lb = [nk + 2] + ks + [nk + 3]            # one N-only mode + the output wire
    lb2 = [nk + 2] + ks
    j = lambda l: ",".join(str(i) for i in l)
    txt = ("# version: 0.4.0\n"
           f"load a dA {j([2] * len(la))}\nload b dB {j([2] * len(lb))}\nload w dW 2,2\noutput o1 1 2\n"
           f"ncon b2 {j(lb2)} b {j(lb)} o1 {nk + 3}\n"
           f"ncon c {nk + 1},{nk + 2} a {j(la)} b2 {j(lb2)}\n"       # K = 2^17, nC = 2
           f"ncon z 0 c 1,2 w 1,2\nsave output z\n")
    W = rng.normal(size=(2, 2)) + 1j * rng.normal(size=(2, 2))
    data = {"dA": A, "dB": B, "dW": W}
    bs = ["0", "1", "+", "-"]
    ref = orc.amplitudes(orc.parse_dsl(txt), data, bs)
    g = Graph.from_dsl(txt, data, dtype).compile()
    got = g.amplitudes(bs)
    d = [o for o in g.describe()["ops"] if o["name"] == "c"][0]
    assert d["nK"] == nk and d["nC"] == 2
    assert np.max(np.abs(got - ref)) / np.max(np.abs(ref)) < (1e-12 if dtype == "c64" else 5e-5)
    # replay: the partial sums meet by atomicAdd, so the order of the additions (not their set) may differ between runs
    assert np.max(np.abs(g.amplitudes(bs) - got)) / np.max(np.abs(ref)) < (1e-13 if dtype == "c64" else 5e-5)
