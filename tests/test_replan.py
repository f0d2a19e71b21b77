"""The batch-aware re-planner is a DSL -> DSL rewrite that must not change the value of
the program: same leaves and views, same amplitudes for every slice range."""
import numpy as np
import pytest

import qxb200 as q
from qxb200.executor import Graph, bits_from_strings
from replan_prototype import replan_dsl, recover_network
from oracle import qx_oracle as orc
import lowered_emulator as em
from cases import kat0, rqc_case, circuit_case


def _leaf_lines(txt):
    return [ln for ln in txt.splitlines() if ln.split() and ln.split()[0] in ("load", "output", "view")]


@pytest.mark.parametrize("shape", [(3, 3, 8, 2), (4, 4, 12, 4), (4, 5, 14, 5)])
def test_replanned_program_is_equivalent(lib_built, shape):
    r, c, d, ns = shape
    txt, data, bs = rqc_case(r, c, d, ns, n_amp=4)
    new, info = replan_dsl(txt, n_amp=64, candidates=6)
    assert info["replanned"] and info["bytes"] < info["given_bytes"]
    assert _leaf_lines(new) == _leaf_lines(txt)                 # leaves and views untouched
    c0, c1 = orc.parse_dsl(txt), orc.parse_dsl(new)
    assert orc.slice_dims(c0) == orc.slice_dims(c1)
    assert np.allclose(orc.amplitudes(c1, data, bs), orc.amplitudes(c0, data, bs), atol=1e-15)
    S = 2 ** ns
    for (b, e) in [(0, 1), (1, S - 1), (S - 1, S)]:               # per-slice values are preserved too
        assert np.allclose(orc.amplitudes(c1, data, bs, slice_begin=b, slice_end=e),
                           orc.amplitudes(c0, data, bs, slice_begin=b, slice_end=e), atol=1e-15)
    # and the lowered form of the re-planned program executes correctly (random topological order)
    g = Graph.from_dsl(new, data)
    assert np.allclose(em.amplitudes(g, data, bits_from_strings(bs, r * c), shuffle_seed=3),
                       orc.amplitudes(c0, data, bs), atol=1e-14)


def test_replan_docs_example_and_unsliced(lib_built):
    txt, data = kat0()
    new, info = replan_dsl(txt, n_amp=4, candidates=6)
    got = orc.amplitudes(orc.parse_dsl(new), data, ["00", "11", "01", "10"])
    assert np.allclose(got, [1 / np.sqrt(2), 1 / np.sqrt(2), 0, 0], atol=1e-15)
    txt, data, _ = circuit_case(q.create_qft_circuit(6))
    new, info = replan_dsl(txt, n_amp=16, candidates=6)
    allb = list(q.amplitudes_all(6))
    assert np.allclose(orc.amplitudes(orc.parse_dsl(new), data, allb), orc.amplitudes(orc.parse_dsl(txt), data, allb),
                       atol=1e-15)


def test_recover_network_classes(lib_built):
    txt, data = kat0()
    header, leaf_lines, leaves, dims, sliced, scalar = recover_network(txt)
    assert scalar and len(leaves) == 7 and set(sliced.values()) == {"v1", "v2"}
    # every index class is shared by at least two leaves (closed network)
    cnt = {}
    for lf in leaves:
        for cl in set(lf.classes):
            cnt[cl] = cnt.get(cl, 0) + 1
    assert min(cnt.values()) >= 2


def test_from_dsl_replan_flag(lib_built):
    txt, data, bs = rqc_case(4, 4, 12, 4, n_amp=3)
    g = Graph.from_dsl(txt, data, "c64", replan=6, replan_n_amp=64)
    assert g.replan_info["replanned"]
    assert g.cost_bytes(64) < Graph.from_dsl(txt, data, "c64").cost_bytes(64)
    # partition chooser is consistent: every rank gets the same kind of share
    kinds = {g.choose_partition(64, 4, r)[0] for r in range(4)}
    assert len(kinds) == 1


@pytest.mark.parametrize("shape", [(3, 3, 8, 2), (4, 4, 12, 4), (4, 5, 14, 5)])
def test_in_library_replan_is_equivalent(lib_built, shape):
    """qxb_graph_replan (C++, what Graph.from_dsl(replan=N) uses): same contract as replan_dsl."""
    r, c, d, ns = shape
    txt, data, bs = rqc_case(r, c, d, ns, n_amp=4)
    g = Graph.from_dsl(txt, data, "c64")
    info = g.replan(16, 64)
    new = g.program_text()
    assert info["replanned"] and info["bytes"] < info["given_bytes"]
    assert _leaf_lines(new) == _leaf_lines(txt)
    c0, c1 = orc.parse_dsl(txt), orc.parse_dsl(new)
    assert np.allclose(orc.amplitudes(c1, data, bs), orc.amplitudes(c0, data, bs), atol=1e-15)
    S = 2 ** ns
    assert np.allclose(orc.amplitudes(c1, data, bs, slice_begin=1, slice_end=S - 1),
                       orc.amplitudes(c0, data, bs, slice_begin=1, slice_end=S - 1), atol=1e-15)
    assert np.allclose(em.amplitudes(g, data, bits_from_strings(bs, r * c), shuffle_seed=5),
                       orc.amplitudes(c0, data, bs), atol=1e-14)
    # seeded: the same call gives the same program
    g2 = Graph.from_dsl(txt, data, "c64")
    g2.replan(16, 64)
    assert g2.program_text() == new


def test_in_library_replan_keeps_programs_it_cannot_improve(lib_built):
    txt, data = kat0()
    g = Graph.from_dsl(txt, data)
    info = g.replan(8, 4)
    got = orc.amplitudes(orc.parse_dsl(g.program_text()), data, ["00", "11", "01", "10"])
    assert np.allclose(got, [1 / np.sqrt(2), 1 / np.sqrt(2), 0, 0], atol=1e-15)
    assert info["bytes"] <= info["given_bytes"]


@pytest.mark.parametrize("case", ["rqc_unsliced", "rqc_sliced", "fsim"])
def test_autoslice_adds_exact_slice_variables(lib_built, case):
    """GPU-aware slicing (qxb_graph_replan_ex, n_free = -3): with a budget far below the largest tensor of the
    searched tree the library must ADD slice variables (views on every leaf of the chosen index classes) and the
    sliced program must still produce the original amplitudes -- summed over all slices, and range by range."""
    if case == "rqc_unsliced":
        txt, data, bs = circuit_case(q.create_rqc_circuit(4, 4, 14, 7), n_slice=0, n_amp=4)
        n_q = 16
    elif case == "rqc_sliced":
        txt, data, bs = rqc_case(4, 4, 14, 2, n_amp=4)
        n_q = 16
    else:
        txt, data, bs = circuit_case(q.create_sycamore_like_circuit(8, seed=3, n_qubits=18), n_slice=0, n_amp=4)
        n_q = 18
    lim = 5 if case == "fsim" else 4                                      # no tensor above 2^lim elements
    c0 = orc.parse_dsl(txt)
    g = Graph.from_dsl(txt, data, "c64")
    k0 = len(g.slice_dims)
    info = g.replan(8, 1, n_free=-3, budget_bytes=3 * 16 * 2 ** lim)
    g2 = Graph.from_dsl(g.text, data, "c64")
    assert info["replanned"] and len(g2.slice_dims) > k0 and info["n_free"] == k0
    assert g2.slice_dims[:k0] == g.slice_dims[:k0]
    c1 = orc.parse_dsl(g.text)
    ref = orc.amplitudes(c0, data, bs)
    assert np.allclose(orc.amplitudes(c1, data, bs), ref, atol=1e-14)
    S = g2.n_slices
    parts = sum(orc.amplitudes(c1, data, bs, slice_begin=b, slice_end=min(b + 5, S)) for b in range(0, S, 5))
    assert np.allclose(parts, ref, atol=1e-14)
    # every tensor of the sliced program, one slice at a time, is within the budget
    d = g2.describe(k0)
    assert max(o["nC"] for o in d["ops"] if o["phase"] != "const") <= lim
    # and the lowered form executes correctly
    assert np.allclose(em.amplitudes(g2, data, bits_from_strings(bs, n_q), shuffle_seed=5), ref, atol=1e-13)


def test_replan_for_partially_batched_runs(lib_built):
    """qxb_graph_replan_ex with n_free >= 0 / -2: the plan is searched and scored for a run that batches only the
    first n_free slice variables; the program stays equivalent for every slice range."""
    txt, data, bs = rqc_case(4, 4, 14, 4, n_amp=4)
    c0 = orc.parse_dsl(txt)
    ref = orc.amplitudes(c0, data, bs)
    for n_free in (0, 2, -2):
        g = Graph.from_dsl(txt, data, "c64")
        info = g.replan(8, 16, n_free=n_free, budget_bytes=3 * 16 * 2 ** 9 if n_free == -2 else 0)
        assert 0 <= info["n_free"] <= 4 and (n_free < 0 or info["n_free"] == n_free)
        assert _leaf_lines(g.text) == _leaf_lines(txt)
        c1 = orc.parse_dsl(g.text)
        assert np.allclose(orc.amplitudes(c1, data, bs), ref, atol=1e-14)
        assert np.allclose(orc.amplitudes(c1, data, bs, slice_begin=3, slice_end=11),
                           orc.amplitudes(c0, data, bs, slice_begin=3, slice_end=11), atol=1e-14)
        if info["replanned"]:
            assert np.isfinite(info["model_seconds_per_block"]) and info["model_seconds_per_block"] > 0


def test_replan_is_deterministic_and_tree_search_beats_orders(lib_built):
    """Same seed, same program; and on a sliced RQC the tree search must not be worse than the order-based
    re-planner of the first version (qxb200.replan.replan_dsl scores elimination orders only)."""
    txt, data, bs = rqc_case(4, 5, 14, 5, n_amp=4)
    a = Graph.from_dsl(txt, data, "c64"); ia = a.replan(8, 1024)
    b = Graph.from_dsl(txt, data, "c64"); ib = b.replan(8, 1024)
    assert a.text == b.text and ia["bytes"] == ib["bytes"]
    _, info = replan_dsl(txt, n_amp=1024, candidates=8)
    assert ia["bytes"] <= info["bytes"] * 1.0001


def test_l1_aware_model_avoids_long_k_untiled_nodes(lib_built, monkeypatch):
    """The planner's time model charges the SIMT kernel's operand loads (qxb_treeopt.h: l1_bandwidth).  On the headline
    workload the plan it finds must have clearly fewer MACs than the single-rate model of r1p (QXB_PLAN_L1_BW=0) at about
    the same bytes, and no node with K >= 32 and no register tile (the 552 GB/s node of profiles/r1p_ops.md)."""
    import bench
    from template_emulator import templates
    txt, data, w = bench.build_workload("rqc_7x7_d20_c64_s4096")
    n_amp = 131072

    def plan(l1):
        if l1 is None:
            monkeypatch.delenv("QXB_PLAN_L1_BW", raising=False)
        else:
            monkeypatch.setenv("QXB_PLAN_L1_BW", l1)
        g = Graph.from_dsl(txt, data, "c64", replan=48, replan_n_amp=n_amp)
        ops = [(p, o) for p, o in zip(templates(g), g.describe()["ops"]) if o["phase"] == "chunk"]
        macs = sum(2.0 ** (p.nC + p.nK) for p, _ in ops)
        untiled_long_k = [o["name"] for p, o in ops if p.nK >= 5 and p.ma + p.nb == 0 and p.nC >= 8]
        return g.replan_info["bytes"], macs, untiled_long_k

    b_new, m_new, bad_new = plan(None)
    b_old, m_old, _ = plan("0")
    assert m_new < 0.9 * m_old                      # measured here: 0.78
    assert b_new < 1.15 * b_old                     # measured here: 1.06
    assert not bad_new
