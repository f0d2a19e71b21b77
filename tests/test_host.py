"""Host-side mirror of the reference API: the reference's own test expectations
(test/test_tn_conversion.jl, test/test_compute_graph.jl, test/test_simulation.jl,
test/test_contraction_planning.jl, test/test_bin.jl) restated on this repo."""
import os

import numpy as np
import pytest
import yaml

import qxb200 as q
from qxb200.compute_graph import TensorCache, save_cache
from qxb200.circuits import H


def test_tn_conversion_counts():
    """test/test_tn_conversion.jl:6-28."""
    circ = q.create_test_circuit()
    assert len(q.convert_to_tnc(circ, no_input=True, no_output=True, decompose=False)) == 3
    assert len(q.convert_to_tnc(circ, no_input=True, no_output=True, decompose=True)) == 5
    tnc = q.convert_to_tnc(circ, decompose=True)
    assert len(tnc) == 11
    first = next(iter(tnc.keys()))
    assert np.allclose(tnc.tensor_data(first), H)


def test_hyperindices_of_cx():
    """The CX control half is a pure hyper-index: stored rank-1 (users_guide.md:78 'load t3 data_4 2')."""
    tnc = q.convert_to_tnc(q.create_ghz_circuit(2), no_input=True, no_output=True)
    shapes = sorted(t.shape for t in tnc.tensors.values())
    assert shapes == [(2,), (2, 2), (2, 2, 2)]


def test_tensor_cache():
    """test/test_compute_graph.jl:5-24."""
    tc = TensorCache()
    rng = np.random.default_rng(0)
    a = rng.random((3, 4, 5))
    sym = tc.push(a)
    assert sym == tc.push(a)
    assert sym != tc.push(rng.random((3, 4, 5)))
    eps = np.finfo(np.float64).eps
    assert sym == tc.push(a + 0.1 * eps)
    assert sym != tc.push(a + 1.1 * eps)
    assert np.array_equal(tc[sym], a.astype(np.complex128))
    assert len(tc) == 3
    assert list(tc.key_dim_map) == ["data_1", "data_2", "data_3"]


def test_tensor_cache_save(tmp_path):
    tc = TensorCache()
    a = np.arange(6.0).reshape(2, 3)
    sym = tc.push(a)
    # test/test_compute_graph.jl:19-22: save_cache to tmp.jld2, load, compare every key
    from qxb200.jld2 import load_jld2
    save_cache(tc, str(tmp_path / "tmp.jld2"))
    got = load_jld2(str(tmp_path / "tmp.jld2"))
    assert list(got) == [sym] and np.array_equal(got[sym], a) and got[sym].dtype == np.complex128
    save_cache(tc, str(tmp_path / "t.npz"))
    assert np.array_equal(np.load(tmp_path / "t.npz")[sym], a)
    with pytest.raises(ValueError):
        save_cache(tc, str(tmp_path / "t.jld"))


def test_compute_graph_node_count():
    """test/test_compute_graph.jl:26-34: nodes = tensors + plan + 1."""
    tnc = q.convert_to_tnc(q.create_test_circuit())
    plan = q.min_fill_contraction_plan(tnc)
    cg = q.build_compute_graph(tnc, plan)
    assert len(cg.root) == len(tnc) + len(plan) + 1
    assert len(plan) == len(tnc) - 1


def test_compute_graph_views_and_dsl():
    tnc = q.convert_to_tnc(q.create_rqc_circuit(3, 3, 8, 1))
    bg, plan, meta = q.contraction_scheme(tnc, 2, time=0)
    assert len(bg) == 2
    cg = q.build_compute_graph(tnc, plan, bg)
    cmds = cg.commands()
    views = [c for c in cmds if isinstance(c, q.ViewCommand)]
    assert {v.slice_sym for v in views} == {"v1", "v2"}
    for v in views:                       # compute_graph.jl:45 naming, :43 extent
        assert v.name == v.target + "_s" and v.bond_dim == 2
    txt = cg.dsl(meta)
    assert txt.startswith("# version: 0.4.0\n")
    assert txt.rstrip().splitlines()[-1].startswith("save output ")
    # every name is defined before use (users_guide.md:71-90)
    defined = set()
    for ln in txt.splitlines():
        t = ln.split()
        if not t or t[0].startswith("#"):
            continue
        if t[0] == "view":
            assert t[2] in defined
        if t[0] == "ncon":
            assert t[3] in defined and t[5] in defined
        if t[0] in ("load", "output", "view", "ncon"):
            defined.add(t[1])


def test_amplitude_generators():
    """test/test_simulation.jl:4-14."""
    amps = list(q.amplitudes_all(5))
    assert len(amps) == 32 and all(len(a) == 5 for a in amps)
    assert amps[0] == "00000" and amps[1] == "00001" and amps[-1] == "11111"
    assert len(list(q.amplitudes_uniform(5, None, 10))) == 10
    assert list(q.amplitudes_uniform(7, 3, 4)) == list(q.amplitudes_uniform(7, 3, 4))


def test_line_graph_and_plans():
    """Shapes of test/test_contraction_planning.jl: plans of the GHZ-3 network have
    len(tnc)-1 steps whichever planner built them."""
    tnc = q.convert_to_tnc(q.create_test_circuit())
    lg = q.convert_to_line_graph(tnc)
    assert set(lg) == set(tnc.index_dim)
    for plan in (q.min_fill_contraction_plan(tnc), q.flow_cutter_contraction_plan(tnc, time=0),
                 q.flow_cutter_contraction_plan(tnc, time=0.2, seed=3)):
        assert len(plan) == len(tnc) - 1
        outs = [c for _, _, c in plan]
        assert len(set(outs)) == len(outs)


def test_contraction_scheme_slices():
    tnc = q.convert_to_tnc(q.create_rqc_circuit(4, 4, 12, 42))
    bg, plan, meta = q.contraction_scheme(tnc, 3, time=0)
    assert len(bg) == 3 and len({b[0] for b in bg}) == 3
    tws = meta["Slicing"]["Treewidths after slicing consecutive edges"]
    assert len(tws) == 3 and tws == sorted(tws, reverse=True)
    bg0, plan0, _ = q.contraction_scheme(tnc, 0, time=0)
    assert bg0 == [] and len(plan0) > 0


def test_generate_simulation_files(tmp_path):
    """test/test_bin.jl:10-52 (file triple + YAML schema, outputs.jl:47-78)."""
    prefix = str(tmp_path / "rqc_3_3_8")
    circ = q.create_rqc_circuit(3, 3, 8, 42)
    q.generate_simulation_files(circ, prefix, 2, seed=42, time=0,
                                output_args=q.output_params_dict(9, 15, seed=1))
    for ext in (".qx", ".jld2", ".yml"):                       # test/test_bin.jl:18
        assert os.path.exists(prefix + ext)
    y = yaml.safe_load(open(prefix + ".yml"))
    assert y["output"]["method"] == "List"
    assert y["output"]["params"]["num_samples"] == 15 and len(y["output"]["params"]["bitstrings"]) == 15
    q.generate_simulation_files(circ, prefix, 2, seed=42, time=0,
                                output_args=q.output_params_dict(9, 20, output_method="Rejection"))
    y = yaml.safe_load(open(prefix + ".yml"))
    assert y["output"]["method"] == "Rejection" and y["output"]["params"]["num_samples"] == 20
    assert set(y["output"]["params"]) == {"num_qubits", "M", "fix_M", "seed", "num_samples"}
    with pytest.raises(ValueError):
        q.output_params_dict(9, 3, output_method="Nope")
