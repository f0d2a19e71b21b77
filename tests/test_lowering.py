"""The lowered program (bit-segment maps, arena plan, block decomposition) executed
by the numpy emulator must reproduce the oracle -- this is the CPU-side proof that
what the CUDA executor is told to do is right."""
import json
import os

import numpy as np
import pytest

import qxb200 as q
from qxb200.executor import Graph, bits_from_strings
from oracle import qx_oracle as orc
import lowered_emulator as em
from cases import kat0, rqc_case, circuit_case

HERE = os.path.dirname(os.path.abspath(__file__))


def test_kat0_all_ranges(lib_built):
    txt, data = kat0()
    g = Graph.from_dsl(txt, data)
    bs = ["00", "11", "01", "10"]
    bits = bits_from_strings(bs, 2)
    cmds = orc.parse_dsl(txt)
    for b in range(4):
        for e in range(b + 1, 5):
            ref = orc.amplitudes(cmds, data, bs, slice_begin=b, slice_end=e)
            assert np.allclose(em.amplitudes(g, data, bits, b, e), ref, atol=1e-15)


@pytest.mark.parametrize("shape", [(3, 3, 8, 2), (3, 4, 10, 3), (4, 4, 12, 4)])
def test_rqc_emulated_lowering(lib_built, shape):
    r, c, d, ns = shape
    txt, data, bs = rqc_case(r, c, d, ns, n_amp=5)
    g = Graph.from_dsl(txt, data)
    bits = bits_from_strings(bs, r * c)
    cmds = orc.parse_dsl(txt)
    assert np.allclose(em.amplitudes(g, data, bits), orc.amplitudes(cmds, data, bs), atol=1e-14)
    S = g.n_slices
    for (b, e) in [(1, S - 1), (S // 2, S), (3, 4)]:
        ref = orc.amplitudes(cmds, data, bs, slice_begin=b, slice_end=e)
        assert np.allclose(em.amplitudes(g, data, bits, b, e), ref, atol=1e-14)


def test_block_decomposition_covers_range():
    dims = [2, 3, 2, 4]
    total = 48
    for b in range(0, total, 5):
        for e in range(b, total + 1, 7):
            seen = []
            for n_free, vals in em.decompose(dims, b, e):
                place = 1
                base = 0
                for i, d in enumerate(dims):
                    if i >= n_free:
                        base += vals[i] * place
                    else:
                        assert vals[i] == 0
                    place *= d
                span = int(np.prod(dims[:n_free])) if n_free else 1
                seen.extend(range(base, base + span))
            assert seen == list(range(b, e))


def test_non_power_of_two_extents(lib_built):
    """Leaves and slice variables with extent 3: padded to 4 with zeros."""
    rng = np.random.default_rng(3)
    A = rng.normal(size=(3, 2)) + 1j * rng.normal(size=(3, 2))
    B = rng.normal(size=(3, 3)) + 1j * rng.normal(size=(3, 3))
    Cc = rng.normal(size=(3, 2)) + 1j * rng.normal(size=(3, 2))
    txt = ("# version: 0.4.0\nload a dA 3,2\nview a_s a v1 1 3\nload b dB 3,3\nview b_s b v1 1 3\n"
           "load c dC 3,2\noutput o1 1 2\noutput o2 2 2\n"
           "ncon ab 1,2,3 a_s 1,2 b_s 1,3\nncon abc 1,2,4 ab 1,2,3 c 3,4\nncon x 1,4 abc 1,2,4 o1 2\n"
           "ncon y 0 x 1,4 o2 4\nsave output y\n")
    data = {"dA": A, "dB": B, "dC": Cc}
    g = Graph.from_dsl(txt, data)
    assert g.slice_dims == [3] and g.n_slices == 3
    bs = ["00", "01", "10", "11"]
    bits = bits_from_strings(bs, 2)
    cmds = orc.parse_dsl(txt)
    for (b, e) in [(0, 3), (0, 1), (1, 3), (2, 3)]:
        ref = orc.amplitudes(cmds, data, bs, slice_begin=b, slice_end=e)
        assert np.allclose(em.amplitudes(g, data, bits, b, e), ref, atol=1e-14)


def test_golden_vectors_through_lowering(lib_built):
    gold = json.load(open(os.path.join(HERE, "golden", "golden_rqc.json")))
    for case in gold["cases"][:2]:
        txt = open(os.path.join(HERE, "golden", case["qx"])).read()
        data = dict(np.load(os.path.join(HERE, "golden", case["npz"])))
        g = Graph.from_dsl(txt, data)
        bits = bits_from_strings(case["bitstrings"], case["n_qubits"])
        ref = np.array(case["re"]) + 1j * np.array(case["im"])
        assert np.allclose(em.amplitudes(g, data, bits), ref, atol=1e-14)
        # a single slice of bitstring 0
        s = 1
        one = em.amplitudes(g, data, bits[:1], s, s + 1)[0]
        assert abs(one - (case["slice_re_bs0"][s] + 1j * case["slice_im_bs0"][s])) < 1e-14


def test_early_sum_and_root_sum_agree(lib_built):
    """Summing a batched slice variable at the lowest covering node (default) or at the
    root (sum_at_root) is the same number; the early form must do no more work."""
    txt, data, bs = rqc_case(4, 4, 12, 4, n_amp=4)
    bits = bits_from_strings(bs, 16)
    ref = orc.amplitudes(orc.parse_dsl(txt), data, bs)
    work = {}
    for mode in (False, True):
        g = Graph.from_dsl(txt, data).configure(sum_at_root=mode)
        d = g.describe()
        assert d["early_sum"] == (not mode)
        assert np.allclose(em.amplitudes(g, data, bits), ref, atol=1e-14)
        work[mode] = sum(2.0 ** (op["nC"] + op["nK"]) for op in d["ops"])
    assert work[False] <= work[True]


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_dependency_edges_allow_any_topological_order(lib_built, seed):
    """The step runs as a CUDA graph: any order that respects `deps` (producer edges +
    write-after-read edges of the arena plan) must give the same amplitudes."""
    txt, data, bs = rqc_case(4, 4, 12, 4, n_amp=3)
    g = Graph.from_dsl(txt, data)
    bits = bits_from_strings(bs, 16)
    ref = orc.amplitudes(orc.parse_dsl(txt), data, bs)
    assert np.allclose(em.amplitudes(g, data, bits, shuffle_seed=seed), ref, atol=1e-14)
    assert np.allclose(em.amplitudes(g, data, bits, 3, 14, shuffle_seed=seed),
                       orc.amplitudes(orc.parse_dsl(txt), data, bs, slice_begin=3, slice_end=14), atol=1e-14)


def test_memory_plan_has_no_unordered_overlap(lib_built):
    """Two tensors that share arena space must be ordered by the dependency edges:
    the later writer is a descendant of every reader of the earlier tenant."""
    txt, data, _ = rqc_case(6, 6, 16, 8, n_amp=1)
    g = Graph.from_dsl(txt, data)
    d = g.describe()
    ops, T = d["ops"], d["tensors"]
    anc = []                                   # transitive closure of deps
    for i, o in enumerate(ops):
        a = set()
        for p in o["deps"]:
            a |= anc[p] | {p}
        anc.append(a)
    prod = {o["c"]: i for i, o in enumerate(ops)}
    readers = {}
    for i, o in enumerate(ops):
        for t in (o["a"], o["b"]):
            readers.setdefault(t, []).append(i)
    size = lambda t: max(2, 1 << T[t]["span_bits"])
    inter = [t for t in prod if ops[prod[t]]["phase"] != "const"]
    n_shared = 0
    for x in inter:
        for y in inter:
            if prod[x] >= prod[y] or T[x]["phase"] != T[y]["phase"]:
                continue
            if T[x]["offset"] < T[y]["offset"] + size(y) and T[y]["offset"] < T[x]["offset"] + size(x):
                n_shared += 1
                for r in readers.get(x, []):
                    assert r in anc[prod[y]] or r == prod[y], (T[x]["name"], T[y]["name"])
    assert n_shared > 0


def test_subspace_partition_sums_to_full(lib_built):
    txt, data, bs = rqc_case(4, 4, 12, 4, n_amp=3)
    g = Graph.from_dsl(txt, data)
    bits = bits_from_strings(bs, 16)
    ref = orc.amplitudes(orc.parse_dsl(txt), data, bs)
    for n_parts in (2, 4, 8):
        total = np.zeros(len(bs), dtype=np.complex128)
        seen = set()
        for part in range(n_parts):
            fv, fx = g.partition_assignment(n_parts, part)
            assert len(fv) == int(np.log2(n_parts)) and len(set(fv)) == len(fv)
            seen.add(tuple(fx))
            total += em.amplitudes_subspace(g, data, bits, fv, fx, shuffle_seed=part)
        assert len(seen) == n_parts
        assert np.allclose(total, ref, atol=1e-14)
    assert g.partition_vars(1) == []
    assert g.partition_vars(3) == []          # extents are all 2: 3 ranks cannot be factored


def test_fsim_circuit_extent4_bonds(lib_built):
    """Sycamore-like fSim gates have operator-Schmidt rank 4: extent-4 bonds and extent-4 slice
    variables (two address bits per mode)."""
    circ = q.create_sycamore_like_circuit(4, seed=3, n_qubits=12)
    tnc = q.convert_to_tnc(circ)
    assert sorted(set(tnc.index_dim.values())) == [2, 4]
    bg, plan, _ = q.contraction_scheme(tnc, 3, time=0)
    cg = q.build_compute_graph(tnc, plan, bg)
    txt, data = cg.dsl(), dict(cg.tensors)
    bs = list(q.amplitudes_uniform(12, 1, 3))
    g = Graph.from_dsl(txt, data)
    assert 4 in g.slice_dims
    ref = orc.amplitudes(orc.parse_dsl(txt), data, bs)
    assert np.allclose(em.amplitudes(g, data, bits_from_strings(bs, 12), shuffle_seed=2), ref, atol=1e-14)
    S = g.n_slices
    assert np.allclose(em.amplitudes(g, data, bits_from_strings(bs, 12), 5, S - 3),
                       orc.amplitudes(orc.parse_dsl(txt), data, bs, slice_begin=5, slice_end=S - 3), atol=1e-14)
