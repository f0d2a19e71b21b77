"""Open networks: ``contract_tn!`` on a network built with ``no_output=True`` returns the tensor over the open
wires -- the form in which the reference's own tests pin the contraction
(/root/reference/test/test_contraction_planning.jl:50-61,111-114,126-129,142-145,158-161 and
/root/reference/test/test_tn_conversion.jl:30-48: GHZ-3 state vector [1/sqrt2, 0, 0, 0, 0, 0, 0, 1/sqrt2]).
CPU part: the oracle's tensor-valued ``contract`` against those known answers, and the library's lowering of a
tensor-valued ``save`` (root mode positions, gather order, summed slice bits) replayed by the lowered emulator."""
import numpy as np
import pytest

import lowered_emulator as emu
import qxb200 as q
from oracle import qx_oracle as orc
from qxb200.executor import Graph, expand_open_indices

S2 = 1 / np.sqrt(2)
GHZ3 = np.array([S2, 0, 0, 0, 0, 0, 0, S2])
NO_BITS = np.zeros((1, 0), dtype=np.uint8)


def open_case(circ, n_slice=0, decompose=True, no_input=False, planner="flow_cutter"):
    tnc = q.convert_to_tnc(circ, no_input=no_input, no_output=True, decompose=decompose)
    if n_slice:
        bond_groups, plan, _ = q.contraction_scheme(tnc, n_slice, time=0)
    else:
        bond_groups = None
        plan = q.min_fill_contraction_plan(tnc) if planner == "min_fill" else q.flow_cutter_contraction_plan(tnc, time=0)
    cg = q.build_compute_graph(tnc, plan, bond_groups)
    wires = [w for w in tnc.wire if w in set(cg.root_indices)]
    wires += [w for w in tnc.wire_first if w in set(cg.root_indices) and w not in tnc.wire]
    return tnc, cg, wires


def sliced_open_case(circ, n_slice=2):
    """Open network sliced on INNER bonds only (the reference's slicer knows nothing about open wires and may pick a
    hyper-edge that ends in one; see test_open_network_equals_closed_amplitudes for that case being refused)."""
    tnc = q.convert_to_tnc(circ, no_output=True)
    plan = q.min_fill_contraction_plan(tnc)
    open_ids = set(tnc.wire)
    carriers = {}
    for t in tnc.tensors.values():
        for i in t.indices:
            carriers[i] = carriers.get(i, 0) + 1
    inner = [i for i, c in sorted(carriers.items()) if c >= 2 and i not in open_ids]
    bond_groups = [[i] for i in inner[len(inner) // 2: len(inner) // 2 + n_slice]]
    assert len(bond_groups) == n_slice
    cg = q.build_compute_graph(tnc, plan, bond_groups)
    wires = [w for w in tnc.wire if w in set(cg.root_indices)]
    return tnc, cg, wires


@pytest.mark.parametrize("decompose", [True, False])
@pytest.mark.parametrize("planner", ["flow_cutter", "min_fill"])
def test_oracle_ghz3_state_vector_from_open_network(decompose, planner):
    tnc, cg, wires = open_case(q.create_test_circuit(), decompose=decompose, planner=planner)
    assert len(wires) == 3
    saved = orc.contract(orc.parse_dsl(cg.dsl()), cg.tensors)
    full = expand_open_indices(saved, cg.root_indices, wires)
    assert full.shape == (2, 2, 2)
    assert np.allclose(full.reshape(-1, order="F"), GHZ3, atol=1e-15)     # reshape(output, prod(size(output)))


def statevector(circ):
    """All amplitudes through CLOSED networks (the form tests/test_oracle.py pins), as a tensor with qubit k on axis k."""
    n = circ.num_qubits
    tnc = q.convert_to_tnc(circ)
    cg = q.build_compute_graph(tnc, q.min_fill_contraction_plan(tnc))
    allb = list(q.amplitudes_all(n))
    amps = orc.amplitudes(orc.parse_dsl(cg.dsl()), cg.tensors, allb)
    out = np.zeros((2,) * n, dtype=np.complex128)
    for b, a in zip(allb, amps):
        out[tuple(int(c) for c in b)] = a
    return out


@pytest.mark.parametrize("n_slice", [0, 2])
def test_open_network_equals_closed_amplitudes(lib_built, n_slice):
    """RQC 2x3 (no diagonal structure left on most wires) and GHZ-4 (one hyper-edge over every wire): open-network
    tensor == table of closed-network amplitudes; oracle and the library's lowering (emulated) agree with both."""
    for circ in (q.create_rqc_circuit(2, 3, 6, 3), q.create_ghz_circuit(4)):
        ref = statevector(circ)
        tnc, cg, wires = open_case(circ, n_slice=n_slice)
        txt = cg.dsl()
        sliced_open = any(ln.startswith("view") for ln in txt.splitlines()) and \
            any(set(bg) & set(cg.root_indices) for bg in q.contraction_scheme(tnc, n_slice, time=0)[0])
        if sliced_open:
            # the slicer picked a hyper-edge that ends in an open wire: the slices are then different elements of the
            # result, which neither the oracle nor the library accept as a sum
            with pytest.raises(ValueError):
                orc.contract(orc.parse_dsl(txt), cg.tensors)
            with pytest.raises(Exception) as e:
                Graph.from_dsl(txt, cg.tensors).describe()
            assert "open indices cannot be slice bonds" in str(e.value)
            continue
        saved = orc.contract(orc.parse_dsl(txt), cg.tensors)
        assert np.allclose(expand_open_indices(saved, cg.root_indices, wires), ref, atol=1e-14)
        g = Graph.from_dsl(txt, cg.tensors)
        assert g.root_dims == list(saved.shape) and g.n_outputs == 0
        d = g.describe()
        assert [m[0] for m in d["root_modes"]] == list(saved.shape)
        low = emu.amplitudes(g, cg.tensors, NO_BITS)                      # [1, prod(dims)] in Julia (column-major) order
        assert low.shape == (1, saved.size)
        assert np.allclose(low[0], saved.reshape(-1, order="F"), atol=1e-14)
        if n_slice:                                                        # partial slice ranges: fixed + batched variables
            part = emu.amplitudes(g, cg.tensors, NO_BITS, 1, g.n_slices - 1)
            want = orc.contract(orc.parse_dsl(txt), cg.tensors, slice_begin=1, slice_end=g.n_slices - 1)
            assert np.allclose(part[0], want.reshape(-1, order="F"), atol=1e-14)


def test_sliced_open_network_lowering(lib_built):
    circ = q.create_rqc_circuit(2, 3, 6, 3)
    ref = statevector(circ)
    tnc, cg, wires = sliced_open_case(circ, 3)
    cmds = orc.parse_dsl(cg.dsl())
    saved = orc.contract(cmds, cg.tensors)
    assert np.allclose(expand_open_indices(saved, cg.root_indices, wires), ref, atol=1e-14)
    g = Graph.from_dsl(cg.dsl(), cg.tensors)
    assert g.n_slices == 8
    assert np.allclose(emu.amplitudes(g, cg.tensors, NO_BITS)[0], saved.reshape(-1, order="F"), atol=1e-14)
    part = orc.contract(cmds, cg.tensors, slice_begin=1, slice_end=7)      # blocks with fixed AND batched variables
    assert np.allclose(emu.amplitudes(g, cg.tensors, NO_BITS, 1, 7)[0], part.reshape(-1, order="F"), atol=1e-14)
    sub = emu.amplitudes_subspace(g, cg.tensors, NO_BITS, [1], [1]) + emu.amplitudes_subspace(g, cg.tensors, NO_BITS, [1], [0])
    assert np.allclose(sub[0], saved.reshape(-1, order="F"), atol=1e-14)


def test_open_inputs_and_outputs_give_the_unitary(lib_built):
    """no_input and no_output: the saved tensor is the circuit's matrix (wires: outputs then inputs)."""
    circ = q.create_test_circuit()
    tnc, cg, wires = open_case(circ, no_input=True)
    saved = orc.contract(orc.parse_dsl(cg.dsl()), cg.tensors)
    full = expand_open_indices(saved, cg.root_indices, wires)
    n = circ.num_qubits
    assert full.ndim == 2 * n
    col0 = full[(slice(None),) * n + (0,) * n]                            # inputs |000>
    assert np.allclose(col0.reshape(-1, order="F"), GHZ3, atol=1e-15)
    U = full.reshape(2 ** n, 2 ** n)
    assert np.allclose(U.conj().T @ U, np.eye(2 ** n), atol=1e-14)
    g = Graph.from_dsl(cg.dsl(), cg.tensors)
    low = emu.amplitudes(g, cg.tensors, NO_BITS)
    assert np.allclose(low[0], saved.reshape(-1, order="F"), atol=1e-14)


def test_random_programs_with_open_root(lib_built):
    """Random tensor-network programs (extents 2..3: zero-padded modes; hyper-indices; slices) whose last ncon keeps
    some labels: the lowering's root gather must reproduce the oracle's saved tensor element for element."""
    from cases import random_program
    seen = 0
    for seed in range(40):
        txt, data, bitstrings = random_program(seed, n_out=2, n_slice=2)
        lines = txt.strip().splitlines()
        last = next(i for i in range(len(lines) - 1, -1, -1) if lines[i].startswith("ncon "))
        tok = lines[last].split()
        a_l = [int(x) for x in tok[4].split(",")] if tok[4] != "0" else []
        b_l = [int(x) for x in tok[6].split(",")] if tok[6] != "0" else []
        keep = sorted(set(a_l) | set(b_l))[:2]
        if not keep:
            continue
        tok[2] = ",".join(map(str, keep))
        lines[last] = " ".join(tok)
        open_txt = "\n".join(lines) + "\n"
        cmds = orc.parse_dsl(open_txt)
        try:
            want = np.stack([orc.contract(cmds, data, b) for b in bitstrings[:3]])
        except ValueError:
            continue                                                       # label kept that the oracle rejects (private, summed)
        g = Graph.from_dsl(open_txt, data)
        bits = np.array([[int(c) for c in b] for b in bitstrings[:3]], dtype=np.uint8)
        low = emu.amplitudes(g, data, bits)
        assert low.shape == (3, want[0].size)
        assert np.allclose(low, np.stack([w.reshape(-1, order="F") for w in want]), atol=1e-12), seed
        seen += 1
    assert seen >= 10
