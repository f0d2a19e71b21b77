"""Open networks on the GPU: ``contract_tn!`` with ``no_output=True`` -- the reference's own known-answer form for
the contraction (/root/reference/test/test_contraction_planning.jl:50-61,111-114,126-129,142-145,158-161;
/root/reference/test/test_tn_conversion.jl:30-48) -- through the tensor-valued ``save`` of the library
(``qxb_graph_root_dims``, ``reduce_root_open_kernel``), checked against those answers and the numpy oracle."""
import numpy as np
import pytest

import qxb200 as q
from cases import random_program
from oracle import qx_oracle as orc
from qxb200.executor import Graph, expand_open_indices
from test_open_network import GHZ3, open_case, sliced_open_case, statevector

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dtype,tol", [("c64", 1e-12), ("c32", 1e-6)])
@pytest.mark.parametrize("decompose", [True, False])
def test_contract_tn_ghz3_state_vector(gpu, dtype, tol, decompose):
    """contract_tn!(tnc, plan) under the reference's planners; reshape(output, 8) == [1/sqrt2, 0, ..., 0, 1/sqrt2]."""
    circ = q.create_test_circuit()
    tnc = q.convert_to_tnc(circ, no_input=False, no_output=True, decompose=decompose)
    before = {k: t.indices for k, t in tnc.tensors.items()}
    for plan in (q.flow_cutter_contraction_plan(tnc, time=0), q.min_fill_contraction_plan(tnc),
                 q.flow_cutter_contraction_plan(tnc, time=0, hypergraph=True)):
        out = q.contract_tn(tnc, plan, dtype=dtype)
        assert out.shape == (2, 2, 2)
        assert np.allclose(out.reshape(-1, order="F"), GHZ3, atol=tol)
    assert {k: t.indices for k, t in tnc.tensors.items()} == before        # the caller's network is untouched
    closed = q.convert_to_tnc(circ)                                         # closed network: 1-element result (simulation.jl:89-90)
    out = q.contract_tn(closed, q.min_fill_contraction_plan(closed), dtype=dtype)
    assert out.shape == (1,) and abs(out[0] - 1 / np.sqrt(2)) < tol


@pytest.mark.parametrize("dtype,tol", [("c64", 1e-10), ("c32", 1e-5)])
def test_open_rqc_equals_closed_amplitudes(gpu, dtype, tol):
    for circ in (q.create_rqc_circuit(2, 3, 6, 3), q.create_ghz_circuit(4), q.create_rqc_circuit(3, 3, 8, 42)):
        n = circ.num_qubits
        tnc, cg, wires = open_case(circ)
        g = Graph.from_compute_graph(cg, dtype).compile()
        out = g.amplitudes([""])
        assert list(out.shape[1:]) == g.root_dims
        ref = orc.contract(orc.parse_dsl(cg.dsl()), cg.tensors)
        scale = max(np.max(np.abs(ref)), 2.0 ** (-n / 2))
        assert np.max(np.abs(out[0] - ref)) / scale < tol
        if n <= 6:
            assert np.max(np.abs(expand_open_indices(out[0], cg.root_indices, wires) - statevector(circ))) / scale < tol
        assert g.stats()["kernel_launches"] > 0


def test_sliced_open_network(gpu):
    """Slice bonds inside the network, open wires outside: full range, partial ranges and the sub-space form."""
    for circ, n_slice in ((q.create_rqc_circuit(2, 3, 6, 3), 3), (q.create_rqc_circuit(3, 3, 8, 7), 2)):
        tnc, cg, wires = sliced_open_case(circ, n_slice)
        cmds = orc.parse_dsl(cg.dsl())
        g = Graph.from_compute_graph(cg, "c64").compile()
        S = g.n_slices
        assert S == 2 ** n_slice
        ref = orc.contract(cmds, cg.tensors)
        assert np.max(np.abs(g.amplitudes([""])[0] - ref)) < 1e-12
        part = orc.contract(cmds, cg.tensors, slice_begin=1, slice_end=S - 1)
        assert np.max(np.abs(g.amplitudes([""], 1, S - 1)[0] - part)) < 1e-12
        total = sum(g.amplitudes_subspace([""], [0], [v])[0] for v in range(g.slice_dims[0]))
        assert np.max(np.abs(total - ref)) < 1e-12


@pytest.mark.parametrize("dtype,tol", [("c64", 1e-10), ("c32", 2e-4)])      # as tests/test_fuzz.py: cancellations
def test_random_programs_with_open_root(gpu, dtype, tol):
    """Tensor-valued save on random programs: padded (extent-3) modes, hyper-indices, slice variables still open in the
    root, several bitstrings per call."""
    seen = 0
    for seed in range(30):
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
            want = np.stack([orc.contract(cmds, data, b) for b in bitstrings])
        except ValueError:
            continue
        g = Graph.from_dsl(open_txt, data, dtype).compile()
        out = g.amplitudes(bitstrings)
        assert out.shape == want.shape
        scale = max(float(np.max(np.abs(want))), 1e-3)
        assert np.max(np.abs(out - want)) / scale < tol, seed
        seen += 1
    assert seen >= 8
