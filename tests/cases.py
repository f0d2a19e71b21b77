"""Shared workload builders for the tests (seeded; small enough for the oracle)."""
import numpy as np

import qxb200 as q


def kat0():
    """2-qubit GHZ program of the reference docs with the tensor values of SURVEY.md Appendix B."""
    import os
    here = os.path.dirname(os.path.abspath(__file__))
    txt = open(os.path.join(here, "golden", "kat0_ghz2.qx")).read()
    s = 1 / np.sqrt(2)
    T = np.zeros((2, 2, 2))
    for o in range(2):
        for i in range(2):
            for c in range(2):
                T[o, i, c] = 1.0 if o == (i ^ c) else 0.0
    data = {"data_1": np.array([1, 0.0]), "data_2": np.array([[1, 1], [1, -1]]) * s,
            "data_3": T, "data_4": np.array([1.0, 1.0])}
    return txt, data


def rqc_case(rows, cols, depth, n_slice, seed=42, n_amp=8, amp_seed=1, **kw):
    circ = q.create_rqc_circuit(rows, cols, depth, seed)
    tnc = q.convert_to_tnc(circ)
    bond_groups, plan, meta = q.contraction_scheme(tnc, n_slice, time=0, **kw)
    cg = q.build_compute_graph(tnc, plan, bond_groups)
    bitstrings = list(q.amplitudes_uniform(rows * cols, amp_seed, n_amp))
    return cg.dsl(meta), dict(cg.tensors), bitstrings


def circuit_case(circ, n_slice=0, n_amp=8, amp_seed=1, decompose=True):
    tnc = q.convert_to_tnc(circ, decompose=decompose)
    if n_slice:
        bond_groups, plan, _ = q.contraction_scheme(tnc, n_slice, time=0)
    else:
        bond_groups, plan = None, q.min_fill_contraction_plan(tnc)
    cg = q.build_compute_graph(tnc, plan, bond_groups)
    bitstrings = list(q.amplitudes_uniform(circ.num_qubits, amp_seed, n_amp))
    return cg.dsl(), dict(cg.tensors), bitstrings


def rel_err(got, ref, n_qubits):
    """Relative to max(|amp_ref|, 2^{-n/2}) (zero amplitudes need an absolute floor)."""
    scale = max(float(np.max(np.abs(ref))), 2.0 ** (-n_qubits / 2))
    return float(np.max(np.abs(np.asarray(got) - np.asarray(ref)))) / scale
