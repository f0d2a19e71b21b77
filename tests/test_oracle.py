"""Pin the oracle against every known answer the reference holds for the path
(SURVEY.md 8c): docs DSL example (KAT-0), README GHZ-3, GHZ-3 state vector,
GHZ-5 over all bitstrings -- plus self-consistency the reference never tests
(sliced == unsliced, norm, QFT moduli)."""
import json
import os

import numpy as np
import pytest

import qxb200 as q
from oracle import qx_oracle as orc
from cases import kat0, rqc_case, circuit_case

HERE = os.path.dirname(os.path.abspath(__file__))
S2 = 1 / np.sqrt(2)


def test_kat0_docs_program():
    """docs/src/users_guide.md:71-90: sum over the 4 slices of the 2-qubit GHZ."""
    txt, data = kat0()
    cmds = orc.parse_dsl(txt)
    assert orc.slice_dims(cmds) == [("v1", 2), ("v2", 2)]
    got = orc.amplitudes(cmds, data, ["00", "11", "01", "10"])
    assert np.allclose(got, [S2, S2, 0, 0], atol=1e-15)


def test_slice_enumeration_v1_fastest():
    dims = [("v1", 2), ("v2", 3), ("v3", 2)]
    seen = [tuple(orc.slice_values(s, dims).values()) for s in range(12)]
    assert seen[0] == (0, 0, 0) and seen[1] == (1, 0, 0) and seen[2] == (0, 1, 0) and seen[6] == (0, 0, 1)
    assert len(set(seen)) == 12


def test_ghz3_readme_amplitudes():
    """README.md:38-50: single_amplitude "000"/"111" -> 1/sqrt2, "100" -> 0."""
    txt, data, _ = circuit_case(q.create_ghz_circuit(3))
    got = orc.amplitudes(orc.parse_dsl(txt), data, ["000", "111", "100"])
    assert np.allclose(got, [S2, S2, 0], atol=1e-15)


@pytest.mark.parametrize("decompose", [True, False])
def test_ghz3_state_vector(decompose):
    """test/test_contraction_planning.jl:58-61, test/test_tn_conversion.jl:36-48."""
    txt, data, _ = circuit_case(q.create_test_circuit(), decompose=decompose)
    allb = list(q.amplitudes_all(3))
    got = orc.amplitudes(orc.parse_dsl(txt), data, allb)
    assert np.allclose(got, [S2, 0, 0, 0, 0, 0, 0, S2], atol=1e-15)


def test_ghz5_all_bitstrings():
    """test/test_simulation.jl:16-26."""
    txt, data, _ = circuit_case(q.create_ghz_circuit(5))
    allb = list(q.amplitudes_all(5))
    got = orc.amplitudes(orc.parse_dsl(txt), data, allb)
    assert len(got) == 32 and abs(got[0] - S2) < 1e-15 and abs(got[-1] - S2) < 1e-15
    assert abs(np.sum(np.abs(got) ** 2) - 1) < 1e-14


def test_sliced_equals_unsliced_and_norm():
    circ = q.create_rqc_circuit(3, 3, 8, 5)
    t0, d0, _ = circuit_case(circ, n_slice=0)
    t1, d1, _ = circuit_case(circ, n_slice=3)
    allb = list(q.amplitudes_all(9))
    a0 = orc.amplitudes(orc.parse_dsl(t0), d0, allb)
    assert abs(np.sum(np.abs(a0) ** 2) - 1) < 1e-12
    sub = allb[::37]
    a1 = orc.amplitudes(orc.parse_dsl(t1), d1, sub)
    assert np.allclose(a1, a0[::37], atol=1e-14)


def test_qft_moduli():
    n = 6
    txt, data, _ = circuit_case(q.create_qft_circuit(n))
    got = orc.amplitudes(orc.parse_dsl(txt), data, list(q.amplitudes_all(n))[::5])
    assert np.allclose(np.abs(got), 2.0 ** (-n / 2), atol=1e-14)


def test_batch_and_scalar_labels():
    A = np.arange(8).reshape(2, 2, 2) + 1j
    B = np.arange(4).reshape(2, 2) - 2j
    # label 2 in A, B and C -> batch (users_guide.md:146)
    got = orc.ncon_pair(A, [1, 2, 3], B, [3, 2], [1, 2])
    assert np.allclose(got, np.einsum("abc,cb->ab", A, B))
    s = np.array(2.0 + 1j)
    assert np.allclose(orc.ncon_pair(s, [], B, [1, 2], [1, 2]), s * B)


def test_golden_fixture_matches_oracle():
    """Committed golden vectors (tests/golden/make_golden.py)."""
    path = os.path.join(HERE, "golden", "golden_rqc.json")
    gold = json.load(open(path))
    for case in gold["cases"]:
        txt = open(os.path.join(HERE, "golden", case["qx"])).read()
        data = dict(np.load(os.path.join(HERE, "golden", case["npz"])))
        got = orc.amplitudes(orc.parse_dsl(txt), data, case["bitstrings"])
        ref = np.array(case["re"]) + 1j * np.array(case["im"])
        assert np.allclose(got, ref, rtol=0, atol=1e-14)
