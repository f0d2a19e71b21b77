"""Samplers in front of the path, with the oracle standing in for the GPU amplitudes call."""
import numpy as np

import qxb200 as q
from qxb200.samplers import rejection_sample, uniform_bitstrings, bits_to_strings
from oracle import qx_oracle as orc
from cases import circuit_case


def _amp_fn(circ):
    txt, data, _ = circuit_case(circ)
    cmds = orc.parse_dsl(txt)
    return lambda bits: orc.amplitudes(cmds, data, bits_to_strings(bits))


def test_rejection_sampling_ghz():
    """GHZ-4 has only two outcomes, each with p * 2^n = 8."""
    bs, amps, info = rejection_sample(_amp_fn(q.create_ghz_circuit(4)), 4, 40, M=8.0, fix_M=True, seed=1, batch=64)
    assert len(bs) == 40 and set(bs) <= {"0000", "1111"} and len(set(bs)) == 2
    assert np.allclose(np.abs(amps), 1 / np.sqrt(2))
    assert info["accepted"] == 40 and info["drawn"] >= 40
    # frugal mode finds the bound by itself
    bs2, _, info2 = rejection_sample(_amp_fn(q.create_ghz_circuit(4)), 4, 20, M=0.001, fix_M=False, seed=2, batch=64)
    assert set(bs2) <= {"0000", "1111"} and abs(info2["M"] - 8.0) < 1e-9


def test_rejection_sampling_follows_distribution():
    circ = q.create_rqc_circuit(2, 3, 8, 3)
    fn = _amp_fn(circ)
    allb = np.array([[int(c) for c in s] for s in q.amplitudes_all(6)], dtype=np.uint8)
    p = np.abs(fn(allb)) ** 2
    bs, _, info = rejection_sample(fn, 6, 3000, M=float(p.max() * 64) * 1.0001, fix_M=True, seed=7, batch=512)
    counts = np.zeros(64)
    for s in bs:
        counts[int(s, 2)] += 1
    # chi-square-ish: empirical frequencies within 5 sigma of the exact distribution
    sigma = np.sqrt(3000 * p * (1 - p)) + 1e-9
    assert np.all(np.abs(counts - 3000 * p) < 5 * sigma + 3)


def test_uniform_bitstrings_seeded():
    a, b = uniform_bitstrings(9, 5, 3), uniform_bitstrings(9, 5, 3)
    assert a.shape == (5, 9) and np.array_equal(a, b) and set(np.unique(a)) <= {0, 1}


def test_uniform_bitstrings_of_execute_are_the_library_stream(tmp_path):
    """`execute` (Python), `qxrun` and the Julia shim read a Uniform parameter file through `qxb_params_read`: one RNG
    stream (splitmix64, documented in include/qxb200.h), so the same file and seed give the same bitstrings everywhere."""
    import yaml
    from qxb200.execute import _bitstrings_from_params
    from qxb200.jld2 import read_params
    path = str(tmp_path / "u.yml")
    doc = {"output": {"method": "Uniform", "params": {"num_qubits": 9, "num_samples": 7, "seed": 123}}}
    with open(path, "w") as f:
        yaml.safe_dump(doc, f)
    via_execute = _bitstrings_from_params(yaml.safe_load(open(path)), None, path)
    via_library = read_params(path)["bitstrings"]
    assert via_execute == via_library and len(via_execute) == 7 and all(len(b) == 9 and set(b) <= {"0", "1"} for b in via_execute)
    assert _bitstrings_from_params(yaml.safe_load(open(path)), 3, path) == via_library[:3]          # qxrun.jl:32-39: the FIRST n
    doc["output"]["params"]["seed"] = 124
    with open(path, "w") as f:
        yaml.safe_dump(doc, f)
    assert read_params(path)["bitstrings"] != via_library
