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
