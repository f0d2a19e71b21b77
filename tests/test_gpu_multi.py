"""Several GPUs behind the C ABI (qxb_multi_*, include/qxb200.h): results must equal the single-GPU amplitudes --
bitstring shards (sub_comm_size 1: nothing is summed) bit for bit, slice splits (partial sums over disjoint slice
ranges) to ComplexF64 rounding -- and the oracle.  The one-device cases run on any GPU box; the two-device cases need
`gpurun --gpus 2` (skipped otherwise)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from qxb200.executor import Graph, MultiGraph
from oracle import qx_oracle as orc
from cases import rqc_case, rel_err

pytestmark = pytest.mark.gpu


def n_gpus():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("n_dev", [1, 2])
@pytest.mark.parametrize("dtype", ["c64", "c32"])
def test_multi_equals_single(gpu, n_dev, dtype):
    if n_gpus() < n_dev:
        pytest.skip(f"needs {n_dev} GPUs")
    txt, data, bs = rqc_case(3, 4, 10, 3, n_amp=37)
    ref = orc.amplitudes(orc.parse_dsl(txt), data, bs)
    single = Graph.from_dsl(txt, data, dtype).compile().amplitudes(bs)
    m = MultiGraph(Graph.from_dsl(txt, data, dtype), n_devices=n_dev)
    assert m.n_devices == n_dev
    tol = 1e-10 if dtype == "c64" else 1e-5
    got = m.amplitudes(bs)                                   # auto: bitstring shards
    assert np.array_equal(got, single)                       # same kernels on the same rows: bit for bit
    assert rel_err(got, ref, 12) < tol
    got_s = m.amplitudes(bs, sub_comm_size=n_dev)            # every device all bitstrings, the slices split
    assert np.max(np.abs(got_s - single)) <= (1e-12 if dtype == "c64" else 1e-5) * np.max(np.abs(single))
    S = m.n_slices
    ref2 = orc.amplitudes(orc.parse_dsl(txt), data, bs, slice_begin=1, slice_end=S - 1)
    assert rel_err(m.amplitudes(bs, 1, S - 1, sub_comm_size=n_dev), ref2, 12) < tol
    one = m.amplitudes(bs[:1])                               # fewer amplitudes than devices: auto shares them, slices split
    assert rel_err(one, ref[:1], 12) < tol


def test_multi_bench_workload_two_devices(gpu):
    if n_gpus() < 2:
        pytest.skip("needs 2 GPUs")
    import bench
    txt, data, w = bench.build_workload("rqc_7x7_d20_c64_s4096")
    bits = bench.synth_bits(2048, 49)
    g = Graph.from_dsl(txt, data, "c64", replan=32, replan_n_amp=131072)
    plan = g.text
    m = MultiGraph(g, n_devices=2)
    single = Graph.from_dsl(plan, data, "c64").compile().amplitudes(bits)
    assert np.max(np.abs(m.amplitudes(bits) - single)) <= 1e-12 * np.max(np.abs(single))
    assert np.max(np.abs(m.amplitudes(bits[:64], sub_comm_size=2) - single[:64])) <= 1e-12 * np.max(np.abs(single))


def test_qxrun_multi_flag(gpu, tmp_path):
    """bin/qxrun -m [-s K] (bin/qxrun.jl:40-46 of the reference): same results file as the single-device run."""
    import qxb200 as q
    from qxb200.jld2 import load_jld2
    here = os.path.dirname(os.path.abspath(__file__))
    exe = os.path.join(os.path.dirname(here), "qxtools.jl_b200", "bin", "qxrun")
    prefix = str(tmp_path / "rqc")
    q.generate_simulation_files(q.create_rqc_circuit(3, 3, 8, 42), prefix, 3, seed=42, time=0,
                                output_args=q.output_params_dict(9, 12, seed=5))
    outs = {}
    for tag, flags in (("one", []), ("multi", ["-m"]), ("multi_s", ["-m", "-s", str(max(1, n_gpus()))])):
        out = str(tmp_path / f"{tag}.jld2")
        r = subprocess.run([exe, "-d", prefix + ".qx", "-o", out, "--dtype", "c64"] + flags, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        outs[tag] = load_jld2(out)["amplitudes"]
    assert np.array_equal(outs["one"], outs["multi"])
    assert np.max(np.abs(outs["one"] - outs["multi_s"])) <= 1e-12 * np.max(np.abs(outs["one"]))
