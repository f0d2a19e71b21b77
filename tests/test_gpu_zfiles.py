"""The file seam on the GPU: a `.qx` / `.jld2` / `.yml` triple executed by the library alone
(`qxb_execute_files`, the plain-C `qxrun`), checked against the numpy oracle on the same files.
Mirrors `julia --project bin/qxrun.jl -d X.qx -o out.jld2` (/root/reference/docs/src/distributed.md:30-33)
and the `-a` / `-n` truncation of /root/reference/bin/qxrun.jl:32-39."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import qxb200 as q
from oracle import qx_oracle as orc
from qxb200._lib import load
from qxb200.jld2 import load_jld2, read_params

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
QXRUN = os.path.join(ROOT, "qxtools.jl_b200", "bin", "qxrun")


def rel_err(got, ref, n_qubits):
    scale = max(float(np.max(np.abs(ref))) if len(ref) else 0.0, 2.0 ** (-n_qubits / 2))
    return float(np.max(np.abs(np.asarray(got) - np.asarray(ref)))) / scale if len(ref) else 0.0


@pytest.fixture()
def triple(tmp_path):
    prefix = str(tmp_path / "rqc_3_3_8")
    q.generate_simulation_files(q.create_rqc_circuit(3, 3, 8, 42), prefix, 3, seed=42, time=0,
                                output_args=q.output_params_dict(9, 6, seed=5))
    txt = open(prefix + ".qx").read()
    data = dict(load_jld2(prefix + ".jld2"))
    bitstrings = read_params(prefix + ".yml")["bitstrings"]
    return prefix, orc.parse_dsl(txt), data, bitstrings


@pytest.mark.parametrize("dtype,tol", [(1, 1e-10), (0, 1e-5)])
def test_execute_files_matches_oracle(gpu, triple, dtype, tol):
    prefix, cmds, data, bitstrings = triple
    lib = load()
    n = C.c_int64()
    sec = (C.c_double * 4)()
    out = prefix + "_out.jld2"
    rc = lib.qxb_execute_files((prefix + ".qx").encode(), None, None, out.encode(), dtype, -1, -1, 0, C.byref(n), sec)
    assert rc == 0, lib.qxb_last_error()
    assert n.value == len(bitstrings) == 6 and all(s >= 0 for s in sec)
    res = load_jld2(out)
    assert [b.decode() for b in res["bitstrings"]] == bitstrings
    assert res["amplitudes"].dtype == (np.complex128 if dtype == 1 else np.complex64)
    ref = orc.amplitudes(cmds, data, bitstrings)
    assert rel_err(res["amplitudes"], ref, 9) < tol


def test_execute_files_truncation_and_replan(gpu, triple):
    """-a keeps the first N bitstrings, -n the first S slices (qxrun.jl:32-39); re-planning does not change values."""
    prefix, cmds, data, bitstrings = triple
    lib = load()
    n = C.c_int64()
    out = prefix + "_out2.jld2"
    rc = lib.qxb_execute_files((prefix + ".qx").encode(), (prefix + ".jld2").encode(), (prefix + ".yml").encode(),
                               out.encode(), 1, 2, 3, 0, C.byref(n), None)
    assert rc == 0, lib.qxb_last_error()
    res = load_jld2(out)
    ref = orc.amplitudes(cmds, data, bitstrings[:2], slice_begin=0, slice_end=3)
    assert n.value == 2 and rel_err(res["amplitudes"], ref, 9) < 1e-10
    rc = lib.qxb_execute_files((prefix + ".qx").encode(), None, None, out.encode(), 1, -1, -1, 8, C.byref(n), None)
    assert rc == 0, lib.qxb_last_error()
    assert rel_err(load_jld2(out)["amplitudes"], orc.amplitudes(cmds, data, bitstrings), 9) < 1e-10


def test_qxrun_binary(gpu, triple):
    prefix, cmds, data, bitstrings = triple
    if not os.path.exists(QXRUN):                            # normally built by __graft_entry__.build() (csrc/Makefile)
        subprocess.run(["make", "-C", os.path.join(ROOT, "qxtools.jl_b200", "csrc"), "../bin/qxrun"], check=True)
    try:
        os.chmod(QXRUN, 0o755)                               # the snapshot that reaches the GPU box may drop the mode bits
    except OSError:
        pass
    out = prefix + "_cli.jld2"
    r = subprocess.run([QXRUN, "-d", prefix + ".qx", "-o", out, "--dtype", "c64", "-t", "-g", "-b", "1"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    assert "Simulation" in r.stdout and "amplitudes/s" in r.stdout
    res = load_jld2(out)
    assert [b.decode() for b in res["bitstrings"]] == bitstrings
    assert rel_err(res["amplitudes"], orc.amplitudes(cmds, data, bitstrings), 9) < 1e-10


def test_python_execute_reads_and_writes_jld2(gpu, triple):
    from qxb200.execute import execute
    prefix, cmds, data, bitstrings = triple
    res = execute(prefix + ".qx", output_file=prefix + "_py.jld2", dtype="c64")
    back = load_jld2(prefix + "_py.jld2")
    assert [b.decode() for b in back["bitstrings"]] == list(res.keys()) == bitstrings
    assert np.array_equal(back["amplitudes"], np.array(list(res.values())))
    assert rel_err(back["amplitudes"], orc.amplitudes(cmds, data, bitstrings), 9) < 1e-10


def test_native_rejection_sampler(gpu, tmp_path):
    """A Rejection parameter file (outputs.jl:57-62) through the native runner: GHZ-5 yields only its two outcomes, each
    with |amplitude| = 1/sqrt2; fix_M keeps M; the file also records M and the number of candidates drawn."""
    prefix = str(tmp_path / "ghz5")
    q.generate_simulation_files(q.create_ghz_circuit(5), prefix, 1, time=0,
                                output_args=q.output_params_dict(5, 12, output_method="Rejection", M=16.0, fix_M=True, seed=4))
    lib = load()
    n = C.c_int64()
    out = prefix + "_out.jld2"
    rc = lib.qxb_execute_files((prefix + ".qx").encode(), None, None, out.encode(), 1, -1, -1, 0, C.byref(n), None)
    assert rc == 0, lib.qxb_last_error()
    res = load_jld2(out)
    assert n.value == 12 and len(res["bitstrings"]) == 12
    assert set(b.decode() for b in res["bitstrings"]) <= {"00000", "11111"}
    assert np.allclose(np.abs(res["amplitudes"]), 1 / np.sqrt(2), atol=1e-12)
    assert float(res["M"]) == 16.0 and float(res["drawn"]) >= 12
    # frugal mode: M rises to the largest p 2^n seen (16 for GHZ-5), -a truncates the number of samples
    q.generate_parameter_file(prefix, q.output_params_dict(5, 12, output_method="Rejection", M=0.0001, fix_M=False, seed=5))
    rc = lib.qxb_execute_files((prefix + ".qx").encode(), None, None, out.encode(), 0, 5, -1, 0, C.byref(n), None)
    assert rc == 0, lib.qxb_last_error()
    res = load_jld2(out)
    assert n.value == 5 and abs(float(res["M"]) - 16.0) < 1e-4 and res["amplitudes"].dtype == np.complex64


def test_execute_with_autotune(gpu, triple):
    """execute(..., autotune=True): the measured choice among re-planned trees / register-tile knobs must not change
    a single amplitude beyond rounding (every candidate is an exact re-association run by the same kernels)."""
    from qxb200.execute import execute
    prefix, cmds, data, bitstrings = triple
    env_before = dict(os.environ)
    res = execute(prefix + ".qx", dtype="c64", autotune=True)
    assert dict(os.environ) == env_before                  # the winning knobs travel as qxb_options, not as environment
    assert list(res.keys()) == bitstrings
    assert rel_err(np.array(list(res.values())), orc.amplitudes(cmds, data, bitstrings), 9) < 1e-10
