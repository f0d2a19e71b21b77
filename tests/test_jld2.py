"""The file seam without a GPU: native JLD2 (HDF5-subset) reader / writer, the YAML-subset
parameter reader and the host part of ``qxb_execute_files`` (csrc/qxb_jld2.cpp, csrc/qxb_files.cpp).

Pins: (1) the lookup3 checksum against Jenkins' published vectors; (2) the reader against a file
written by the real HDF5 C library (scipy ships one MATLAB v7.3 file: 512-byte user block,
superblock v0, symbol-table group, v1 object headers -- the library-written counterpart of the
structures JLD2.jl writes in their v2 form); (3) the JLD2-shaped layout against this repo's own
writer (no file produced by JLD2.jl exists in the image: that part is unpinned by the reference,
see DESIGN.md §6)."""
import ctypes as C
import glob
import os
import struct
import subprocess

import numpy as np
import pytest
import yaml

import qxb200 as q
from qxb200._lib import QxbError, load
from qxb200.jld2 import load_jld2, read_params, save_jld2

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def test_lookup3_published_vectors():
    """lookup3.c, driver5(): hashlittle("Four score and seven years ago", 30, 0) = 0x17770551, initval 1 -> 0xcd628161;
    the empty key with initval 0 -> 0xdeadbeef."""
    lib = load()
    key = b"Four score and seven years ago"
    assert lib.qxb_debug_lookup3(key, len(key), 0) == 0x17770551
    assert lib.qxb_debug_lookup3(key, len(key), 1) == 0xCD628161
    assert lib.qxb_debug_lookup3(b"", 0, 0) == 0xDEADBEEF


def _scipy_hdf5_file():
    try:
        import scipy.io
    except ImportError:
        return None
    hits = glob.glob(os.path.join(os.path.dirname(scipy.io.__file__), "matlab", "tests", "data", "testhdf5_7.4_GLNX86.mat"))
    return hits[0] if hits else None


@pytest.mark.skipif(_scipy_hdf5_file() is None, reason="scipy's HDF5 test file not installed")
def test_reads_file_written_by_the_hdf5_library():
    """MATLAB 7.4 `testdouble = 0:pi/4:2*pi` saved with -v7.3: a 1 x 9 double written by libhdf5."""
    info = {}
    d = load_jld2(_scipy_hdf5_file(), info=info)
    assert list(d) == ["testdouble"] and info["skipped"] == []
    assert d["testdouble"].shape == (1, 9) and d["testdouble"].dtype == np.float64
    assert np.array_equal(d["testdouble"].ravel(), np.arange(9) * (np.pi / 4))
    as_c = load_jld2(_scipy_hdf5_file(), as_c64=True)["testdouble"]
    assert as_c.dtype == np.complex128 and np.array_equal(as_c.real.ravel(), np.arange(9) * (np.pi / 4))


def _arrays(rng):
    c = lambda *s: rng.normal(size=s) + 1j * rng.normal(size=s)
    return {
        "data_1": c(2, 2, 2, 2), "data_2": np.array([1, 0], dtype=np.complex128), "data_3": c(2, 3, 5),
        "scalar": np.array(3 - 4j), "f32": rng.normal(size=(7,)).astype(np.float32), "f64": rng.normal(size=(3, 1, 2)),
        "c32": c(4, 3).astype(np.complex64), "ints": np.arange(-3, 9).reshape(2, 6), "names": np.array([b"0101", b"1+-0"]),
        "a_much_longer_dataset_name_" + "x" * 300: c(2),
    }


@pytest.mark.parametrize("commit_types", [True, False])
def test_roundtrip(tmp_path, commit_types):
    arrays = _arrays(np.random.default_rng(7))
    path = str(tmp_path / "t.jld2")
    save_jld2(path, arrays, commit_types=commit_types)
    info = {}
    back = load_jld2(path, info=info)
    assert info == {"checksum_failures": 0, "skipped": []}
    assert list(back) == list(arrays)                        # link order = write order
    for k, v in arrays.items():
        assert back[k].shape == v.shape and back[k].dtype == (np.int64 if v.dtype.kind == "i" else v.dtype), k
        assert np.array_equal(back[k], v), k
        if v.ndim > 1:
            assert back[k].flags.f_contiguous                # Julia memory order preserved
    raw = open(path, "rb").read()
    assert raw.startswith(b"HDF5-based Julia Data Format, version ")       # JLD2's required file header
    assert raw[512:520] == b"\x89HDF\r\n\x1a\n" and raw[520] == 2          # superblock v2 behind the 512-byte header
    base, ext, eof, root = struct.unpack_from("<4Q", raw, 524)
    assert base == 512 and eof + 512 == len(raw) and raw[512 + root:516 + root] == b"OHDR"
    assert (b"_types" in raw) == commit_types


def test_layout_is_column_major(tmp_path):
    """Julia writes reverse(size(A)) as the HDF5 extents and the memory as is (column-major)."""
    a = (np.arange(24).reshape(2, 3, 4) * (1 + 0.5j)).astype(np.complex128)
    path = str(tmp_path / "a.jld2")
    save_jld2(path, {"data_1": a})
    raw = open(path, "rb").read()
    want = np.asfortranarray(a).tobytes(order="F")
    at = raw.find(want)
    assert at > 512 and at % 8 == 0
    assert struct.pack("<3Q", 4, 3, 2) in raw                # dataspace extents reversed


def test_corruption_is_reported(tmp_path):
    path = str(tmp_path / "t.jld2")
    save_jld2(path, {"data_1": np.ones((2, 2), dtype=np.complex128)})
    raw = bytearray(open(path, "rb").read())
    with pytest.raises(QxbError) as e:
        open(path, "wb").write(bytes(raw[:len(raw) - 40])), load_jld2(path)
    assert "past the end" in str(e.value) or "offset" in str(e.value)
    with pytest.raises(QxbError):
        open(path, "wb").write(b"not an hdf5 file" * 64), load_jld2(path)
    with pytest.raises(QxbError):
        load_jld2(str(tmp_path / "missing.jld2"))
    flipped = bytearray(raw)
    flipped[raw.rfind(b"data_1") + 1] ^= 1                   # a link name inside the root group header
    open(path, "wb").write(bytes(flipped))
    info = {}
    load_jld2(path, info=info)
    assert info["checksum_failures"] == 1


def test_unsupported_features_fail_loudly(tmp_path):
    """A chunked layout (what JLD2 writes with compress=true) must be refused, never mis-read."""
    path = str(tmp_path / "t.jld2")
    save_jld2(path, {"data_1": np.ones((2, 2), dtype=np.complex128)}, commit_types=False)
    raw = bytearray(open(path, "rb").read())
    at = raw.find(bytes([0x08, 18, 0, 0, 3, 1]))             # layout message: type 8, size 18, flags 0, v3, contiguous
    assert at > 0
    raw[at + 5] = 2                                          # class 2 = chunked
    open(path, "wb").write(bytes(raw))
    with pytest.raises(QxbError) as e:
        load_jld2(path)
    assert e.value.code == -4 and "chunked" in str(e.value)


@pytest.mark.parametrize("name", sorted(os.path.basename(p) for p in glob.glob(os.path.join(ROOT, "workloads", "*.yml"))))
def test_params_reader_matches_pyyaml_on_workloads(name):
    path = os.path.join(ROOT, "workloads", name)
    want = yaml.safe_load(open(path))["output"]
    got = read_params(path)
    assert got["method"] == want["method"]
    if want["method"] == "List":
        assert got["bitstrings"] == [str(b) for b in want["params"]["bitstrings"]]
        assert got["num_samples"] == want["params"]["num_samples"]


def test_params_reader_schema(tmp_path):
    """Every method of output_params_dict (outputs.jl:54-77) as pyyaml / YAML.jl-style text."""
    p = str(tmp_path / "p.yml")
    for method in ("List", "Uniform", "Rejection"):
        q.generate_parameter_file(str(tmp_path / "p"), q.output_params_dict(5, 7, output_method=method, seed=11, M=0.25, fix_M=True))
        got = read_params(p)
        want = yaml.safe_load(open(p))["output"]["params"]
        assert got["method"] == method and got["num_samples"] == 7
        if method == "List":
            assert got["bitstrings"] == want["bitstrings"] and got["num_qubits"] == 5
        else:
            assert got["num_qubits"] == 5 and got["seed"] == 11
        if method == "Rejection":
            assert got["M"] == 0.25 and got["fix_M"] is True
        if method == "Uniform":                               # drawn with replacement, reproducible per seed
            assert len(got["bitstrings"]) == 7 and all(len(b) == 5 and set(b) <= {"0", "1"} for b in got["bitstrings"])
            assert got["bitstrings"] == read_params(p)["bitstrings"]
    # hand-written variants: comments, document marker, double quotes, flow sequence, null seed, nested indentation
    open(p, "w").write("# generated\n---\noutput:\n    method: \"List\"   # explicit\n    params:\n        seed: ~\n"
                       "        num_samples: 3\n        bitstrings: ['01', \"1+\", -0]\n")
    got = read_params(p)
    assert got["bitstrings"] == ["01", "1+", "-0"] and got["seed"] is None and got["num_qubits"] == 2
    open(p, "w").write("output:\n  method: List\n  params:\n    bitstrings:\n      - '012'\n")
    with pytest.raises(QxbError) as e:
        read_params(p)
    assert "0 1 + -" in str(e.value)
    open(p, "w").write("output:\n  method: Sampling\n  params:\n    num_samples: 1\n")
    with pytest.raises(QxbError) as e:
        read_params(p)
    assert "not supported" in str(e.value)                   # outputs.jl:74
    open(p, "w").write("output:\n  method: List\n   params: {}\n")
    with pytest.raises(QxbError):
        read_params(p)


def test_params_reader_matches_pyyaml_on_styles(tmp_path):
    """The same parameter dictionaries dumped by pyyaml in block style, with other indents, and with leaf collections
    in flow style (`bitstrings: ['01', ...]`, `params: {num_qubits: 5, ...}`): same result as yaml.safe_load."""
    import random
    rng = random.Random(5)
    p = str(tmp_path / "p.yml")
    for trial in range(60):
        nq = rng.randrange(1, 40)
        method = rng.choice(["List", "Uniform", "Rejection"])
        seed = rng.choice([None, 0, 7, 2 ** 40 + 3])
        if method == "List":
            bs = ["".join(rng.choice("01+-") for _ in range(nq)) for _ in range(rng.randrange(1, 12))]
            params = {"num_samples": len(bs), "bitstrings": bs}
        elif method == "Uniform":
            params = {"num_qubits": nq, "num_samples": rng.randrange(0, 20), "seed": seed}
        else:
            params = {"num_qubits": nq, "M": rng.choice([0.0001, 2.5, 1e-3]), "fix_M": rng.choice([True, False]), "seed": seed,
                      "num_samples": rng.randrange(0, 20)}
        doc = {"output": {"method": method, "params": params}}
        text = yaml.safe_dump(doc, sort_keys=rng.choice([True, False]), default_flow_style=rng.choice([False, None]),
                              indent=rng.choice([2, 4, 6]), width=rng.choice([40, 80, 10 ** 6]), explicit_start=rng.choice([True, False]))
        open(p, "w").write(text)
        want = yaml.safe_load(text)["output"]
        try:
            got = read_params(p)
        except QxbError as e:
            # the one construct of these dumps outside the subset: a flow sequence folded over several lines
            assert "same line" in str(e), (text, str(e))
            continue
        assert got["method"] == want["method"], text
        wp = want["params"]
        assert got["num_samples"] == wp["num_samples"], text
        if method == "List":
            assert got["bitstrings"] == [str(b) for b in wp["bitstrings"]] and got["num_qubits"] == nq, text
        else:
            assert got["num_qubits"] == nq and got["seed"] == wp["seed"], text
        if method == "Rejection":
            assert got["M"] == wp["M"] and got["fix_M"] == wp["fix_M"], text


def _triple(tmp_path):
    prefix = str(tmp_path / "rqc_3_3_8")
    q.generate_simulation_files(q.create_rqc_circuit(3, 3, 8, 42), prefix, 2, seed=42, time=0,
                                output_args=q.output_params_dict(9, 6, seed=5))
    return prefix


def test_graph_load_jld2(tmp_path):
    """.qx text + .jld2 data through the ABI alone: every data label of the program gets its tensor."""
    prefix = _triple(tmp_path)
    lib = load()
    g = C.c_void_p()
    assert lib.qxb_graph_create(C.byref(g), 1) == 0
    text = open(prefix + ".qx", "rb").read()
    assert lib.qxb_graph_parse_dsl(g, text, len(text)) == 0
    n = C.c_int()
    assert lib.qxb_graph_load_jld2(g, (prefix + ".jld2").encode(), C.byref(n)) == 0
    labels = {ln.split()[2] for ln in text.decode().splitlines() if ln.startswith("load ")}
    assert n.value == len(labels) == len(load_jld2(prefix + ".jld2"))
    need = lib.qxb_graph_describe(g, -1, None, 0)             # lowering needs extents only, data was accepted above
    assert need > 0
    lib.qxb_graph_destroy(g)
    # same tensors as the harness' in-memory cache
    tnc = q.convert_to_tnc(q.create_rqc_circuit(3, 3, 8, 42))
    bond_groups, plan, _ = q.contraction_scheme(tnc, 2, seed=42, time=0)
    cg = q.build_compute_graph(tnc, plan, bond_groups)
    back = load_jld2(prefix + ".jld2")
    assert set(back) == set(cg.tensors) and all(np.array_equal(back[k], cg.tensors[k]) for k in back)


def test_execute_files_host_part_and_no_cpu_fallback(tmp_path):
    """Without a GPU the native runner must get through parsing of all three files and then fail with
    QXB_ERR_CUDA at compile -- never fall back to a CPU path (tests/test_abi.py has the same for the ABI)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by tests/test_gpu_zfiles.py")
    prefix = _triple(tmp_path)
    lib = load()
    n = C.c_int64()
    rc = lib.qxb_execute_files((prefix + ".qx").encode(), None, None, (prefix + "_out.jld2").encode(), 1, -1, -1, 0, C.byref(n), None)
    assert rc == -3, lib.qxb_last_error()
    assert not os.path.exists(prefix + "_out.jld2")
    # a Rejection parameter file takes the same route (sampling loop = GPU calls)
    q.generate_parameter_file(prefix + "_rej", q.output_params_dict(9, 4, output_method="Rejection", M=8.0, fix_M=True, seed=3))
    rc = lib.qxb_execute_files((prefix + ".qx").encode(), None, (prefix + "_rej.yml").encode(), None, 1, -1, -1, 0, C.byref(n), None)
    assert rc == -3, lib.qxb_last_error()
    os.remove(prefix + ".jld2")
    rc = lib.qxb_execute_files((prefix + ".qx").encode(), None, None, None, 1, -1, -1, 0, None, None)
    assert rc == -1 and b".jld2" in lib.qxb_last_error()
    exe = os.path.join(ROOT, "qxtools.jl_b200", "bin", "qxrun")
    r = subprocess.run([exe, "-d", prefix + ".qx", "-o", prefix + "_out.jld2"], capture_output=True, text=True)
    assert r.returncode == 1 and "no CUDA device" in r.stderr
    r = subprocess.run([exe, "-o", "x.jld2"], capture_output=True, text=True)
    assert r.returncode == 2 and "--dsl" in r.stderr
