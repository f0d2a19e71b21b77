"""The C-ABI library loads on a CPU-only box, exports every symbol include/qxb200.h
declares, and its host-side logic (parser, analysis, slice enumeration, lowering
report, error behaviour) works without a GPU.  No compute calls here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from qxb200 import _lib
from qxb200._lib import QxbError
from qxb200.executor import Graph
from oracle import qx_oracle as orc
from cases import kat0, rqc_case

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_exported(lib_built):
    hdr = open(os.path.join(ROOT, "include", "qxb200.h")).read()
    declared = set(re.findall(r"\b(qxb_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(_lib.SYMBOLS)
    for s in declared:
        assert hasattr(lib_built, s), s
    assert lib_built.qxb_version() >= 100


def test_parse_and_queries(lib_built):
    txt, data = kat0()
    g = Graph.from_dsl(txt, data)
    assert g.n_outputs == 2 and g.slice_dims == [2, 2] and g.n_slices == 4
    cmds = orc.parse_dsl(txt)
    dims = orc.slice_dims(cmds)
    for s in range(4):                       # bit-exact slice bookkeeping vs the oracle
        assert g.slice_values(s) == list(orc.slice_values(s, dims).values())
    with pytest.raises(QxbError):
        g.slice_values(4)


def test_builder_calls_equal_parser(lib_built):
    import qxb200 as q
    tnc = q.convert_to_tnc(q.create_rqc_circuit(3, 3, 8, 3))
    bg, plan, _ = q.contraction_scheme(tnc, 2, time=0)
    cg = q.build_compute_graph(tnc, plan, bg)
    a = Graph.from_compute_graph(cg).describe()
    b = Graph.from_dsl(cg.dsl(), cg.tensors).describe()
    assert a == b


@pytest.mark.parametrize("bad,code", [
    ("load t1 data_1 2\n", -1),                                   # no version line
    ("# version: 0.4.0\nfoo t1\n", -1),                           # unknown instruction
    ("# version: 0.4.0\nload t1 data_1 2\nncon t2 1 t1 1 t9 1\nsave output t2\n", -1),   # undefined symbol
    ("# version: 0.4.0\nload t1 data_1 2\nload t2 data_1 2\nncon t3 0 t1 1,2 t2 1\nsave output t3\n", -1),  # rank
    ("# version: 0.4.0\nload t1 data_1 2\nload t2 data_1 2\nncon t3 0 t1 1 t2 1\n", -1),  # no save
    ("# version: 0.4.0\nload t1 data_1 2\nview t1_s t1 v1 2 2\n", -1),                    # mode out of range
    ("# version: 0.4.0\nload t1 data_1 2\nview t1_s t1 v1 1 3\n", -1),                    # bond dim mismatch
    ("# version: 0.4.0\nload t1 data_1 2,2\nload t2 data_1 2\nncon t3 1 t1 1,2 t2 2\nview x t3 v1 1 2\nsave output t3\n", -4),
])
def test_malformed_programs(lib_built, bad, code):
    g = Graph()
    b = bad.encode()
    rc = g._lib.qxb_graph_parse_dsl(g._h, b, len(b))
    if rc == 0:
        n = C.c_int64()
        rc = g._lib.qxb_graph_num_slices(g._h, C.byref(n))      # forces analysis
    assert rc == code, g._lib.qxb_last_error()
    assert len(g._lib.qxb_last_error()) > 0


def test_tensor_valued_root(lib_built):
    """A non-scalar ``save`` (open network) is accepted: its shape is reported in Julia order and its modes are
    located in the lowered root; slicing one of its modes is refused (tests/test_open_network.py has the values)."""
    g = Graph.from_dsl("# version: 0.4.0\nload a d 2,2\nload b e 2,3\nncon c 1,3 a 1,2 b 2,3\nsave output c\n",
                       {"d": np.eye(2), "e": np.ones((2, 3))})
    assert g.root_dims == [2, 3] and g.n_outputs == 0
    d = g.describe()
    assert [m[0] for m in d["root_modes"]] == [2, 3] and [m[1] for m in d["root_modes"]] == [1, 2]
    assert Graph.from_dsl("# version: 0.4.0\nload a d 2\nload b d 2\nncon c 0 a 1 b 1\nsave output c\n",
                          {"d": np.ones(2)}).root_dims == []
    g = Graph.from_dsl("# version: 0.4.0\nload a d 2,2\nview a_s a v1 1 2\nload b d 2,2\nncon c 1,3 a_s 1,2 b 2,3\nsave output c\n",
                       {"d": np.eye(2)})
    with pytest.raises(QxbError) as e:
        g.describe()
    assert e.value.code == -4 and "open indices cannot be slice bonds" in str(e.value)


def test_compute_fails_loudly_without_gpu(lib_built):
    """No CPU fallback: on a box without CUDA the compute entry points must error."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    txt, data = kat0()
    g = Graph.from_dsl(txt, data)
    with pytest.raises(QxbError) as e:
        g.compile()
    assert e.value.code == -3
    with pytest.raises(QxbError):
        g.amplitudes(["00"])


def test_describe_split_is_consistent(lib_built):
    txt, data, _ = rqc_case(4, 4, 12, 4)
    g = Graph.from_dsl(txt, data)
    for n_free in (0, 2, 4):
        d = g.describe(n_free)
        assert d["n_free"] == n_free and d["n_slice_vars"] == 4
        for op in d["ops"]:
            assert op["batch_bits"] + op["m_bits"] + op["n_bits"] == op["nC"]
            assert sum(s[2] for s in op["segA"]) + sum(s[2] for s in op["segKA"]) <= op["a_bits"]
        # fully fixed: nothing is batched over slices, so no op has more bits than the unsliced tensors
        if n_free == 0:
            assert max(op["nC"] for op in d["ops"]) <= 8


def _build_consumer(tmp_path):
    import subprocess
    exe = str(tmp_path / "abi_consumer")
    pkg = os.path.join(ROOT, "qxtools.jl_b200")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "abi_consumer.c"), "-o", exe, "-L", pkg, "-lqxb200", "-lm",
                    f"-Wl,-rpath,{pkg}"], check=True)
    return exe


def test_plain_c_consumer_host_side(lib_built, tmp_path):
    """The header is valid C99 and the ABI is usable without Python (what a Julia ccall sees)."""
    import subprocess
    exe = _build_consumer(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "abi consumer ok" in r.stdout


@pytest.mark.gpu
def test_plain_c_consumer_on_gpu(gpu, tmp_path):
    import subprocess
    exe = _build_consumer(tmp_path)
    r = subprocess.run([exe, "gpu"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "gpu amplitudes ok" in r.stdout


def _header_option_fields():
    """(type, name) of every field of `qxb_options` in include/qxb200.h, in order."""
    import re
    txt = open(os.path.join(ROOT, "include", "qxb200.h")).read()
    body = txt[txt.index("typedef struct qxb_options {"):txt.index("} qxb_options;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    return re.findall(r"\b(int64_t|int32_t)\s+(\w+)\s*;", body)


def test_options_mirrors_match_the_header():
    """The ctypes mirror (qxb200/_lib.py) and the Julia mirror (julia/QXB200.jl) of `qxb_options` list the header's fields
    in the header's order with the header's widths; sizeof agrees with a C compile of the header."""
    import ctypes as C
    import re
    import subprocess
    import tempfile
    from qxb200 import _lib
    fields = _header_option_fields()
    assert len(fields) >= 23
    want = [(n, C.c_int64 if t == "int64_t" else C.c_int32) for t, n in fields]
    assert [(n, t) for n, t in _lib.Options._fields_] == want
    jl = open(os.path.join(ROOT, "julia", "QXB200.jl")).read()
    body = jl[jl.index("struct Options"):]
    body = body[:body.index("\nend")]
    jfields = re.findall(r"^\s+(\w+)::(Int64|Int32)", body, flags=re.M)
    assert jfields == [(n, "Int64" if t == "int64_t" else "Int32") for t, n in fields]
    d = tempfile.mkdtemp()
    src = os.path.join(d, "sz.c")
    with open(src, "w") as f:
        f.write('#include <stdio.h>\n#include "%s"\nint main(void) { printf("%%zu", sizeof(qxb_options)); return 0; }\n'
                % os.path.join(ROOT, "include", "qxb200.h"))
    subprocess.check_call(["gcc", src, "-o", os.path.join(d, "sz")])
    assert int(subprocess.check_output([os.path.join(d, "sz")]).decode()) == C.sizeof(_lib.Options)
