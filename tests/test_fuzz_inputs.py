"""Malformed inputs must come back as error codes, never as a crash or a mis-sized buffer: seeded byte / character
mutations of a valid `.jld2`, `.yml` and `.qx` through the library's native readers (host logic, no GPU)."""
import ctypes as C
import os
import random

import numpy as np

from cases import rqc_case
from qxb200._lib import QxbError, load
from qxb200.executor import Graph
from qxb200.jld2 import load_jld2, read_params, save_jld2

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_jld2_reader_survives_mutations(lib_built, tmp_path):
    rng = random.Random(11)
    arrays = {"data_1": np.arange(16).reshape(2, 2, 2, 2) * (1 + 1j), "data_2": np.array([1, 0], dtype=complex),
              "f": np.arange(5.0), "s": np.array([b"01", b"10"])}
    path = str(tmp_path / "m.jld2")
    outcomes = {"ok": 0, "err": 0}
    for commit_types in (True, False):
        save_jld2(path, arrays, commit_types=commit_types)
        raw = open(path, "rb").read()
        for _ in range(400):
            b = bytearray(raw)
            for _ in range(rng.choice([1, 1, 2, 4, 8])):
                pos = rng.randrange(512, len(b))
                b[pos] = rng.choice([0, 0xFF, b[pos] ^ (1 << rng.randrange(8)), rng.randrange(256)])
            if rng.random() < 0.1:
                b = b[:rng.randrange(512, len(b))]
            open(path, "wb").write(bytes(b))
            try:
                for v in load_jld2(path).values():
                    assert v.size <= len(b)                  # never more elements than the file has bytes
                outcomes["ok"] += 1
            except QxbError as e:
                assert e.code in (-1, -4)
                outcomes["err"] += 1
    assert outcomes["ok"] > 50 and outcomes["err"] > 50


def test_params_reader_survives_mutations(lib_built, tmp_path):
    rng = random.Random(12)
    text = open(os.path.join(ROOT, "workloads", "rqc_4x4_d12_c64_s16.yml")).read()
    path = str(tmp_path / "m.yml")
    errs = 0
    for _ in range(600):
        t = list(text)
        for _ in range(rng.choice([1, 2, 4])):
            t[rng.randrange(len(t))] = rng.choice(list(" \n:-'\"[]#\tx0~{"))
        if rng.random() < 0.2:
            t = t[:rng.randrange(len(t))]
        open(path, "w").write("".join(t))
        try:
            p = read_params(path)
            assert all(set(b) <= set("01+-") and len(b) == p["num_qubits"] for b in p["bitstrings"])
        except QxbError as e:
            assert e.code == -1
            errs += 1
    assert errs > 100


def test_dsl_parser_survives_mutations(lib_built):
    """Token-level mutations of a valid program: parse + analysis + lowering either succeed or return ERR_ARG / ERR_UNSUPP."""
    rng = random.Random(13)
    txt, data, _ = rqc_case(3, 3, 8, 2)
    lines = txt.strip().splitlines()
    lib = load()
    seen = {"ok": 0, "err": 0}
    for _ in range(300):
        ls = list(lines)
        for _ in range(rng.choice([1, 1, 2])):
            i = rng.randrange(1, len(ls))
            tok = ls[i].split()
            kind = rng.randrange(5)
            if kind == 0 and tok:
                tok[rng.randrange(len(tok))] = rng.choice(["0", "-1", "99", "t1", "v1", "1,2,3", "", "ncon", "1,,2", "9" * 25])
            elif kind == 1:
                tok = tok[:rng.randrange(len(tok) + 1)]
            elif kind == 2:
                ls.insert(i, ls[rng.randrange(1, len(ls))])
                continue
            elif kind == 3:
                del ls[i]
                continue
            else:
                tok.append(rng.choice(["1", "x", "2,2"]))
            ls[i] = " ".join(tok)
        prog = ("\n".join(ls) + "\n").encode()
        g = Graph()
        rc = lib.qxb_graph_parse_dsl(g._h, prog, len(prog))
        if rc == 0:
            try:
                g.set_data(data)
                rc = lib.qxb_graph_describe(g._h, -1, None, 0)
            except QxbError as e:
                rc = e.code
        assert rc >= 0 or rc in (-1, -2, -4), (rc, lib.qxb_last_error())
        seen["ok" if rc >= 0 else "err"] += 1
    assert seen["err"] > 50
