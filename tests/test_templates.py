"""contract_kernel's launch templates replayed on the CPU (tests/template_emulator.py): for every
lowered op of the golden programs and of random tensor-network programs, the addresses the kernel
computes -- thread bits, register tile, tile index, K chunks -- must enumerate every C element exactly
once and read exactly the A / B elements the lowered op defines.  Run for the default split and for
the QXB_MIN_LOB variants (register tiles for nodes with few elements per bitstring), which changes
only this composition, never the kernel."""
import os

import numpy as np
import pytest

from cases import kat0, random_program, rqc_case
from qxb200.executor import Graph
from template_emulator import kernel_addresses, loads_per_output, reference_addresses, templates


def check_graph(g, n_free=-1):
    d = g.describe(n_free)
    tm = templates(g, n_free)
    assert len(tm) == len(d["ops"])
    stats = []
    for p, op in zip(tm, d["ops"]):
        assert (p.nC, p.nK) == (op["nC"], op["nK"])
        c, a, b = kernel_addresses(p)
        assert np.array_equal(np.sort(c), np.arange(1 << p.nC)), op["name"]          # every output exactly once
        ra, rb = reference_addresses(op, c)
        # same multiset of (A, B) address pairs per output (the kernel may walk k in another order)
        key = lambda x, y: np.sort(x * (1 << 40) + y, axis=1)
        assert np.array_equal(key(a, b), key(ra, rb)), op["name"]
        # warps store contiguous runs: lanes 0..31 of a tile element differ only in C bits 0..4
        if p.lob >= 5:
            lane_span = c.reshape(1 << p.hb, 1 << p.lob, -1)[:, :32, :]
            assert np.array_equal(np.sort(lane_span[0, :, 0] - lane_span[0, 0, 0]), np.arange(32)), op["name"]
        stats.append((op["name"], p.nC, p.nK, p.lob, p.ma, p.nb, p.kc, loads_per_output(p)))
    return stats


CASES = {"kat0": None, "rqc_3x3_d8_s2": (3, 3, 8, 2), "rqc_4x4_d12_s4": (4, 4, 12, 4), "rqc_4x5_d14_s5": (4, 5, 14, 5),
         "random_11": 11, "random_12": 12, "random_13": 13}


def load_case(name, dtype="c64", replan=0):
    if name == "kat0":
        txt, data = kat0()
    elif name.startswith("random"):                       # extents 2..3 (zero-padded modes), hyper-indices, slices
        txt, data, _ = random_program(CASES[name])
    else:
        txt, data, _ = rqc_case(*CASES[name])
    return Graph.from_dsl(txt, data, dtype, replan=replan, replan_n_amp=4096)


@pytest.mark.parametrize("min_lob", [None, "7", "6", "5"])
@pytest.mark.parametrize("name", CASES)
def test_templates_match_lowered_ops(lib_built, monkeypatch, name, min_lob):
    if min_lob is None:
        monkeypatch.delenv("QXB_MIN_LOB", raising=False)
    else:
        monkeypatch.setenv("QXB_MIN_LOB", min_lob)
    for dtype in ("c64", "c32"):
        g = load_case(name, dtype)
        k = len(g.slice_dims)
        for n_free in sorted({-1, 0, k // 2}):
            check_graph(g, n_free)


@pytest.mark.parametrize("min_lob", [None, "5"])
def test_templates_of_replanned_programs(lib_built, monkeypatch, min_lob):
    """The tree-searched programs are the ones with many small nodes and long K (K chunks, multi-segment k maps)."""
    if min_lob is None:
        monkeypatch.delenv("QXB_MIN_LOB", raising=False)
    else:
        monkeypatch.setenv("QXB_MIN_LOB", min_lob)
    long_k = False
    for name in ("rqc_3x3_d8_s2", "rqc_4x4_d12_s4", "rqc_4x5_d14_s5"):
        stats = check_graph(load_case(name, "c64", replan=8))
        long_k |= any(s[2] >= 3 for s in stats)
    assert long_k                                            # some op has K >= 8


def test_min_lob_gives_small_nodes_a_register_tile(lib_built, monkeypatch):
    """What the knob is for: ops with <= 2^8 elements per bitstring get M/N register-tile bits (fewer operand loads per
    output), ops that already had 8 thread bits left keep their split."""
    g = load_case("rqc_4x5_d14_s5", "c64", replan=8)
    monkeypatch.delenv("QXB_MIN_LOB", raising=False)
    base = {s[0]: s for s in check_graph(g)}
    monkeypatch.setenv("QXB_MIN_LOB", "5")
    tiled = {s[0]: s for s in check_graph(g)}
    assert all(s[3] == min(8, s[1]) or s[4] + s[5] > 0 for s in base.values())
    fewer = [n for n in base if tiled[n][7] < base[n][7]]
    more = [n for n in base if tiled[n][7] > base[n][7]]
    assert fewer and not more
    assert all(tiled[n][3] >= 5 or tiled[n][1] < 5 + tiled[n][4] + tiled[n][5] for n in tiled)


@pytest.mark.parametrize("min_lob", [None, "6"])
@pytest.mark.parametrize("workload,replan", [("rqc_6x6_d16_c32_s64", 0), ("rqc_7x7_d20_c64_s4096", 16)])
def test_templates_of_bench_workloads(lib_built, monkeypatch, workload, replan, min_lob):
    """The committed benchmark triples have the nodes with > 2^8 elements per bitstring, i.e. the 2^ma x 2^nb
    register tiles and multi-chunk K walks of the default split (the headline step's dominant contractions)."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    if min_lob is None:
        monkeypatch.delenv("QXB_MIN_LOB", raising=False)
    else:
        monkeypatch.setenv("QXB_MIN_LOB", min_lob)
    txt, data, w = bench.build_workload(workload)
    g = Graph.from_dsl(txt, data, w["dtype"], replan=replan, replan_n_amp=32768)
    stats = check_graph(g)
    tiles = {(s[4], s[5]) for s in stats}
    assert len(tiles) >= 4 and any(m and n for m, n in tiles)
    assert any(s[6] < s[2] for s in stats)                   # K walked in more than one register chunk somewhere
