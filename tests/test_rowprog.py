"""Row programs (csrc/qxb_rowprog.h: a whole phase of the tree as one persistent kernel, intermediates in shared
memory) replayed on the CPU: the descriptors the library builds -- levels, warp units, aligned arena plan, XOR-combined
tables, K-splitting lanes -- must reproduce the oracle.  The GPU parity tests then only have to show that the kernel
follows the same loops (tests/test_gpu_rowprog.py)."""
import numpy as np
import pytest

from qxb200.executor import Graph, bits_from_strings
from oracle import qx_oracle as orc
import rowprog_emulator as rpe
import lowered_emulator as le
from cases import kat0, rqc_case, circuit_case, random_program
import qxb200 as q


def test_kat0_all_ranges(lib_built):
    txt, data = kat0()
    g = Graph.from_dsl(txt, data)
    bs = ["00", "11", "01", "10"]
    bits = bits_from_strings(bs, 2)
    cmds = orc.parse_dsl(txt)
    for b in range(4):
        for e in range(b + 1, 5):
            ref = orc.amplitudes(cmds, data, bs, slice_begin=b, slice_end=e)
            assert np.allclose(rpe.amplitudes(g, data, bits, b, e), ref, atol=1e-15)


@pytest.mark.parametrize("shape", [(3, 3, 8, 2), (3, 4, 10, 3), (4, 4, 12, 4)])
def test_rqc_rowprog_matches_oracle(lib_built, shape):
    r, c, d, ns = shape
    txt, data, bs = rqc_case(r, c, d, ns, n_amp=4)
    g = Graph.from_dsl(txt, data)
    bits = bits_from_strings(bs + ["+" * (r * c), "-" * (r * c)], r * c)
    cmds = orc.parse_dsl(txt)
    ref = orc.amplitudes(cmds, data, bs + ["+" * (r * c), "-" * (r * c)])
    assert np.allclose(rpe.amplitudes(g, data, bits), ref, atol=1e-14)
    S = g.n_slices
    for (b, e) in [(1, S - 1), (3, 4)]:
        ref = orc.amplitudes(cmds, data, bs, slice_begin=b, slice_end=e)
        assert np.allclose(rpe.amplitudes(g, data, bits[:len(bs)], b, e), ref, atol=1e-14)


def test_replanned_rqc_rowprog(lib_built):
    """The re-planned (batch-aware) tree is what the bench runs: bigger nodes, register tiles, several units per op."""
    txt, data, bs = rqc_case(4, 4, 12, 4, n_amp=3)
    g = Graph.from_dsl(txt, data, replan=16, replan_n_amp=1024)
    bits = bits_from_strings(bs, 16)
    ref = orc.amplitudes(orc.parse_dsl(txt), data, bs)
    assert np.allclose(rpe.amplitudes(g, data, bits), ref, atol=1e-14)
    rp = rpe.dump(g, (1 << len(g.slice_dims)) - 1, 2)
    assert rp.n_levels < len(rp.lop)                      # levels really group independent ops


@pytest.mark.parametrize("dmma", [False, True])
@pytest.mark.parametrize("workload", ["rqc_7x7_d20_c64_s4096", "rqc_6x6_d16_c32_s64"])
def test_bench_workloads_rowprog(lib_built, workload, dmma, monkeypatch):
    """The plans the bench runs (every slice variable batched: nodes of up to 2^11 elements per bitstring, 2 x 2
    register tiles, several warp units per node, K chunks) against the emulator of the lowered program -- itself
    pinned to the oracle by tests/test_lowering.py -- and, on a short slice range, against the oracle directly."""
    import bench
    if dmma:                                              # 8 x 8 tiles on the FP64 tensor pipe (mma.sync.m8n8k4 fragments)
        monkeypatch.setenv("QXB_ROW_DMMA", "1")
    txt, data, w = bench.build_workload(workload)
    g = Graph.from_dsl(txt, data, "c64", replan=8, replan_n_amp=131072)
    nq = w["rows"] * w["cols"]
    bits = bench.synth_bits(2, nq)
    bits[1, :3] = (2, 3, 2)                              # '+' / '-' outputs as well
    got = rpe.amplitudes(g, data, bits)
    ref = le.amplitudes(g, data, bits)
    assert np.max(np.abs(got - ref)) <= 1e-12 * np.max(np.abs(ref))
    k = len(g.slice_dims)
    rp = rpe.dump(g, (1 << k) - 1, 2)
    kinds = {op.hot.kind for op in rp.descs}
    assert 0 in kinds and max(kinds) > 1                  # both the K-splitting path and register tiles are exercised
    if workload.startswith("rqc_7x7"):
        assert (254 in kinds) == dmma
        assert max(op.hot.ma + op.hot.nb for op in rp.descs) >= 3
    # a 3-slice range (blocks with fixed variables) against the oracle
    bs = ["".join("01"[b] for b in row) for row in bench.synth_bits(1, nq)]
    ref = orc.amplitudes(orc.parse_dsl(g.text), data, bs, slice_begin=5, slice_end=8)
    got = rpe.amplitudes(g, data, bits_from_strings(bs, nq), 5, 8)
    assert np.max(np.abs(got - ref)) <= 1e-12 * max(np.max(np.abs(ref)), 2.0 ** (-nq / 2))


@pytest.mark.parametrize("seed", range(6))
def test_random_programs_rowprog(lib_built, seed):
    """Hyper-edges, outer products, non-power-of-two extents, views on outputs (fixed offsets into arena leaves)."""
    txt, data, bs = random_program(seed)
    g = Graph.from_dsl(txt, data)
    n_out = g.n_outputs
    bits = bits_from_strings(bs, n_out)
    cmds = orc.parse_dsl(txt)
    ref = orc.amplitudes(cmds, data, bs)
    try:
        got = rpe.amplitudes(g, data, bits)
    except rpe.Unavailable as e:                          # a row too large for shared memory: the per-op path runs it
        assert "exceeds the budget" in str(e) or "2^16" in str(e)
        pytest.skip(f"no row program: {e}")
    assert np.allclose(got, ref, atol=1e-12 * max(1.0, float(np.max(np.abs(ref)))))
    S = g.n_slices
    if S > 2:
        ref = orc.amplitudes(cmds, data, bs, slice_begin=1, slice_end=S - 1)
        got = rpe.amplitudes(g, data, bits, 1, S - 1)
        assert np.allclose(got, ref, atol=1e-12 * max(1.0, float(np.max(np.abs(ref)))))


def test_ghz_and_qft_rowprog(lib_built):
    for circ, ns in ((q.create_ghz_circuit(5), 0), (q.create_qft_circuit(6), 0), (q.create_ghz_circuit(4), 2)):
        txt, data, bs = circuit_case(circ, n_slice=ns, n_amp=4)
        g = Graph.from_dsl(txt, data)
        bits = bits_from_strings(bs, circ.num_qubits)
        ref = orc.amplitudes(orc.parse_dsl(txt), data, bs)
        assert np.allclose(rpe.amplitudes(g, data, bits), ref, atol=1e-14)


def test_arena_plan_stages_shared_operands_and_fits(lib_built):
    """Operands shared by all rows are copied into the arena one level ahead of their first reader, and the headline
    workload's row fits a B200 SM's shared memory."""
    import bench
    txt, data, w = bench.build_workload("rqc_7x7_d20_c64_s4096")
    g = Graph.from_dsl(txt, data, "c64", replan=8, replan_n_amp=131072)
    k = len(g.slice_dims)
    rp = rpe.dump(g, (1 << k) - 1, 2)
    assert rp is not None
    d = g.describe()
    # no two tensors that are live in the same level overlap: checked dynamically by the emulator (NaN-filled arena,
    # per-level race check); here: every staged operand gets a copy unit one level before its first reader
    staged = [j for j in range(len(rp.lop)) if rp.lop[j] < 0]
    assert staged, "the bench plan has operands shared by all rows: they must be staged"
    lvl_of_desc = {}
    for lv in range(rp.n_levels):
        for sl in range(rp.level_start[lv], rp.level_start[lv + 1]):
            if rp.slots[sl] != 0xFFFF:
                lvl_of_desc[rp.slots[sl]] = lv
    first_use = {}
    for di, op in enumerate(rp.descs):
        j = rp.desc_op[di]
        if rp.lop[j] < 0:
            continue
        for t in (rp.ref_a[j], rp.ref_b[j]):
            first_use[t] = min(first_use.get(t, 1 << 30), lvl_of_desc[di])
        assert op.hot.gen == 0                            # nothing on the chunk path reads global memory at use time
    for di, op in enumerate(rp.descs):
        j = rp.desc_op[di]
        if rp.lop[j] < 0:
            assert lvl_of_desc[di] == first_use[rp.ref_a[j]] - 1
    assert rp.arena_elems * 16 <= 226 * 1024              # fits one SM's shared memory (<= 113 KB: two CTAs per SM)


def _fused_plan(graph, free_mask):
    import ctypes as C
    lib = graph._lib
    need = lib.qxb_debug_fused_plan(graph._h, free_mask, None, 0)
    assert need > 0, lib.qxb_last_error().decode()
    buf = C.create_string_buffer(need)
    assert lib.qxb_debug_fused_plan(graph._h, free_mask, buf, need) == need
    fused, ops, tensors = (-1, -1), [], {}
    for ln in buf.value.decode().splitlines():
        f = ln.split()
        if f[0] == "fused":
            fused = (int(f[1]), int(f[2]))
        elif f[0] == "op":
            ops.append((int(f[1]), f[2], int(f[3]), int(f[4]), int(f[5])))
        else:
            tensors[int(f[1])] = dict(off=int(f[2]), size=int(f[3]), amp=int(f[4]), leaf=int(f[5]))
    return fused, ops, tensors


def test_fused_launch_results_never_alias_its_inputs():
    """One fused launch works through all rows of the batch: a result it writes to HBM must not share arena space with
    ANY tensor the launch reads from HBM (another CTA may still have to read that row).  plan_memory therefore releases no
    operand of the fused ops before the last one.  Checked on the headline plan (where the unfixed planner put the chain's
    result on an operand of an earlier chain op once side branches changed the allocation order) and on a 6x6 plan."""
    import bench
    for wl, fm, replan in (("rqc_7x7_d20_c64_s4096", 4095, 32), ("rqc_6x6_d16_c32_s64", 63, 16)):
        txt, data, w = bench.build_workload(wl)
        g = Graph.from_dsl(txt, data, w["dtype"], replan=replan, replan_n_amp=131072)
        for chain_side in (0, 1, 2):                     # a pure path, and with side branches taken in
            g.configure(chain_side=chain_side)
            (first, last), ops, T = _fused_plan(g, fm)
            if first < 0:
                continue
            inside = [o for o in ops if first <= o[0] <= last]
            assert len(inside) == last - first + 1 >= 2
            produced = {o[4] for o in inside}
            read_later = {t for o in ops if o[0] > last for t in (o[2], o[3])}
            inputs = {t for o in inside for t in (o[2], o[3]) if t not in produced and T[t]["off"] >= 0 and T[t]["amp"]}
            results = {o[4] for o in inside if o[4] in read_later or o[0] == last}
            assert inputs and results
            for r in results:
                for i in inputs:
                    a, b = T[r], T[i]
                    assert a["off"] + a["size"] <= b["off"] or b["off"] + b["size"] <= a["off"], (wl, chain_side, r, i, a, b)


@pytest.mark.parametrize("knobs", [dict(), dict(row_bank_opt=False), dict(chain_side=1)])
def test_fused_chain_program_on_the_cpu(knobs):
    """The fused chain of the headline plan, replayed from its exact device tables (levels with a barrier in between,
    staged copies of the per-row inputs, lane tables with the bank-aware lane bits, re-laid-out intermediates in a
    NaN-filled arena, results written to global rows) against the lowered program executed op by op."""
    import bench
    txt, data, w = bench.build_workload("rqc_7x7_d20_c64_s4096")      # the headline plan: a chain of six 2^9 - 2^11-element nodes
    g = Graph.from_dsl(txt, data, "c64", replan=32, replan_n_amp=131072)
    g.configure(**knobs)
    bits = bench.synth_bits(2, 49)
    n = bits.shape[0]
    fm = (1 << 12) - 1                                                 # all twelve slice variables batched, as the plan runs
    desc = g.describe(fm)
    rp = rpe.dump(g, fm, 3)
    if rp is None:
        pytest.skip("no chain in this plan: " + g._lib.qxb_last_error().decode())
    assert len(rp.fused_names) >= 2
    # the reference keeps EVERY tensor (the memory plan recycles arena space after a tensor's last reader; here the
    # chain's inputs must still be there afterwards): disjoint offsets per phase
    nxt = {"const": 0, "block": 0, "chunk": 0}
    for t in desc["tensors"]:
        if t["leaf"] and not t["output_leaf"]:
            continue
        t["offset"] = nxt[t["phase"]]
        nxt[t["phase"]] += max(2, 1 << t["span_bits"])
    desc["arena_elems"] = {"const": nxt["const"], "block": nxt["block"], "chunk_per_amp": nxt["chunk"]}
    state = {}
    le.run_block(desc, data, bits, np.zeros(64, dtype=np.int64), state=state)
    T, locate = desc["tensors"], state["locate"]
    outs = {}                                            # tensors the chain writes to global memory: fresh buffers

    for u in range(n):
        def resolve(ti, base=False):
            t = {**T[ti], "fixed": []} if base else T[ti]
            if ti in outs:
                return rpe._Mem(outs[ti], u << T[ti]["span_bits"])
            buf, b0, sU = locate(t)
            return rpe._Mem(buf, b0 + u * sU)
        for j in range(len(rp.ref_c)):
            if rp.lop[j] >= 0 and not rp.in_arena_c[j] and rp.ref_c[j] not in outs:
                outs[rp.ref_c[j]] = np.full(n << T[rp.ref_c[j]]["span_bits"], np.nan + 0j, dtype=np.complex128)
        arena = np.full(max(rp.arena_elems, 1), np.nan + 0j, dtype=np.complex128)
        rpe.run_program(rp, resolve, arena, np.complex128)
    assert outs
    for ti, got in outs.items():
        buf, b0, sU = locate(T[ti])
        want = np.concatenate([buf[b0 + u * sU: b0 + u * sU + (1 << T[ti]["span_bits"])] for u in range(n)])
        assert np.allclose(got, want, rtol=1e-12, atol=1e-14), ti
