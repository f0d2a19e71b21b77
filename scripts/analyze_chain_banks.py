"""Shared-memory bank-conflict model of the fused-chain row program of the headline plan, from the exact device tables
(qxb_debug_rowprog phase 3): per op, wavefronts per 128-bit load of A / B and per store of C against the 4 a conflict-free
access needs (ComplexF64: 16-byte elements, a quarter-warp covers the 32 banks once).  CPU only."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import bench
import rowprog_emulator as emu
from qxb200.executor import Graph

wl = os.environ.get("PROBE_WORKLOAD", "rqc_7x7_d20_c64_s4096")
txt, data, w = bench.build_workload(wl)
g = Graph.from_dsl(txt, data, w["dtype"], replan=128, replan_n_amp=131072)
es = 16 if w["dtype"] == "c64" else 8
fm = int(os.environ.get("FREE_MASK", "-1"))
d = g.describe(0)
if fm < 0:
    fm = 0
rp = emu.dump(g, fm, 3)
assert rp is not None, g._lib.qxb_last_error().decode()
print("fused ops:", rp.fused_names, "levels", rp.n_levels, "descs", len(rp.descs), "arena elems", rp.arena_elems)


def wavefronts(offs, act):
    """offs: element offsets of the 32 lanes; 128-bit (c64) accesses go a quarter-warp at a time, 64-bit a half-warp."""
    per = 8 if es == 16 else 16
    tot = 0
    for q in range(0, 32, per):
        o = [int(offs[l]) for l in range(q, q + per) if act[l]]
        if not o:
            continue
        uniq = set(o)
        groups = {}
        for e in uniq:
            groups.setdefault(e % per, set()).add(e)
        tot += max(len(v) for v in groups.values())
    return tot


tot_w = tot_ideal = 0
for lv in range(rp.n_levels):
    seen = {}
    for s in range(rp.level_start[lv], rp.level_start[lv + 1]):
        di = rp.slots[s]
        if di == emu.K_NULL:
            continue
        op = rp.descs[di]
        h = op.hot
        if h.kind in (0, 254, 255):
            continue
        j = rp.desc_op[di]
        act = np.array(op.lC[:]) != emu.K_NULL
        TM, TN = 1 << h.ma, 1 << h.nb
        wa = wavefronts(op.lA[:], act); wb = wavefronts(op.lB[:], act); wc = wavefronts(op.lC[:], act)
        ideal = 32 * es // 128
        nk = 1 << h.nK
        loads = nk * (TM * wa + TN * wb) + TM * TN * wc
        ideal_l = nk * (TM + TN) * ideal + TM * TN * ideal
        # broadcast loads: fewer distinct addresses than lanes need fewer wavefronts than "ideal"
        tot_w += loads; tot_ideal += ideal_l
        key = j
        if key not in seen:
            seen[key] = [0, wa, wb, wc, h.ma, h.nb, h.kc, h.nK]
        seen[key][0] += 1
    for j, v in seen.items():
        print(f"level {lv} op {j} units {v[0]} tile {1 << v[4]}x{1 << v[5]} kc {v[6]} nK {v[7]}: wavefronts per access A {v[1]} B {v[2]} C {v[3]} (full-width conflict-free = {32 * es // 128})")
print("total wavefronts per row (model)", tot_w, "if every access took", 32 * es // 128, ":", tot_ideal)
