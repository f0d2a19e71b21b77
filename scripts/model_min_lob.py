"""CPU-side model of the headline step (RQC 7x7 d20 c64, tree-searched plan, 131072 bitstrings) under QXB_MIN_LOB.

No GPU needed: re-plans the committed workload exactly as bench.py does, reads the launch templates of every
contraction (qxb_debug_templates), counts the operand loads contract_kernel issues per output, and models each op as
    t = max(HBM bytes / BW_hbm, L1 bytes / BW_l1)
with L1 bytes = loads x element size.  BW_l1 is CALIBRATED on the measured per-op times of the same plan
(profiles/r1p_ops.md, CUDA events on B200), then the model is re-evaluated with the register-tile split that
QXB_MIN_LOB = 7 / 6 / 5 produces.  Output: profiles/r1q_model_min_lob.md."""
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np            # noqa: E402
import bench                  # noqa: E402
from qxb200.executor import Graph   # noqa: E402
from template_emulator import templates   # noqa: E402

N_AMP = 131072
os.environ["QXB_PLAN_L1_BW"] = "0"     # the plan that was MEASURED (r1p): single-rate planner model; see scripts/model_plan_rate.py
ES = 16
BW_HBM = 6.0e12               # what the streaming nodes reach (93 % of the measured 6.55 TB/s copy rate)

measured = {}
for ln in open(os.path.join(ROOT, "profiles", "r1p_ops.md")):
    m = re.match(r"\| (R\d+) \| (\d+) \| (\d+) \| (\d+)/(\d+)/(\d+) \| ([\d.]+) \|", ln)
    if m:
        measured[m.group(1)] = dict(nC=int(m.group(2)), nK=int(m.group(3)), split="/".join(m.group(4, 5, 6)), ms=float(m.group(7)))

txt, data, w = bench.build_workload("rqc_7x7_d20_c64_s4096")


def plan(min_lob):
    if min_lob is None:
        os.environ.pop("QXB_MIN_LOB", None)
    else:
        os.environ["QXB_MIN_LOB"] = str(min_lob)
    g = Graph.from_dsl(txt, data, "c64", replan=128, replan_n_amp=N_AMP)
    d = g.describe()
    rows = []
    for p, op in zip(templates(g), d["ops"]):
        if op["phase"] != "chunk":
            continue
        u = N_AMP
        out = u * 2.0 ** p.nC
        hbm = ES * ((2.0 ** op["a_bits"]) * (u if op["a_amp"] else 1) + (2.0 ** op["b_bits"]) * (u if op["b_amp"] else 1) + out)
        loads = out * ((1 << p.ma) + (1 << p.nb)) * 2.0 ** p.nK / (1 << (p.ma + p.nb))
        rows.append(dict(name=op["name"], nC=p.nC, nK=p.nK, lob=p.lob, ma=p.ma, nb=p.nb, kc=p.kc, hbm=hbm, l1=loads * ES,
                         split=f'{op["batch_bits"]}/{op["m_bits"]}/{op["n_bits"]}'))
    return rows


base = plan(None)
# the tree search has been extended since the r1p capture, so node names moved: match measured ops to today's ops by
# shape signature (nC, nK, batch/M/N bits), largest first; unmatched ops on either side are left out of the calibration
common, used = [], set()
for mname, mm in sorted(measured.items(), key=lambda kv: -kv[1]["ms"]):
    for r in base:
        if r["name"] not in used and (r["nC"], r["nK"], r["split"]) == (mm["nC"], mm["nK"], mm["split"]):
            used.add(r["name"])
            common.append(r)
            measured[r["name"]] = dict(mm, was=mname)
            break
assert len(common) >= 12, "the re-planned program no longer resembles profiles/r1p_ops.md"
# calibrate BW_l1 on the ops the model calls L1-bound for any plausible bandwidth (L1 bytes >> HBM bytes)
l1b = [r for r in common if r["l1"] / r["hbm"] > 12 and measured[r["name"]]["ms"] > 0.15]
bw_l1 = np.median([r["l1"] / (measured[r["name"]]["ms"] * 1e-3) for r in l1b])


def model_ms(r):
    return max(r["hbm"] / BW_HBM, r["l1"] / bw_l1) * 1e3


lines = ["# Model of QXB_MIN_LOB on the headline step (CPU-only analysis, `scripts/model_min_lob.py`)", "",
         "RQC 7x7 d20 c64, plan of `bench.py` (`replan=128`), 131072 bitstrings.  Per op `t = max(HBM bytes / 6.0 TB/s, L1 bytes / BW_l1)`,",
         "L1 bytes = operand loads x 16 B with loads per output = `(2^ma + 2^nb) K / 2^(ma+nb)`.",
         f"`BW_l1` = **{bw_l1 / 1e12:.1f} TB/s**, the median over the {len(l1b)} measured ops with L1 bytes > 12x HBM bytes",
         "(`profiles/r1p_ops.md`; 148 SMs x 128 B/clk x 1.9 GHz = 36 TB/s nominal).", "",
         "## Calibration: model vs measured (default split), ops >= 0.15 ms", "",
         "| op | nC | nK | b/M/N | lob/ma/nb/kc | loads per output | measured ms | model ms | model / measured |", "|---|---:|---:|---|---|---:|---:|---:|---:|"]
tot_meas = tot_model = 0.0
for r in sorted(common, key=lambda r: -measured[r["name"]]["ms"]):
    ms = measured[r["name"]]["ms"]
    tot_meas += ms
    tot_model += model_ms(r)
    if ms >= 0.15:
        lines.append(f"| {r['name']} (r1p: {measured[r['name']].get('was', r['name'])}) | {r['nC']} | {r['nK']} | {r['split']} | {r['lob']}/{r['ma']}/{r['nb']}/{r['kc']} | "
                     f"{r['l1'] / ES / (N_AMP * 2.0 ** r['nC']):.1f} | {ms:.3f} | {model_ms(r):.3f} | {model_ms(r) / ms:.2f} |")
lines += ["", f"Sum over the {len(common)} matched ops: measured {tot_meas:.2f} ms, model {tot_model:.2f} ms "
              f"(ratio {tot_model / tot_meas:.2f}); the model has no latency term, so it under-estimates the ops that ncu shows as",
          "latency-bound (R474, R571: 25 % occupancy, no eligible warp half of the cycles).", "",
          "## Prediction per setting", "", "| setting | modelled step ms | vs default | ops whose split changes | largest changes (op: loads per output, model ms) |",
          "|---|---:|---:|---:|---|"]
base_by = {r["name"]: r for r in base}
base_total = sum(model_ms(r) for r in base)
for ml in (None, 7, 6, 5):
    rows = plan(ml)
    total = sum(model_ms(r) for r in rows)
    changed = [r for r in rows if (r["ma"], r["nb"], r["lob"]) != (base_by[r["name"]]["ma"], base_by[r["name"]]["nb"], base_by[r["name"]]["lob"])]
    changed.sort(key=lambda r: model_ms(base_by[r["name"]]) - model_ms(r), reverse=True)
    desc = "; ".join(f"{r['name']}: {base_by[r['name']]['l1'] / ES / (N_AMP * 2.0 ** r['nC']):.0f} -> {r['l1'] / ES / (N_AMP * 2.0 ** r['nC']):.0f}, "
                     f"{model_ms(base_by[r['name']]):.2f} -> {model_ms(r):.2f}" for r in changed[:5])
    lines.append(f"| {'default (8)' if ml is None else ml} | {total:.2f} | {total / base_total:.3f} | {len(changed)} | {desc} |")
os.environ.pop("QXB_MIN_LOB", None)
lines += ["", "The prediction is an upper bound on the gain (L1 term only).  Functional equivalence of the changed splits is checked on the CPU by",
          "`tests/test_templates.py` (kernel index arithmetic replayed on the library's own launch templates); the GPU A/B is",
          "`PROBE_CONFIGS=lob python scripts/probe_variants.py`."]
open(os.path.join(ROOT, "profiles", "r1q_model_min_lob.md"), "w").write("\n".join(lines) + "\n")
print("\n".join(lines))
