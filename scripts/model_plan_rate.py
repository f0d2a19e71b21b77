"""CPU-side experiment: re-plan the headline workload with different FLOP rates in the tree search's time model
(QXB_PLAN_FLOP_RATE, TFLOP/s) and score every resulting plan with the SAME calibrated per-op model
(scripts/model_min_lob.py: t = max(HBM bytes / 6.0 TB/s, operand-load bytes / BW_l1), BW_l1 from profiles/r1p_ops.md).
Prints one line per setting; no GPU needed."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench                                   # noqa: E402
from qxb200.executor import Graph              # noqa: E402
from template_emulator import templates        # noqa: E402

N_AMP, ES, BW_HBM = int(os.environ.get("MODEL_N_AMP", "131072")), 16, 6.0e12
BW_L1 = float(os.environ.get("MODEL_BW_L1", "19.9e12"))
txt, data, w = bench.build_workload(os.environ.get("MODEL_WORKLOAD", "rqc_7x7_d20_c64_s4096"))
ES = 16 if w["dtype"] == "c64" else 8


def score(rate, min_lob=None):
    for k, v in (("QXB_PLAN_L1_BW", rate), ("QXB_MIN_LOB", min_lob)):
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = str(v)
    g = Graph.from_dsl(txt, data, w["dtype"], replan=int(os.environ.get("MODEL_CANDIDATES", "128")), replan_n_amp=N_AMP)
    d = g.describe()
    hbm_t = l1_t = t = hbm_b = flops = 0.0
    worst = []
    for p, op in zip(templates(g), d["ops"]):
        if op["phase"] != "chunk":
            continue
        out = N_AMP * 2.0 ** p.nC
        hbm = ES * ((2.0 ** op["a_bits"]) * (N_AMP if op["a_amp"] else 1) + (2.0 ** op["b_bits"]) * (N_AMP if op["b_amp"] else 1) + out)
        l1 = ES * out * ((1 << p.ma) + (1 << p.nb)) * 2.0 ** p.nK / (1 << (p.ma + p.nb))
        ti = max(hbm / BW_HBM, l1 / BW_L1)
        t += ti; hbm_b += hbm; flops += 8 * out * 2.0 ** p.nK
        hbm_t += hbm / BW_HBM; l1_t += l1 / BW_L1
        worst.append((ti, op["name"], p.nC, p.nK, f'{op["batch_bits"]}/{op["m_bits"]}/{op["n_bits"]}', p.ma, p.nb))
    worst.sort(reverse=True)
    print(f"rate {rate or 'default':>7} min_lob {min_lob or 8}: model {t * 1e3:6.2f} ms | HBM {hbm_b / 1e9:6.1f} GB ({hbm_t * 1e3:5.2f} ms) | "
          f"{flops / 1e9:7.1f} GFLOP | L1 term {l1_t * 1e3:6.2f} ms | planner bytes {g.replan_info['bytes'] / 1e9:.1f} GB", flush=True)
    print("      top: " + "  ".join(f"{n}[{c},{k},{s},t{ma}{nb}] {ti * 1e3:.2f}" for ti, n, c, k, s, ma, nb in worst[:6]), flush=True)


for rate in [None] + [float(x) for x in os.environ.get("MODEL_RATES", "0,30,14").split(",") if x]:
    score(rate)
for lob in [int(x) for x in os.environ.get("MODEL_LOBS", "").split(",") if x]:     # QXB_MIN_LOB on the default plan
    score(None, lob)
