"""Measured choice among exact alternatives of one program.

The re-planner (``qxb_graph_replan``) and the kernel-selection knobs decide by MODEL; the models are calibrated on
measurements (profiles/r1p_ops.md) but the tree search is noisy at the few-per-cent level between settings, and a knob
that helps one plan can hurt another.  ``tune`` turns the models' proposals into candidates and lets the GPU pick:

* trees: the planner under several settings of its L1 term (``QXB_PLAN_L1_BW``: default 19.9 TB/s, 0 = single-rate
  model, 14, 30) -- every tree is an exact re-association of the same network;
* register-tile knob ``min_lob`` = 8 / 7 / 6 (``qxb_options``: how many C bits stay thread bits: below 8 the nodes with
  <= 2^8 elements per bitstring get a register tile; same kernel, different launch template).

Every candidate is compiled, run on the same probe bitstrings and timed; one that raises, or whose amplitudes differ
from the baseline's, is discarded (``executor.autotune``).  Used by ``bench.py`` (device buffers, CUDA events, max over
ranks) and by ``execute(..., autotune=True)`` (host buffers, wall clock around a synchronised call).
"""
from __future__ import annotations

import os
import time
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

PLANNER_SETTINGS = (("l1model", None), ("r1pmodel", "0"), ("l1bw14", "14"), ("l1bw30", "30"))
MIN_LOBS = (8, 7, 6)


def plan_candidates(txt: str, data: Dict[str, np.ndarray], dtype: str, replan_candidates: int, n_amp: int,
                    first=None, settings=PLANNER_SETTINGS) -> List[Tuple[str, str, Optional[dict]]]:
    """``[(tag, plan text, replan_info)]``: one re-planned program per planner setting, duplicates dropped.
    ``first`` = an already re-planned ``Graph`` for the first setting (saves one search)."""
    from .executor import Graph
    plans: List[Tuple[str, str, Optional[dict]]] = []
    for tag, bw in settings:
        if first is not None and not plans:
            plans.append((tag, first.text, first.replan_info))
            continue
        saved = os.environ.get("QXB_PLAN_L1_BW")
        try:
            if bw is None:
                os.environ.pop("QXB_PLAN_L1_BW", None)
            else:
                os.environ["QXB_PLAN_L1_BW"] = bw
            g = Graph.from_dsl(txt, data, dtype, replan=replan_candidates, replan_n_amp=n_amp)
            if all(g.text != t for _, t, _ in plans):           # Graph.text is the re-planned program after a re-plan
                plans.append((tag, g.text, g.replan_info))
            del g
        except Exception:                                    # noqa: BLE001  (one tree less to choose from)
            pass
        finally:
            if saved is None:
                os.environ.pop("QXB_PLAN_L1_BW", None)
            else:
                os.environ["QXB_PLAN_L1_BW"] = saved
    return plans


def candidates_of(plans, min_lobs: Sequence[int] = MIN_LOBS):
    """(planner setting) x (min_lob) -> ``[(tag, plan text, compile options)]`` for ``executor.autotune``."""
    return [(f"{tag}/lob{lob}", text, {"min_lob": lob}) for tag, text, _ in plans for lob in min_lobs]


def tune(txt: str, data: Dict[str, np.ndarray], dtype: str, probe_bits: np.ndarray, replan_candidates: int = 32,
         compile_kw: Optional[dict] = None, rel_tol: Optional[float] = None):
    """-> (uncompiled ``Graph`` of the winning tree, report).  ``report["options"]`` = the winning compile options
    (pass them to ``Graph.compile``; nothing is left in ``os.environ``)."""
    from .executor import Graph, autotune, synchronize
    compile_kw = compile_kw or {}
    plans = plan_candidates(txt, data, dtype, replan_candidates, int(probe_bits.shape[0]))
    if not plans:
        return Graph.from_dsl(txt, data, dtype), {"error": "no plan candidate"}
    cands = candidates_of(plans)

    def build(text, **kw):
        return Graph.from_dsl(text, data, dtype).compile(**dict(compile_kw, **kw))

    def probe(g):
        for _ in range(2):
            out = g.amplitudes(probe_bits)
        synchronize()
        t0 = time.perf_counter()
        for _ in range(3):
            out = g.amplitudes(probe_bits)
        synchronize()
        return (time.perf_counter() - t0) * 1e3 / 3.0, out

    best, report = autotune(cands, build, probe, None, rel_tol if rel_tol is not None else (1e-4 if dtype == "c32" else 1e-9))
    tag, text, kw = cands[best]
    g = Graph.from_dsl(text, data, dtype)
    g.replan_info = next(i for t, _, i in plans if tag.startswith(t + "/"))
    return g, {"chosen": tag, "options": kw, "probe_bitstrings": int(probe_bits.shape[0]), "candidates": report}
