"""executor.autotune: the control flow bench.py relies on to choose, by measurement, among exact alternatives
(re-planner models, kernel-selection knobs) -- exercised here with fake graphs and fake timings."""
import os

import numpy as np

from qxb200.executor import autotune


class Fake:
    def __init__(self, text, knob=None):
        self.text = text
        self.env = knob


def run(cands, times, results, reduce=None, raising=()):
    def build(text, fake_knob=None):
        if text in raising:
            raise RuntimeError("boom " + text)
        return Fake(text, fake_knob)

    def probe(g):
        key = (g.text, g.env)
        return times[key], results[key]
    return autotune(cands, build, probe, reduce)


def test_picks_fastest_exact_candidate_and_leaves_no_state():
    env_before = dict(os.environ)
    a = np.array([1 + 1j, 0.5])
    cands = [("base", "A", {}), ("knob7", "A", {"fake_knob": "7"}), ("planB", "B", {}), ("wrong", "C", {}), ("bad", "D", {})]
    times = {("A", None): 10.0, ("A", "7"): 8.0, ("B", None): 9.0, ("C", None): 1.0}
    results = {("A", None): a, ("A", "7"): a * (1 + 1e-13), ("B", None): a, ("C", None): a * 1.001}
    best, rep = run(cands, times, results, raising={"D"})
    assert best == 1 and rep[1]["ms"] == 8.0
    assert rep[3]["ms"] is None and "differs" in rep[3]["note"]           # fast but wrong: discarded
    assert rep[4]["ms"] is None and "boom" in rep[4]["note"]              # raising candidate: discarded, run continues
    assert dict(os.environ) == env_before                                  # knobs travel as compile options, not as environment


def test_baseline_failure_falls_back_to_index_zero():
    a = np.ones(2)
    best, rep = run([("base", "A", {}), ("other", "B", {})], {("B", None): 1.0}, {("B", None): a}, raising={"A"})
    assert best == 0 and rep[1]["ms"] is None and "baseline" in rep[1]["note"]


def test_multi_rank_reduction_decides():
    a = np.ones(2)
    cands = [("base", "A", {}), ("alt", "B", {})]
    times = {("A", None): 10.0, ("B", None): 9.0}
    results = {("A", None): a, ("B", None): a}
    # another rank saw the alternative fail (inf): the max over ranks keeps every rank on the baseline
    best, rep = run(cands, times, results, reduce=lambda ts: [max(t, o) for t, o in zip(ts, [10.5, float("inf")])])
    assert best == 0 and rep[1]["ms"] == 9.0 and rep[1]["ms_max_over_ranks"] is None
    best, _ = run(cands, times, results, reduce=lambda ts: [max(t, o) for t, o in zip(ts, [10.5, 9.5])])
    assert best == 1
