"""Contraction planning and bond slicing (planner side of the hot path).

Host-side mirror of ``/root/reference/src/contraction_planning.jl``:
line graph of (hyper)edges (:47-63), min-fill plan (:127-176),
``flow_cutter_contraction_plan`` (:90-113), ``contraction_scheme`` = plan + greedy
treewidth-deletion slicing (:219-299) and elimination order -> pairwise plan
(:386-448).  The reference shells out to the FlowCutter binary through
QXGraphDecompositions (neither is available here) and falls back to min-fill
when FlowCutter returns nothing (:102-106); this module *is* that fallback, made
stronger with wall-clock-bounded randomised restarts, so ``time`` keeps its
meaning ("seconds spent looking for a better tree decomposition").

The executor only consumes the *outputs* of this module: a plan
``[(A, B, C), ...]`` and the sliced bond groups.  In this mirror every hyper-edge
is a single index id, so a bond group is ``[index_id]``.
"""
from __future__ import annotations

import time as _time
from collections import OrderedDict
from typing import Dict, Iterable, List, Optional, Sequence, Set, Tuple

import numpy as np

from .tn import TensorNetworkCircuit

Plan = List[Tuple[str, str, str]]


def convert_to_graph(tnc: TensorNetworkCircuit) -> Dict[str, Set[str]]:
    """Tensor adjacency graph (contraction_planning.jl:16-33)."""
    adj: Dict[str, Set[str]] = {s: set() for s in tnc.keys()}
    owners: Dict[int, List[str]] = {}
    for s, t in tnc.tensors.items():
        for i in t.indices:
            owners.setdefault(i, []).append(s)
    for ts in owners.values():
        for a in ts:
            for b in ts:
                if a != b:
                    adj[a].add(b)
    return adj


def convert_to_line_graph(tnc: TensorNetworkCircuit, use_hyperedges: bool = True,
                          exclude: Iterable[int] = ()) -> Dict[int, Set[int]]:
    """Line graph: one vertex per (hyper)edge = index id, two vertices adjacent
    when their indices meet in a tensor (contraction_planning.jl:47-63)."""
    exclude = set(exclude)
    lg: Dict[int, Set[int]] = {i: set() for i in tnc.index_dim if i not in exclude}
    for t in tnc.tensors.values():
        ids = [i for i in t.indices if i not in exclude]
        for a in ids:
            for b in ids:
                if a != b:
                    lg[a].add(b)
    return lg


def _eliminate(adj: Dict[int, Set[int]], order: Sequence[int]):
    """Width and bags of an elimination order."""
    adj = {v: set(n) for v, n in adj.items()}
    bags = []
    width = 0
    for v in order:
        nb = adj.pop(v)
        bags.append((v, frozenset(nb)))
        width = max(width, len(nb))
        for a in nb:
            adj[a].discard(v)
            adj[a] |= nb
            adj[a].discard(a)
    return width, bags


def min_fill(adj: Dict[int, Set[int]], rng: Optional[np.random.Generator] = None):
    """Min-fill elimination order; returns (treewidth upper bound, order).
    With ``rng`` ties are broken at random (used by the restarts)."""
    adj = {v: set(n) for v, n in adj.items()}
    order: List[int] = []
    width = 0

    def fill(v):
        nb = list(adj[v])
        f = 0
        for i in range(len(nb)):
            ai = adj[nb[i]]
            for j in range(i + 1, len(nb)):
                if nb[j] not in ai:
                    f += 1
        return f

    cost = {v: fill(v) for v in adj}
    while adj:
        best = min(cost.values())
        cands = [v for v, c in cost.items() if c == best]
        if rng is not None and len(cands) > 1:
            md = min(len(adj[v]) for v in cands)
            cands = [v for v in cands if len(adj[v]) == md]
            v = cands[int(rng.integers(len(cands)))]
        else:
            v = min(cands, key=lambda x: (len(adj[x]), x))
        nb = adj.pop(v)
        cost.pop(v)
        order.append(v)
        width = max(width, len(nb))
        touched = set(nb)
        for a in nb:
            adj[a].discard(v)
            new = nb - adj[a] - {a}
            if new:
                adj[a] |= new
                touched |= adj[a]
        for a in nb:
            touched |= adj[a]
        for a in touched:
            if a in adj:
                cost[a] = fill(a)
    return width, order


def _order_cost(tnc: TensorNetworkCircuit, adj, order) -> float:
    """log2-sum-exp style total size of the bags (proxy for contraction flops)."""
    _, bags = _eliminate(adj, order)
    tot = 0.0
    for v, nb in bags:
        b = 1.0
        for i in list(nb) + [v]:
            b *= tnc.index_dim.get(i, 2)
        tot += b
    return tot


def best_order(tnc: TensorNetworkCircuit, adj, time: float = 0.0, seed: Optional[int] = None):
    """Deterministic min-fill, then randomised restarts for ``time`` seconds."""
    width, order = min_fill(adj)
    cost = _order_cost(tnc, adj, order)
    rng = np.random.default_rng(None if seed in (None, -1) else seed)
    t_end = _time.time() + max(0.0, float(time))
    tries = 0
    while _time.time() < t_end:
        w, o = min_fill(adj, rng)
        c = _order_cost(tnc, adj, o)
        tries += 1
        if (w, c) < (width, cost):
            width, order, cost = w, o, c
    return width, order, {"restarts": tries, "cost": cost}


def order_to_contraction_plan(order: Sequence[int], tnc: TensorNetworkCircuit,
                              skip: Iterable[int] = ()) -> Plan:
    """Elimination order over index ids -> pairwise plan (contraction_planning.jl:386-448).
    Eliminating an index contracts every tensor that still carries it; hyper-edges
    with more than two tensors are contracted smallest-result-first (the reference
    uses netcon there, :427-435)."""
    skip = set(skip)
    cur: "OrderedDict[str, Set[int]]" = OrderedDict((s, set(t.indices)) for s, t in tnc.tensors.items())
    owners: Dict[int, Set[str]] = {}
    for s, ids in cur.items():
        for i in ids:
            owners.setdefault(i, set()).add(s)
    dim = tnc.index_dim
    plan: Plan = []
    n_int = 0

    def size(ids):
        r = 1
        for i in ids:
            if i not in skip:
                r *= dim[i]
        return r

    for ix in order:
        group = sorted(owners.get(ix, ()), key=lambda s: (size(cur[s]), s))
        while len(group) > 1:
            best = None
            for a in range(len(group)):
                for b in range(a + 1, len(group)):
                    ia, ib = cur[group[a]], cur[group[b]]
                    res = set(ia | ib)
                    for i in ia & ib:
                        if owners[i] <= {group[a], group[b]}:
                            res.discard(i)
                    key = (size(res), size(ia | ib), a, b)
                    if best is None or key < best[0]:
                        best = (key, a, b, res)
            _, a, b, res = best
            A, B = group[a], group[b]
            n_int += 1
            C = f"I{n_int}"
            plan.append((A, B, C))
            for i in cur[A] | cur[B]:
                owners[i].discard(A)
                owners[i].discard(B)
            for i in res:
                owners[i].add(C)
            del cur[A], cur[B]
            cur[C] = res
            group = [g for k, g in enumerate(group) if k not in (a, b)] + [C]
    return plan


def min_fill_contraction_plan(tnc: TensorNetworkCircuit, hypergraph: bool = True) -> Plan:
    """contraction_planning.jl:127-176."""
    lg = convert_to_line_graph(tnc, use_hyperedges=hypergraph)
    _, order = min_fill(lg)
    return order_to_contraction_plan(order, tnc)


def flow_cutter_contraction_plan(tnc: TensorNetworkCircuit, time: float = 10, seed: int = -1,
                                 hypergraph: bool = True) -> Plan:
    """contraction_planning.jl:90-113.  No FlowCutter binary here: this is the
    reference's own min-fill fallback (:102-106) plus timed random restarts."""
    lg = convert_to_line_graph(tnc, use_hyperedges=hypergraph)
    _, order, _ = best_order(tnc, lg, time=time, seed=seed)
    return order_to_contraction_plan(order, tnc)


def _greedy_treewidth_deletion(tnc, lg, num, order, score_function):
    """Pick ``num`` line-graph vertices (= hyper-edges) to remove
    (QXGraphDecompositions.greedy_treewidth_deletion; call site
    contraction_planning.jl:254-257)."""
    adj = {v: set(n) for v, n in lg.items()}
    order = list(order)
    removed: List[int] = []
    widths: List[int] = []
    for _ in range(num):
        if not adj:
            break
        if score_function == "degree":
            v = max(adj, key=lambda x: (len(adj[x]), -x))
        elif score_function == "direct_treewidth":
            cands = sorted(adj, key=lambda x: (-len(adj[x]), x))[:24]
            best = None
            for c in cands:
                sub = {a: (n - {c}) for a, n in adj.items() if a != c}
                w, _ = _eliminate(sub, [o for o in order if o != c])
                if best is None or (w, c) < best:
                    best = (w, c)
            v = best[1]
        else:   # tree_trimming: the vertex sitting in most of the largest bags
            w, bags = _eliminate(adj, order)
            count: Dict[int, float] = {}
            for u, nb in bags:
                if len(nb) >= w - 1:
                    wgt = 4.0 if len(nb) == w else 1.0
                    for x in list(nb) + [u]:
                        count[x] = count.get(x, 0.0) + wgt
            v = max(count, key=lambda x: (count[x], len(adj[x]), -x))
        removed.append(v)
        for a in adj.pop(v):
            adj[a].discard(v)
        order = [o for o in order if o != v]
        widths.append(_eliminate(adj, order)[0])
    return adj, removed, order, widths


def contraction_scheme(tnc: TensorNetworkCircuit, num: int, time: float = 10, seed: int = -1,
                       score_function: str = "tree_trimming", hypergraph: bool = True):
    """contraction_planning.jl:219-299 -> (bond_groups, plan, metadata)."""
    lg = convert_to_line_graph(tnc, use_hyperedges=hypergraph)
    tw, order, info = best_order(tnc, lg, time=time, seed=seed)
    cmeta = OrderedDict()
    cmeta["Method used"] = "min fill heuristic (randomised restarts)"
    cmeta["Treewidth"] = tw
    cmeta["Time allocated"] = time
    cmeta["Seed used"] = seed
    cmeta["Returned metadata"] = {1: f"restarts {info['restarts']}"}
    cmeta["Hypergraph used"] = hypergraph
    cmeta["Hyperedge contraction method"] = "smallest result first"
    sliced_lg, removed, new_order, widths = _greedy_treewidth_deletion(tnc, lg, num, order, score_function)
    if removed:
        # re-plan the sliced network: the sliced indices have extent 1 from here on
        w2, o2, _ = best_order(tnc, sliced_lg, time=min(time, 2.0) if time else 0.0, seed=seed)
        if (w2, _order_cost(tnc, sliced_lg, o2)) < (_eliminate(sliced_lg, new_order)[0],
                                                     _order_cost(tnc, sliced_lg, new_order)):
            new_order = o2
            if widths:
                widths[-1] = w2
    smeta = OrderedDict()
    smeta["Method used"] = "greedy treewidth deletion"
    smeta["Edges sliced"] = num
    smeta["Score fucntion used"] = score_function
    smeta["Treewidths after slicing consecutive edges"] = widths
    # sliced edges are absent from the order (contraction_planning.jl:267-270): their
    # extent-1 modes ride along; tensors left connected only through them are joined
    # by build_compute_graph (compute_graph.jl:77-90)
    plan = order_to_contraction_plan(list(new_order), tnc, skip=removed)
    bond_groups = [[v] for v in removed]
    meta = OrderedDict()
    meta["Determination of contraction plan"] = cmeta
    meta["Slicing"] = smeta
    return bond_groups, plan, meta
