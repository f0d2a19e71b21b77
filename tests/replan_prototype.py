"""Batch-aware re-planning of a ``.qx`` program (planner-side lowering, host side).

QXTools plans for ONE slice of ONE bitstring at a time: ``contraction_scheme``
(/root/reference/src/contraction_planning.jl:219-299) removes the sliced
hyper-edges from the line graph and orders the rest.  The B200 executor does not
run one slice at a time -- sliced hyper-edges come back as batch bits and the
bitstring is one more batch axis -- so the order that is optimal per slice keeps
up to a dozen batch bits alive on every large intermediate.  On RQC 7x7 depth 20
with 12 sliced bonds that is 16.5 GB of traffic per 1024 bitstrings, against
2-3 GB for an order chosen for the batched network.

``replan_dsl`` is a DSL -> DSL rewrite.  It keeps every leaf statement (``load`` /
``output`` / ``view``) verbatim, recovers the tensor network behind the ``ncon``
tree (index identity = labels matched in an ``ncon``), and emits a new ``ncon``
tree from an elimination order (min-fill on the line graph,
contraction_planning.jl:47-63,127-176, with the bitstring axis as one more
hyper-edge shared by all ``output`` leaves) scored by the executor's own cost
model (bytes moved by the lowered program, ``qxb_graph_describe``).  The value of
the program -- sum over slices of the saved scalar -- is unchanged (exact
re-association; checked against the oracle in tests/test_replan.py).
"""
from __future__ import annotations

import time as _time
from collections import OrderedDict
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from qxb200.planning import min_fill

AMP = -1     # pseudo index: the bitstring axis carried by every output leaf


class _Leaf:
    __slots__ = ("name", "classes", "is_output")

    def __init__(self, name, classes, is_output):
        self.name, self.classes, self.is_output = name, classes, is_output


def _parse(text: str):
    lines = text.splitlines()
    if not lines or not lines[0].startswith("# version:"):
        raise ValueError("first line of a .qx file must be '# version: x.y.z'")
    header, stmts = [], []
    for ln in lines:
        s = ln.strip()
        if not s:
            continue
        if s.startswith("#"):
            header.append(ln)
        else:
            stmts.append(s.split())
    return header, stmts


def _labels(tok: str) -> List[int]:
    return [] if tok == "0" else [int(x) for x in tok.split(",")]


def recover_network(text: str):
    """-> (header, leaf statement lines, leaves, extent per class, sliced classes, root is scalar).
    A class is a set of tensor modes identified by matching labels in ``ncon`` lines."""
    header, stmts = _parse(text)
    parent: List[int] = []

    def new():
        parent.append(len(parent))
        return len(parent) - 1

    def find(x):
        while parent[x] != x:
            parent[x] = parent[parent[x]]
            x = parent[x]
        return x

    def union(a, b):
        a, b = find(a), find(b)
        if a != b:
            parent[b] = a

    modes: Dict[str, List[int]] = {}        # tensor name -> class handle per mode
    ext: Dict[int, int] = {}                 # class handle -> stored extent
    sliced: Dict[int, str] = {}              # class handle -> slice symbol
    is_out: Dict[str, bool] = {}
    used_as_operand = set()
    view_target = set()
    leaf_lines, leaf_names = [], []
    root = None
    for t in stmts:
        op = t[0]
        if op == "load":
            dims = [int(x) for x in t[3].split(",")]
            ms = []
            for d in dims:
                h = new(); ext[h] = d; ms.append(h)
            modes[t[1]] = ms; is_out[t[1]] = False
            leaf_lines.append(" ".join(t)); leaf_names.append(t[1])
        elif op == "output":
            h = new(); ext[h] = int(t[3])
            modes[t[1]] = [h]; is_out[t[1]] = True
            leaf_lines.append(" ".join(t)); leaf_names.append(t[1])
        elif op == "view":
            name, target, sym, pos = t[1], t[2], t[3], int(t[4])
            ms = list(modes[target])
            sliced[ms[pos - 1]] = sym
            modes[name] = ms; is_out[name] = is_out[target]
            view_target.add(target)
            leaf_lines.append(" ".join(t)); leaf_names.append(name)
        elif op == "ncon":
            out, cl, a, al, b, bl = t[1], _labels(t[2]), t[3], _labels(t[4]), t[5], _labels(t[6])
            used_as_operand.add(a); used_as_operand.add(b)
            by_label: Dict[int, int] = {}
            for ls, nm in ((al, a), (bl, b)):
                if len(ls) != len(modes[nm]):
                    raise ValueError(f"ncon {out}: label count does not match rank of {nm}")
                for l, h in zip(ls, modes[nm]):
                    if l in by_label:
                        union(by_label[l], h)
                    else:
                        by_label[l] = h
            modes[out] = [by_label[l] for l in cl]
            is_out[out] = False
        elif op == "save":
            root = t[2]
        else:
            raise ValueError(f"unknown instruction {op}")
    if root is None:
        raise ValueError("program has no save instruction")
    ncon_names = {t[1] for t in stmts if t[0] == "ncon"}
    leaves = []
    for nm in leaf_names:
        if nm in used_as_operand and nm not in ncon_names:
            leaves.append(_Leaf(nm, [find(h) for h in modes[nm]], is_out[nm]))
    dims = {}
    for h, d in ext.items():
        r = find(h)
        if dims.setdefault(r, d) != d:
            raise ValueError("extent mismatch inside an index class")
    sl = {find(h): s for h, s in sliced.items()}
    return header, leaf_lines, leaves, dims, sl, len(modes[root]) == 0


class _Net:
    """Duck-typed stand-in for TensorNetworkCircuit for the planning helpers."""

    def __init__(self, leaves: Sequence[_Leaf], dims: Dict[int, int], n_amp_weight: int):
        self.tensors = OrderedDict()
        self.index_dim = dict(dims)
        self.index_dim[AMP] = n_amp_weight
        for lf in leaves:
            ids = list(dict.fromkeys(lf.classes))
            if lf.is_output:
                ids.append(AMP)
            self.tensors[lf.name] = ids


def _line_graph(net: _Net):
    lg = {i: set() for ids in net.tensors.values() for i in ids}
    for ids in net.tensors.values():
        for a in ids:
            for b in ids:
                if a != b:
                    lg[a].add(b)
    return lg


def _plan_from_order(net: _Net, order: Sequence[int]):
    """Elimination order -> pairwise plan; inside a hyper-edge the pair with the smallest
    (batched) result goes first (contraction_planning.jl:386-448 uses netcon there)."""
    cur = OrderedDict((s, set(ids)) for s, ids in net.tensors.items())
    owners: Dict[int, set] = {}
    for s, ids in cur.items():
        for i in ids:
            owners.setdefault(i, set()).add(s)
    dim = net.index_dim
    plan, n_int = [], 0

    def size(ids):
        r = 1.0
        for i in ids:
            r *= dim[i]
        return r

    def result(a, b):
        ia, ib = cur[a], cur[b]
        res = set(ia | ib)
        for i in ia & ib:
            if i != AMP and owners[i] <= {a, b}:
                res.discard(i)
        return res

    def contract(a, b):
        nonlocal n_int
        res = result(a, b)
        n_int += 1
        c = f"R{n_int}"
        plan.append((a, b, c))
        for i in cur[a] | cur[b]:
            owners[i].discard(a); owners[i].discard(b)
        for i in res:
            owners[i].add(c)
        del cur[a], cur[b]
        cur[c] = res
        return c

    for ix in list(order) + [AMP]:
        group = sorted(owners.get(ix, ()), key=lambda s: (size(cur[s]), s))
        while len(group) > 1:
            best = None
            for x in range(len(group)):
                for y in range(x + 1, len(group)):
                    key = (size(result(group[x], group[y])), x, y)
                    if best is None or key < best[0]:
                        best = (key, x, y)
            _, x, y = best
            c = contract(group[x], group[y])
            group = [g for k, g in enumerate(group) if k not in (x, y)] + [c]
    rest = list(cur.keys())                      # disconnected components (scalars): join them
    while len(rest) > 1:
        rest = [contract(rest[0], rest[1])] + rest[2:]
    return plan, rest[0]


def _emit(header, leaf_lines, leaves, plan, root) -> str:
    ids: Dict[str, List[int]] = {lf.name: list(lf.classes) for lf in leaves}
    cnt: Dict[int, int] = {}
    for ls in ids.values():
        for i in set(ls):
            cnt[i] = cnt.get(i, 0) + 1
    lab = lambda ls: ",".join(map(str, ls)) if ls else "0"
    out = [header[0], "# re-planned for batched execution by qxb200.replan (leaves and views unchanged)"]
    out += [h for h in header[1:]]
    out += leaf_lines
    for a, b, c in plan:
        ia, ib = ids[a], ids[b]
        sa, sb = set(ia), set(ib)
        label: Dict[int, int] = {}
        for i in ia + ib:
            if i not in label:
                label[i] = len(label) + 1
        # an index survives iff some OTHER remaining tensor still carries it (hyper-edge rule of
        # compute_graph.jl:65 / users_guide.md:146); otherwise it is summed here
        others = lambda i: cnt[i] - (i in sa) - (i in sb)
        keep = [i for i in ia if others(i) > 0]
        seen = set(keep)
        keep += [i for i in ib if i not in seen and others(i) > 0]
        # a class repeated inside one operand (should not happen for leaves of QXTools files)
        if len(set(ia)) != len(ia) or len(set(ib)) != len(ib):
            raise ValueError("repeated index inside one tensor: cannot re-plan")
        out.append(f"ncon {c} {lab([label[i] for i in keep])} {a} {lab([label[i] for i in ia])} "
                   f"{b} {lab([label[i] for i in ib])}")
        for i in sa | sb:
            cnt[i] -= (i in sa) + (i in sb)
        for i in set(keep):
            cnt[i] += 1
        ids[c] = keep
    out.append(f"save output {root}")
    return "\n".join(out) + "\n"


def _cost(text: str, data_dims: Dict[str, Tuple[int, ...]], n_amp: int, dtype: str) -> float:
    """Bytes moved by the lowered program (the executor's own model, host-only call)."""
    from qxb200.executor import Graph
    g = Graph(dtype)
    b = text.encode()
    from qxb200._lib import check
    check(g._lib.qxb_graph_parse_dsl(g._h, b, len(b)))
    d = g.describe()
    es = 8 if g.dtype == 0 else 16
    tot = 0.0
    for o in d["ops"]:
        if o["phase"] == "const":
            continue
        tot += (2.0 ** o["nC"] * (n_amp if o["amp"] else 1) + 2.0 ** o["a_bits"] * (n_amp if o["a_amp"] else 1) +
                2.0 ** o["b_bits"] * (n_amp if o["b_amp"] else 1))
    return tot * es


def replan_dsl(text: str, n_amp: int = 1024, time: float = 30.0, seed: int = 0, dtype: str = "c64",
               verbose: bool = False, candidates: int = 24):
    """Return (new_text, info).  Tries the deterministic min-fill order and then randomised
    restarts (seeded) until ``candidates`` orders were scored or ``time`` seconds passed -- with
    the default bounds the count is what stops the search, so the result is reproducible.
    The original program is kept when nothing cheaper is found."""
    header, leaf_lines, leaves, dims, sliced, scalar = recover_network(text)
    if not scalar:
        return text, {"replanned": False, "reason": "saved tensor is not a scalar"}
    base = _cost(text, {}, n_amp, dtype)
    net = _Net(leaves, dims, max(2, n_amp))
    lg = _line_graph(net)
    amp_adj = lg.pop(AMP, set())
    for v in lg.values():
        v.discard(AMP)
    best_text, best_cost, tries = text, base, 0
    rng = np.random.default_rng(seed)
    t_end = _time.time() + max(0.0, time)
    first = True
    while first or (_time.time() < t_end and tries < candidates):
        _, order = min_fill(lg, None if first else rng)
        first = False
        plan, root = _plan_from_order(net, order)
        cand = _emit(header, leaf_lines, leaves, plan, root)
        c = _cost(cand, {}, n_amp, dtype)
        tries += 1
        if verbose:
            print(f"  candidate {tries}: {c / 1e9:.3f} GB (best {best_cost / 1e9:.3f}, given {base / 1e9:.3f})")
        if c < best_cost:
            best_text, best_cost = cand, c
    return best_text, {"replanned": best_text is not text, "given_bytes": base, "bytes": best_cost,
                       "candidates": tries, "n_amp_model": n_amp}
