"""Compute-graph emission: the definition of the ops the B200 executor runs.

Host-side mirror of ``/root/reference/src/compute_graph/compute_graph.jl:15-98``
(``build_compute_graph``) and ``tensor_cache.jl:6-111`` (``TensorCache``), plus the
DSL writer the reference gets from QXContexts (``generate_dsl_files``, call site
``src/simulation.jl:73``; format ``docs/src/users_guide.md:93-164``).

Everything here is symbolic/integer bookkeeping and must be bit-exact:
names, view chains ``t -> t_s -> t_s_s`` (:45), 1-based ``index_position`` among the
tensor's (hyper-reduced) modes (:46-48), per-command integer labels, post-order
statement order, ``save output <root>`` (:94).
"""
from __future__ import annotations

import dataclasses
from collections import OrderedDict
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from .tn import TensorNetworkCircuit

DSL_VERSION = "0.4.0"


# --------------------------------------------------------------------------- cache
class TensorCache:
    """Unique leaf tensors keyed by shape; match if max|delta| < eps(Float64)
    (tensor_cache.jl:51-69; tolerance pinned by test/test_compute_graph.jl:12-13)."""

    def __init__(self, label: str = "data_"):
        self.tensors: Dict[Tuple[int, ...], List[Tuple[str, np.ndarray]]] = {}
        self.key_dim_map: "OrderedDict[str, Tuple[int, ...]]" = OrderedDict()
        self.id_val = 1
        self.label = label

    def _next_symbol(self) -> str:
        s = f"{self.label}{self.id_val}"
        self.id_val += 1
        return s

    def push(self, data: np.ndarray) -> str:
        data = np.asarray(data)
        dim = tuple(int(d) for d in data.shape)
        flat = data.reshape(-1, order="F").astype(np.complex128)   # column-major (:52-53)
        bucket = self.tensors.setdefault(dim, [])
        eps = np.finfo(np.float64).eps
        for sym, other in bucket:
            if flat.size == 0 or np.max(np.abs(flat - other)) < eps:
                return sym
        sym = self._next_symbol()
        bucket.append((sym, flat))
        self.key_dim_map[sym] = dim
        return sym

    def __getitem__(self, sym: str) -> np.ndarray:
        if sym not in self.key_dim_map:
            raise KeyError(f"No symbol {sym} in cache")
        dim = self.key_dim_map[sym]
        for s, flat in self.tensors[dim]:
            if s == sym:
                return flat.reshape(dim, order="F")
        raise KeyError(sym)

    def __len__(self):
        return len(self.key_dim_map)

    def to_dict(self) -> "OrderedDict[str, np.ndarray]":
        return OrderedDict((k, self[k]) for k in self.key_dim_map)


def save_cache(tc: TensorCache, filename: str) -> None:
    """tensor_cache.jl:90-106: one JLD2 dataset per label holding the N-d ComplexF64 array,
    written by the library's native JLD2 writer (``csrc/qxb_jld2.cpp``); the suffix must be
    ``.jld2`` as in the reference (:102).  ``.npz`` is still accepted: the committed benchmark
    triples under ``workloads/`` predate the native writer."""
    if filename.endswith(".npz"):
        np.savez(filename, **{k: np.asfortranarray(v) for k, v in tc.to_dict().items()})
        return
    if not filename.endswith(".jld2"):
        raise ValueError('Filename must have suffix ".jld2"')
    from .jld2 import save_jld2
    save_jld2(filename, {k: np.asarray(v, dtype=np.complex128) for k, v in tc.to_dict().items()})


# ------------------------------------------------------------------------ commands
@dataclasses.dataclass
class LoadCommand:
    name: str
    label: str
    dims: Tuple[int, ...]

    def dsl(self):
        return f"load {self.name} {self.label} {','.join(map(str, self.dims))}"


@dataclasses.dataclass
class OutputCommand:
    name: str
    idx: int
    dim: int

    def dsl(self):
        return f"output {self.name} {self.idx} {self.dim}"


@dataclasses.dataclass
class ViewCommand:
    name: str
    target: str
    slice_sym: str
    bond_index: int
    bond_dim: int

    def dsl(self):
        return f"view {self.name} {self.target} {self.slice_sym} {self.bond_index} {self.bond_dim}"


def _lab(ls: Sequence[int]) -> str:
    return ",".join(map(str, ls)) if len(ls) else "0"


@dataclasses.dataclass
class ContractCommand:
    output_name: str
    output_idxs: List[int]
    left_name: str
    left_idxs: List[int]
    right_name: str
    right_idxs: List[int]

    def dsl(self):
        return (f"ncon {self.output_name} {_lab(self.output_idxs)} {self.left_name} {_lab(self.left_idxs)} "
                f"{self.right_name} {_lab(self.right_idxs)}")


@dataclasses.dataclass
class SaveCommand:
    label: str
    name: str

    def dsl(self):
        return f"save {self.label} {self.name}"


class ComputeNode:
    def __init__(self, op):
        self.op = op
        self.children: List["ComputeNode"] = []
        self.parent: Optional["ComputeNode"] = None

    def __len__(self):
        """Number of nodes in the subtree (test/test_compute_graph.jl:33)."""
        n, stack = 0, [self]
        while stack:
            x = stack.pop()
            n += 1
            stack.extend(x.children)
        return n

    def post_order(self) -> List["ComputeNode"]:
        out, stack = [], [(self, False)]
        while stack:
            node, done = stack.pop()
            if done:
                out.append(node)
            else:
                stack.append((node, True))
                for ch in reversed(node.children):
                    stack.append((ch, False))
        return out


class ComputeGraph:
    def __init__(self, root: ComputeNode, tensors: Dict[str, np.ndarray]):
        self.root = root
        self.tensors = tensors
        self.root_indices: List[int] = []

    def commands(self):
        return [n.op for n in self.root.post_order()]

    def dsl(self, metadata=None) -> str:
        return write_dsl(self, metadata)


def _meta_lines(meta, indent=0) -> List[str]:
    out = []
    for k, v in meta.items():
        pad = "  " * indent
        if isinstance(v, dict):
            out.append(f"# {pad}{k}:")
            out.extend(_meta_lines(v, indent + 1))
        elif isinstance(v, (list, tuple)):
            out.append(f"# {pad}{k}:")
            out.extend(f"# {pad}  - {x}" for x in v)
        else:
            out.append(f'# {pad}{k}: {v}')
    return out


def write_dsl(cg: ComputeGraph, metadata=None) -> str:
    """Post-order walk -> ``.qx`` text (users_guide.md:48-90)."""
    lines = [f"# version: {DSL_VERSION}"]
    if metadata:
        lines.extend(_meta_lines(metadata))
        lines.append("#")
    lines.extend(op.dsl() for op in cg.commands())
    return "\n".join(lines) + "\n"


# ------------------------------------------------------------------------- builder
def _contraction_labels(net: "OrderedDict[str, List[int]]", a: str, b: str, cnt: Dict[int, int]):
    """Integer labels for contracting a with b given what else is still in the
    network (QXTns ``contraction_indices``; call site compute_graph.jl:65).
    An index shared by a and b survives iff some third tensor still carries it
    (hyper-edge, users_guide.md:146)."""
    ia, ib = net[a], net[b]
    sa, sb = set(ia), set(ib)
    # cnt[i] = number of tensors currently carrying index i
    others = {i for i in sa & sb if cnt[i] > 2}
    label: Dict[int, int] = {}
    for i in list(ia) + list(ib):
        if i not in label:
            label[i] = len(label) + 1
    c_ids = [i for i in ia if (i not in sb) or (i in others)]
    c_ids += [i for i in ib if i not in sa]
    return ([label[i] for i in c_ids], [label[i] for i in ia], [label[i] for i in ib], c_ids)


def build_compute_graph(tnc: TensorNetworkCircuit, plan: Sequence[Tuple[str, str, str]],
                        bond_groups: Optional[Sequence[Sequence[int]]] = None) -> ComputeGraph:
    """compute_graph.jl:15-98."""
    tnc = tnc.copy()
    net: "OrderedDict[str, List[int]]" = OrderedDict((s, list(t.indices)) for s, t in tnc.tensors.items())
    nodes: Dict[str, ComputeNode] = {}
    tc = TensorCache()

    for t in tnc.keys():                                      # :25-29
        data = tnc.tensor_data(t)
        sym = tc.push(data)
        nodes[t] = ComputeNode(LoadCommand(t, sym, tuple(data.shape)))
    for i, o in enumerate(tnc.output_tensors(), start=1):      # :32-35
        nodes[o] = ComputeNode(OutputCommand(o, i, tnc[o].shape[0]))

    changed: Dict[str, str] = {}
    if bond_groups is not None:                                # :39-58
        for gi, bg in enumerate(bond_groups, start=1):
            bgs = set(bg)
            related = [s for s, ids in net.items() if bgs & set(ids)]
            slice_sym = f"v{gi}"
            slice_dim = tnc.index_dim[list(bg)[0]]
            for t in related:
                new_sym = f"{t}_s"
                pos = next(k for k, i in enumerate(net[t], start=1) if i in bgs)
                node = ComputeNode(ViewCommand(new_sym, t, slice_sym, pos, slice_dim))
                # replace_tensor_symbol! keeps the tensor's place in the network
                net = OrderedDict((new_sym if k == t else k, v) for k, v in net.items())
                changed[t] = new_sym
                node.children.append(nodes[t])
                nodes[t].parent = node
                nodes[new_sym] = node

    def resolve(s):
        while s in changed:
            s = changed[s]
        return s

    cnt: Dict[int, int] = {}
    for ids in net.values():
        for i in ids:
            cnt[i] = cnt.get(i, 0) + 1

    # open wires (no input / no output tensors attached) are held by the outside world: count one extra carrier so
    # that an open index shared by the last two tensors of its hyper-edge stays a mode of the result instead of
    # being summed (QXTns keeps a distinct open Index per wire end; ids are merged per hyper-edge here)
    open_ids = set()
    if not tnc.output_tensors_:
        open_ids |= set(tnc.wire)
    if not tnc.input_tensors:
        open_ids |= set(tnc.wire_first)
    for i in open_ids:
        if i in cnt:
            cnt[i] += 1

    def contract(a, b, c):
        c_l, a_l, b_l, c_ids = _contraction_labels(net, a, b, cnt)
        for i in net[a] + net[b]:
            cnt[i] -= 1
        for i in c_ids:
            cnt[i] += 1
        node = ComputeNode(ContractCommand(c, c_l, a, a_l, b, b_l))
        for s in (a, b):
            nodes[s].parent = node
            node.children.append(nodes[s])
        nodes[c] = node
        # mock contraction: only the index bookkeeping evolves (:73)
        del net[a], net[b]
        net[c] = c_ids

    for (A, B, C) in plan:                                     # :61-74
        contract(resolve(A), resolve(B), C)

    parentless = [s for s in net.keys() if nodes[s].parent is None]     # :77-90
    root = parentless[0]
    for y in parentless[1:]:
        new = tnc.next_tensor_id()
        contract(root, y, new)
        root = new
    left = [s for s in nodes if nodes[s].parent is None]
    assert left == [root], "Only root node should have no parent"
    node = ComputeNode(SaveCommand("output", root))            # :94
    node.children.append(nodes[root])
    nodes[root].parent = node
    cg = ComputeGraph(node, tc.to_dict())
    # index id behind each mode of the saved tensor (empty for a closed network); a hyper-edge is ONE id here, so
    # several open wires of the circuit may map to the same mode (contract_tn expands them, executor.py)
    cg.root_indices = list(net[root])
    return cg
