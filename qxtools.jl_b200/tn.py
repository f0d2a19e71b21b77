"""Tensor-network circuit with hyper-indices (host-side mirror of the QXTns
structures the reference's hot path is driven through).

Mirrors ``/root/reference/src/tn_conversion.jl:10-40`` (``convert_to_tnc``:
one ``push!`` per gate, then ``add_input!``/``add_output!``) and the QXTns behaviour
the compute-graph builder relies on (``src/compute_graph/compute_graph.jl:26-27,46-48``):
``tensor_data`` is stored *hyper-index reduced* -- modes that are forced equal
(diagonal gates, control halves of controlled gates) are merged into one mode, so
a hyper-edge is simply an index id shared by two or more tensors.

Counts pinned by the reference's tests (``test/test_tn_conversion.jl:12,18,24``):
the 3-qubit GHZ circuit gives 3 tensors undecomposed, 5 decomposed, and 6 more
with inputs and outputs.
"""
from __future__ import annotations

import copy as _copy
from collections import OrderedDict
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from .circuits import Circuit, gate_matrix, gate_qubits

_TOL = 1e-12


class TNTensor:
    __slots__ = ("indices", "data")

    def __init__(self, indices: Sequence[int], data: np.ndarray):
        self.indices = tuple(indices)
        self.data = np.asarray(data, dtype=np.complex128)
        assert self.data.ndim == len(self.indices)
        assert len(set(self.indices)) == len(self.indices)

    @property
    def shape(self):
        return self.data.shape


def _reduce_hyper(data: np.ndarray, ids: List[int], forced: Sequence[Tuple[int, int]] = ()):
    """Merge modes that are forced equal (tensor vanishes unless their values
    agree).  ``ids`` are per-mode index ids; merged modes keep the OLDEST (smallest) id.
    Returns (reduced data, ids, {dropped id -> kept id})."""
    data = np.asarray(data)
    ids = list(ids)
    renames: Dict[int, int] = {}
    changed = True
    while changed:
        changed = False
        for i in range(len(ids)):
            for j in range(i + 1, len(ids)):
                if data.shape[i] != data.shape[j]:
                    continue
                d = data.shape[i]
                m = np.moveaxis(data, (i, j), (0, 1))
                off = m[~np.eye(d, dtype=bool)]
                if off.size and np.max(np.abs(off)) > _TOL:
                    continue
                # take the diagonal: new axis goes last -> move back to position i
                diag = np.diagonal(data, axis1=i, axis2=j)
                data = np.moveaxis(diag, -1, i)
                keep, drop = min(ids[i], ids[j]), max(ids[i], ids[j])
                for k, v in list(renames.items()):
                    if v == drop:
                        renames[k] = keep
                renames[drop] = keep
                ids[i] = keep
                ids.pop(j)
                changed = True
                break
            if changed:
                break
    return data, ids, renames


def _operator_schmidt(T4: np.ndarray):
    """T4[o1,i1,o2,i2] -> A[o1,i1,b], B[o2,i2,b] with sum_b A (x) B == T4."""
    d = T4.shape[0]
    M = T4.reshape(d * d, d * d)
    # controlled structure  sum_c |c><c| (x) U_c  -> exact projector decomposition
    # (keeps the control half a pure hyper-index, as in users_guide.md:78 "load t3 data_4 2")
    diag1 = all(np.max(np.abs(T4[o, i])) < _TOL for o in range(d) for i in range(d) if o != i)
    diag2 = all(np.max(np.abs(T4[:, :, o, i])) < _TOL for o in range(d) for i in range(d) if o != i)
    if diag1 and not diag2:
        A = np.zeros((d, d, d), dtype=np.complex128)
        B = np.zeros((d, d, d), dtype=np.complex128)
        for c in range(d):
            A[c, c, c] = 1
            B[:, :, c] = T4[c, c]
        return A, B
    if diag2 and not diag1:
        A = np.zeros((d, d, d), dtype=np.complex128)
        B = np.zeros((d, d, d), dtype=np.complex128)
        for c in range(d):
            B[c, c, c] = 1
            A[:, :, c] = T4[:, :, c, c]
        return A, B
    U, s, Vh = np.linalg.svd(M)
    r = int(np.sum(s > 1e-10 * s[0]))
    sq = np.sqrt(s[:r])
    A = (U[:, :r] * sq).reshape(d, d, r)
    B = (Vh[:r, :].T * sq).reshape(d, d, r)
    return A, B


class TensorNetworkCircuit:
    """Tensor network of a circuit.  ``tensors`` keeps insertion order
    (``keys(tnc)`` order drives leaf creation, compute_graph.jl:25)."""

    def __init__(self, qubits: int):
        self.qubits = int(qubits)
        self.tensors: "OrderedDict[str, TNTensor]" = OrderedDict()
        self.index_dim: Dict[int, int] = {}
        self._next_index = 0
        self._next_tensor = 0
        self.wire_first: List[int] = []
        self.wire: List[int] = []
        for _ in range(self.qubits):
            ix = self._new_index(2)
            self.wire_first.append(ix)
            self.wire.append(ix)
        self.input_tensors: List[str] = []
        self.output_tensors_: List[str] = []

    # -- bookkeeping -----------------------------------------------------
    def _new_index(self, dim: int) -> int:
        self._next_index += 1
        self.index_dim[self._next_index] = int(dim)
        return self._next_index

    def next_tensor_id(self) -> str:
        self._next_tensor += 1
        return f"t{self._next_tensor}"

    def _rename_index(self, old: int, new: int) -> None:
        """old and new are forced equal: fold ``old`` into ``new`` everywhere."""
        for t in self.tensors.values():
            if old in t.indices:
                assert new not in t.indices
                t.indices = tuple(new if i == old else i for i in t.indices)
        self.wire = [new if i == old else i for i in self.wire]
        self.wire_first = [new if i == old else i for i in self.wire_first]
        self.index_dim.pop(old, None)

    def _add_tensor(self, ids: List[int], data: np.ndarray):
        data, ids, renames = _reduce_hyper(data, ids)
        for old, new in renames.items():
            self._rename_index(old, new)
        sym = self.next_tensor_id()
        self.tensors[sym] = TNTensor(ids, data)
        return sym, renames

    def __len__(self):
        return len(self.tensors)

    def keys(self):
        return self.tensors.keys()

    def __getitem__(self, sym: str) -> TNTensor:
        return self.tensors[sym]

    def copy(self) -> "TensorNetworkCircuit":
        return _copy.deepcopy(self)

    def tensor_data(self, sym: str) -> np.ndarray:
        return self.tensors[sym].data

    def bonds(self) -> List[int]:
        return sorted(self.index_dim)

    def tensors_with(self, index: int) -> List[str]:
        return [s for s, t in self.tensors.items() if index in t.indices]

    def output_tensors(self) -> List[str]:
        return list(self.output_tensors_)

    # -- construction ----------------------------------------------------
    def push(self, qubits: Sequence[int], matrix: np.ndarray, decompose: bool = True) -> List[str]:
        """``push!(tnc, qubits, matrix; decompose)`` (tn_conversion.jl:15, basics.md:20-40).
        Qubits are 1-based."""
        if self.output_tensors_:
            raise ValueError("cannot add gates after outputs were attached")
        U = np.asarray(matrix, dtype=np.complex128)
        qs = [int(q) - 1 for q in qubits]
        if len(qs) == 1:
            q = qs[0]
            i_in = self.wire[q]
            i_out = self._new_index(2)
            self.wire[q] = i_out
            # modes (out, in); if diagonal the out index folds back into the wire
            return [self._add_tensor([i_out, i_in], U.reshape(2, 2))[0]]
        if len(qs) != 2:
            raise ValueError("only 1- and 2-qubit gates are supported")
        q1, q2 = qs
        if q1 == q2:
            raise ValueError("two-qubit gate on a single qubit")
        T4 = U.reshape(2, 2, 2, 2).transpose(0, 2, 1, 3)   # [o1,i1,o2,i2]
        in1, in2 = self.wire[q1], self.wire[q2]
        out1, out2 = self._new_index(2), self._new_index(2)
        self.wire[q1], self.wire[q2] = out1, out2
        if not decompose:
            return [self._add_tensor([out1, in1, out2, in2], T4)[0]]
        A, B = _operator_schmidt(T4)
        bond = self._new_index(A.shape[2])
        s1, ren = self._add_tensor([out1, in1, bond], A)
        bond = ren.get(bond, bond)        # the bond may have been folded into wire 1
        s2, _ = self._add_tensor([out2, in2, bond], B)
        return [s1, s2]

    def _state_tensors(self, spec: Optional[str], which: str) -> None:
        vecs = {"0": [1, 0], "1": [0, 1], "+": [1, 1], "-": [1, -1]}
        spec = "0" * self.qubits if spec is None else spec
        if len(spec) != self.qubits or any(ch not in vecs for ch in spec):
            raise ValueError(f"bad {which} specification {spec!r}")
        existing = self.input_tensors if which == "input" else self.output_tensors_
        if existing:
            # replace data only: symbols stay valid for existing plans (simulation.jl:87-88)
            for sym, ch in zip(existing, spec):
                self.tensors[sym].data = np.asarray(vecs[ch], dtype=np.complex128)
            return
        for q, ch in enumerate(spec):
            ix = self.wire_first[q] if which == "input" else self.wire[q]
            sym = self.next_tensor_id()
            self.tensors[sym] = TNTensor([ix], np.asarray(vecs[ch], dtype=np.complex128))
            existing.append(sym)

    def add_input(self, spec: Optional[str] = None) -> None:
        self._state_tensors(spec, "input")

    def add_output(self, spec: Optional[str] = None) -> None:
        self._state_tensors(spec, "output")


def convert_to_tnc(circ: Circuit, input: Optional[str] = None, output: Optional[str] = None,
                   no_input: bool = False, no_output: bool = False,
                   decompose: bool = True) -> TensorNetworkCircuit:
    """tn_conversion.jl:30-40."""
    tnc = TensorNetworkCircuit(circ.num_qubits)
    for g in circ.gates:
        tnc.push(gate_qubits(g), gate_matrix(g), decompose=decompose)
    if not no_input:
        tnc.add_input(input)
    if not no_output:
        tnc.add_output(output)
    return tnc
