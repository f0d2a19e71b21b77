"""numpy restatement of the QXTools/QXContexts contraction hot path (the ORACLE).

TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``.

What it restates
----------------
The reference executes a ``.qx`` DSL program once per (output bitstring, slice
assignment) and sums the resulting scalars over the slices:

* the loop itself lives in QXContexts ``execute`` (call site
  ``/root/reference/bin/qxrun.jl:83-87``) and, for the unsliced in-process path, in
  QXTns ``contract_tn!`` (call site ``/root/reference/src/simulation.jl:86-91``).
  Both are third-party, un-vendored Julia packages (``Project.toml:33-35``:
  ``QXContexts = "1.0"``, ``QXTns = "1.0"``; no Manifest, so no exact pin) and there is
  no Julia toolchain in this image, so the reference itself cannot run here.
* the *semantics* of every instruction are normative in the reference repo and
  that is what this file follows, line by line:
    - DSL grammar and meaning: ``docs/src/users_guide.md:93-164``
      (load :102-105, view :107-122, ncon :124-146, output :149-158, save :161-164)
    - which instructions can appear and how views chain:
      ``src/compute_graph/compute_graph.jl:15-98``
    - column-major tensor data, ComplexF64 on disk:
      ``src/compute_graph/tensor_cache.jl:51-81``
    - bitstring character i <-> qubit i <-> ``output`` index i:
      ``docs/src/basics.md:55``
    - ``-a`` / ``-n`` keep the FIRST amplitudes / slices: ``bin/qxrun.jl:32-39``

Parity pin
----------
Checked (tests/test_oracle.py) against every known answer the reference's own
tests and docs hold for this path: the complete sliced 2-qubit GHZ program of
``docs/src/users_guide.md:71-90`` (KAT-0), GHZ-3 "000"/"111"/"100"
(``README.md:38-50``), the GHZ-3 state vector
(``test/test_contraction_planning.jl:58-61``) and GHZ-5 over all 32 bitstrings
(``test/test_simulation.jl:16-26``).  The reference has NO test that executes a
sliced graph, ComplexF32 or any RQC amplitude -- for those, parity is
"unpinned by the reference" and rests on self-consistency (sliced == unsliced,
norm = 1, analytic QFT moduli).

Conventions fixed here (the reference leaves them to the executor)
------------------------------------------------------------------
* slice enumeration: linear slice id ``s in [0, prod d_i)``; ``v_i = (s // prod_{j<i} d_j) % d_i``
  (``v1`` fastest, 0-based internally, 1-based in Julia).
* ``ncon`` is evaluated as transpose -> batched matmul -> transpose with numpy
  (OpenBLAS) in the requested complex dtype.
"""
from __future__ import annotations

import dataclasses
from typing import Dict, List, Sequence, Tuple

import numpy as np

__all__ = [
    "parse_dsl", "slice_dims", "slice_values", "contract_instance",
    "amplitude", "amplitudes", "ncon_pair",
]


@dataclasses.dataclass
class Cmd:
    op: str
    args: tuple


def _labels(tok: str) -> List[int]:
    # "0" is the scalar placeholder (users_guide.md:138-144)
    if tok == "0":
        return []
    return [int(x) for x in tok.split(",")]


def parse_dsl(text: str) -> List[Cmd]:
    """Parse a ``.qx`` program (users_guide.md:93-164).  Comments are ignored
    except that line 1 must carry a version string (users_guide.md:95-99)."""
    cmds: List[Cmd] = []
    lines = text.splitlines()
    if not lines or not lines[0].startswith("# version:"):
        raise ValueError("first line of a .qx file must be '# version: x.y.z'")
    for ln in lines:
        ln = ln.strip()
        if not ln or ln.startswith("#"):
            continue
        t = ln.split()
        op = t[0]
        if op == "load":
            cmds.append(Cmd("load", (t[1], t[2], tuple(int(x) for x in t[3].split(",")))))
        elif op == "output":
            cmds.append(Cmd("output", (t[1], int(t[2]), int(t[3]))))
        elif op == "view":
            cmds.append(Cmd("view", (t[1], t[2], t[3], int(t[4]), int(t[5]))))
        elif op == "ncon":
            cmds.append(Cmd("ncon", (t[1], _labels(t[2]), t[3], _labels(t[4]), t[5], _labels(t[6]))))
        elif op == "save":
            cmds.append(Cmd("save", (t[1], t[2])))
        else:
            raise ValueError(f"unknown DSL instruction {op!r}")
    return cmds


def slice_dims(cmds: Sequence[Cmd]) -> List[Tuple[str, int]]:
    """Slice symbols in numeric order v1..vk with their extents
    (compute_graph.jl:42-43,49: ``Symbol("v$i")`` and ``dim(bg[1])``)."""
    dims: Dict[str, int] = {}
    for c in cmds:
        if c.op == "view":
            sym, d = c.args[2], c.args[4]
            if dims.setdefault(sym, d) != d:
                raise ValueError(f"inconsistent extent for slice symbol {sym}")
    return sorted(dims.items(), key=lambda kv: int(kv[0][1:]))


def slice_values(s: int, dims: Sequence[Tuple[str, int]]) -> Dict[str, int]:
    """Linear slice id -> {symbol: 0-based value}; v1 is the fastest digit."""
    out = {}
    for sym, d in dims:
        out[sym] = s % d
        s //= d
    return out


def ncon_pair(A, a_l, B, b_l, c_l):
    """Pairwise einsum with integer labels (users_guide.md:124-146).
    label in A and B and not in C -> summed; in A, B and C -> batch (:146)."""
    a_l, b_l, c_l = list(a_l), list(b_l), list(c_l)
    if A.ndim != len(a_l) or B.ndim != len(b_l):
        raise ValueError("label count does not match tensor rank")
    sa, sb, sc = set(a_l), set(b_l), set(c_l)
    ext = {}
    for l, d in list(zip(a_l, A.shape)) + list(zip(b_l, B.shape)):
        if ext.setdefault(l, d) != d:
            raise ValueError(f"extent mismatch on label {l}")
    bat = [l for l in c_l if l in sa and l in sb]
    ks = [l for l in a_l if l in sb and l not in sc]
    ms = [l for l in c_l if l in sa and l not in sb]
    ns = [l for l in c_l if l in sb and l not in sa]
    # labels private to one operand and absent from C are summed out first
    for l in [l for l in a_l if l not in sb and l not in sc]:
        ax = a_l.index(l); A = A.sum(axis=ax); a_l.pop(ax)
    for l in [l for l in b_l if l not in sa and l not in sc]:
        ax = b_l.index(l); B = B.sum(axis=ax); b_l.pop(ax)
    if set(bat + ms + ns) != sc or len(c_l) != len(sc):
        raise ValueError("output labels must be a duplicate-free subset of the input labels")
    pr = lambda ls: int(np.prod([ext[l] for l in ls], dtype=np.int64)) if ls else 1
    At = np.transpose(A, [a_l.index(l) for l in bat + ms + ks]).reshape(pr(bat), pr(ms), pr(ks))
    Bt = np.transpose(B, [b_l.index(l) for l in bat + ks + ns]).reshape(pr(bat), pr(ks), pr(ns))
    Ct = np.matmul(At, Bt).reshape([ext[l] for l in bat + ms + ns])
    order = bat + ms + ns
    return np.transpose(Ct, [order.index(l) for l in c_l])


def contract_instance(cmds: Sequence[Cmd], data: Dict[str, np.ndarray], bits: str,
                      svals: Dict[str, int], dtype=np.complex128, track_sliced: bool = False):
    """Run the whole program for ONE (bitstring, slice assignment) and return the
    ``save``d tensor (a scalar for closed networks)."""
    env: Dict[str, np.ndarray] = {}
    sliced: Dict[str, set] = {}          # track_sliced: per tensor, the positions of modes a view has fixed
    result = None
    for c in cmds:
        if c.op == "load":
            name, label, dims = c.args
            arr = np.asarray(data[label])
            # tensor_cache.jl:52-53,80 -- data is the column-major flattening
            env[name] = np.reshape(arr.reshape(-1, order="F"), dims, order="F").astype(dtype)
        elif c.op == "output":
            name, idx, dim = c.args
            v = np.zeros(dim, dtype=dtype)
            ch = bits[idx - 1]          # basics.md:55: char i <-> qubit i
            if ch in "01":
                v[int(ch)] = 1
            elif ch == "+":            # basics.md:62 (un-normalised "0"+"1")
                v[:2] = 1
            elif ch == "-":
                v[0], v[1] = 1, -1
            else:
                raise ValueError(f"bad bitstring character {ch!r}")
            env[name] = v
        elif c.op == "view":
            name, target, sym, pos, _dim = c.args
            # rank is preserved: the sliced mode keeps extent 1 (users_guide.md:85-88)
            env[name] = np.take(env[target], [svals[sym]], axis=pos - 1)
            sliced[name] = sliced.get(target, set()) | {pos - 1}
        elif c.op == "ncon":
            out, c_l, a, a_l, b, b_l = c.args
            env[out] = ncon_pair(env[a], a_l, env[b], b_l, c_l)
            if track_sliced:
                lab = {a_l[i] for i in sliced.get(a, ())} | {b_l[i] for i in sliced.get(b, ())}
                sliced[out] = {i for i, l in enumerate(c_l) if l in lab}
        elif c.op == "save":
            result = env[c.args[1]]
            if track_sliced and sliced.get(c.args[1]):
                # the slices would be different ELEMENTS of the result, not terms of a sum
                raise ValueError("a mode of the saved tensor is sliced: open indices cannot be slice bonds")
    if result is None:
        raise ValueError("program has no save instruction")
    return result


def amplitude(cmds, data, bits: str, dtype=np.complex128, slice_begin=0, slice_end=None):
    """amplitude(b) = sum over slices of the saved scalar (users_guide.md:14-20)."""
    dims = slice_dims(cmds)
    total = int(np.prod([d for _, d in dims], dtype=np.int64)) if dims else 1
    if slice_end is None:
        slice_end = total
    acc = dtype(0)
    for s in range(slice_begin, slice_end):
        r = contract_instance(cmds, data, bits, slice_values(s, dims), dtype)
        acc = acc + np.asarray(r).reshape(-1)[0]
    return acc


def contract(cmds, data, bits: str = "", dtype=np.complex128, slice_begin=0, slice_end=None) -> np.ndarray:
    """The saved TENSOR summed over slices: what ``contract_tn!`` returns for an open network
    (/root/reference/test/test_contraction_planning.jl:58-61: ``reshape(output, 8)`` is the GHZ state vector).
    For a closed network this is the amplitude as a 0-d array."""
    dims = slice_dims(cmds)
    total = int(np.prod([d for _, d in dims], dtype=np.int64)) if dims else 1
    if slice_end is None:
        slice_end = total
    acc = None
    for s in range(slice_begin, slice_end):
        r = np.asarray(contract_instance(cmds, data, bits, slice_values(s, dims), dtype, track_sliced=True))
        acc = r.copy() if acc is None else acc + r
    return acc


def amplitudes(cmds, data, bitstrings: Sequence[str], dtype=np.complex128,
               slice_begin=0, slice_end=None) -> np.ndarray:
    return np.array([amplitude(cmds, data, b, dtype, slice_begin, slice_end) for b in bitstrings],
                    dtype=dtype)
