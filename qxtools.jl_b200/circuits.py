"""Circuit constructors feeding the contraction hot path.

Host-side mirror of ``/root/reference/src/circuits/circuits.jl`` (GHZ :30-37,
QFT :44-50, grid RQC :61-67, ``gate_matrix`` :74-76).  The reference delegates to
QXZoo (not vendored, no Julia here); these are independent restatements that only
need to produce the same *kind* of network: they are input generators for the
executor, not part of the hot path.  Qubits are numbered 1..N as in
``docs/src/basics.md:24``.

Two-qubit matrices use the usual kron convention: ``U[(o1 o2), (i1 i2)]`` with
``qubits = (q1, q2)`` and q1 the most significant factor.
"""
from __future__ import annotations

import dataclasses
from typing import List, Optional, Sequence, Tuple

import numpy as np

_S2 = 1.0 / np.sqrt(2.0)

H = np.array([[1, 1], [1, -1]], dtype=np.complex128) * _S2
X = np.array([[0, 1], [1, 0]], dtype=np.complex128)
Y = np.array([[0, -1j], [1j, 0]], dtype=np.complex128)
Z = np.array([[1, 0], [0, -1]], dtype=np.complex128)
T = np.diag([1.0, np.exp(1j * np.pi / 4)]).astype(np.complex128)
SX = 0.5 * np.array([[1 + 1j, 1 - 1j], [1 - 1j, 1 + 1j]], dtype=np.complex128)
SY = 0.5 * np.array([[1 + 1j, -1 - 1j], [1 + 1j, 1 + 1j]], dtype=np.complex128)
# sqrt(W), W = (X+Y)/sqrt(2)  (Sycamore single-qubit gate set)
SW = np.array([[1 + 1j, -np.sqrt(2) * 1j], [np.sqrt(2), 1 + 1j]], dtype=np.complex128) * 0.5
CX = np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 0, 1], [0, 0, 1, 0]], dtype=np.complex128)
CZ = np.diag([1, 1, 1, -1]).astype(np.complex128)


def cphase(theta: float) -> np.ndarray:
    return np.diag([1, 1, 1, np.exp(1j * theta)]).astype(np.complex128)


def fsim(theta: float, phi: float) -> np.ndarray:
    c, s = np.cos(theta), np.sin(theta)
    return np.array([[1, 0, 0, 0], [0, c, -1j * s, 0], [0, -1j * s, c, 0],
                     [0, 0, 0, np.exp(-1j * phi)]], dtype=np.complex128)


@dataclasses.dataclass
class Gate:
    name: str
    qubits: Tuple[int, ...]
    matrix: np.ndarray


@dataclasses.dataclass
class Circuit:
    num_qubits: int
    gates: List[Gate] = dataclasses.field(default_factory=list)

    def add(self, name: str, qubits: Sequence[int], matrix: np.ndarray) -> "Circuit":
        qubits = tuple(int(q) for q in qubits)
        if any(q < 1 or q > self.num_qubits for q in qubits):
            raise ValueError(f"qubit out of range in {name}{qubits}")
        self.gates.append(Gate(name, qubits, np.asarray(matrix, dtype=np.complex128)))
        return self


def gate_matrix(gate: Gate) -> np.ndarray:
    """circuits.jl:74-76 -- always ComplexF64."""
    return np.asarray(gate.matrix, dtype=np.complex128)


def gate_qubits(gate: Gate) -> Tuple[int, ...]:
    return gate.qubits


def create_test_circuit() -> Circuit:
    """3-qubit GHZ preparation (circuits.jl:16-22)."""
    return create_ghz_circuit(3)


def create_ghz_circuit(n: int) -> Circuit:
    """H on qubit 1 then a CX chain (circuits.jl:30-37)."""
    c = Circuit(n)
    c.add("h", (1,), H)
    for i in range(1, n):
        c.add("cx", (i, i + 1), CX)     # (control, target)
    return c


def create_qft_circuit(n: int) -> Circuit:
    """QFT on n qubits followed by the two CX of circuits.jl:44-50.

    Standard textbook QFT without the final swaps (H then controlled phases
    2*pi/2^k).  QXZoo's exact gate order is not available here; the amplitudes of
    QFT|0..0> all have modulus 2^{-n/2} either way, which is what the tests pin."""
    c = Circuit(n)
    for j in range(1, n + 1):
        c.add("h", (j,), H)
        for k in range(j + 1, n + 1):
            c.add("cphase", (k, j), cphase(2 * np.pi / 2 ** (k - j + 1)))
    if n >= 2:
        c.add("cx", (1, 2), CX)
    if n >= 3:
        c.add("cx", (2, 3), CX)
    return c


def _grid_cz_patterns(rows: int, cols: int):
    """Eight disjoint CZ layouts covering every grid edge once per 8 cycles
    (Boixo et al. 2018 style)."""
    q = lambda r, c: r * cols + c + 1
    hor = {(a, b): [] for a in range(2) for b in range(2)}
    ver = {(a, b): [] for a in range(2) for b in range(2)}
    for r in range(rows):
        for c in range(cols - 1):
            hor[(c % 2, r % 2)].append((q(r, c), q(r, c + 1)))
    for r in range(rows - 1):
        for c in range(cols):
            ver[(r % 2, c % 2)].append((q(r, c), q(r + 1, c)))
    return [hor[(0, 0)], ver[(0, 0)], hor[(1, 1)], ver[(1, 1)],
            hor[(0, 1)], ver[(0, 1)], hor[(1, 0)], ver[(1, 0)]]


def create_rqc_circuit(rows: int, cols: int, depth: int, seed: Optional[int] = None,
                       final_h: bool = False) -> Circuit:
    """Grid random quantum circuit (circuits.jl:61-67 -> QXZoo.RQC.create_RQC).

    Rules (Boixo et al., "Characterizing quantum supremacy in near-term devices"):
    cycle 0 is a Hadamard layer; each later cycle applies one of 8 CZ layouts; a
    single-qubit gate is placed on a qubit that is idle in this cycle and was in a
    CZ in the previous one: T if it is the qubit's first such gate, otherwise a
    random choice of sqrt(X)/sqrt(Y) different from the qubit's previous one.
    RNG is numpy ``default_rng(seed)`` (Julia's MersenneTwister stream cannot be
    reproduced here), so circuits are seed-stable within this repo only."""
    n = rows * cols
    rng = np.random.default_rng(seed)
    c = Circuit(n)
    for qb in range(1, n + 1):
        c.add("h", (qb,), H)
    patterns = _grid_cz_patterns(rows, cols)
    prev_cz = set()
    last_gate = {qb: None for qb in range(1, n + 1)}
    for t in range(1, depth + 1):
        layer = patterns[(t - 1) % 8]
        busy = set()
        for a, b in layer:
            busy.add(a); busy.add(b)
        for qb in range(1, n + 1):
            if qb in busy or qb not in prev_cz:
                continue
            if last_gate[qb] is None:
                g = "t"
            else:
                opts = [o for o in ("sx", "sy") if o != last_gate[qb]]
                g = opts[int(rng.integers(len(opts)))]
            last_gate[qb] = g
            c.add(g, (qb,), {"t": T, "sx": SX, "sy": SY}[g])
        for a, b in layer:
            c.add("cz", (a, b), CZ)
        prev_cz = busy
    if final_h:
        for qb in range(1, n + 1):
            c.add("h", (qb,), H)
    return c


def create_sycamore_like_circuit(cycles: int, seed: Optional[int] = None, n_qubits: int = 53,
                                 theta: float = np.pi / 2, phi: float = np.pi / 6) -> Circuit:
    """Sycamore-like circuit on a diagonal lattice: per cycle one random
    sqrt(X)/sqrt(Y)/sqrt(W) per qubit (never repeating) then fSim gates on one of
    the coupler classes in the ABCDCDAB sequence.  BASELINE.json config 5 -- NOT
    producible by the reference (circuits.jl only builds CZ grids)."""
    rng = np.random.default_rng(seed)
    # 54-site diagonal lattice (rows of 6 on a 9-row brick pattern), drop sites to reach n_qubits
    sites = []
    for r in range(9):
        for k in range(6):
            sites.append((r, 2 * k + (r % 2)))
    sites = sites[:n_qubits] if n_qubits <= len(sites) else sites
    n = len(sites)
    idx = {s: i + 1 for i, s in enumerate(sites)}
    classes = {"A": [], "B": [], "C": [], "D": []}
    for (r, x) in sites:
        for dx, kind in ((1, 0), (-1, 1)):
            nb = (r + 1, x + dx)
            if nb in idx:
                if kind == 0:
                    cls = "A" if r % 2 == 0 else "B"
                else:
                    cls = "C" if r % 2 == 0 else "D"
                classes[cls].append((idx[(r, x)], idx[nb]))
    seq = "ABCDCDAB"
    c = Circuit(n)
    last = {qb: None for qb in range(1, n + 1)}
    mats = {"sx": SX, "sy": SY, "sw": SW}
    U = fsim(theta, phi)
    for t in range(cycles):
        for qb in range(1, n + 1):
            opts = [o for o in ("sx", "sy", "sw") if o != last[qb]]
            g = opts[int(rng.integers(len(opts)))]
            last[qb] = g
            c.add(g, (qb,), mats[g])
        for a, b in classes[seq[t % 8]]:
            c.add("fsim", (a, b), U)
    for qb in range(1, n + 1):
        opts = [o for o in ("sx", "sy", "sw") if o != last[qb]]
        g = opts[int(rng.integers(len(opts)))]
        c.add(g, (qb,), mats[g])
    return c
