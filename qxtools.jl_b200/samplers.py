"""Output samplers in front of the hot path.

The parameter file selects how output bitstrings are chosen
(/root/reference/src/outputs.jl:47-78, docs/src/features.md:68-110):

* ``List``      explicit bitstrings                              -> amplitudes
* ``Uniform``   ``num_samples`` uniform random bitstrings        -> amplitudes
* ``Rejection`` rejection sampling from the circuit's output distribution: draw a uniform
  bitstring, compute its amplitude, accept with probability ``p * 2^n / M`` (``M`` bounds
  ``p * 2^n``; with ``fix_M = False`` a larger ratio raises ``M`` -- "frugal" sampling).

The reference implements these inside QXContexts (not vendored); the acceptance rule above is
the documented one.  RNG streams differ from Julia's MersenneTwister, so samples are
reproducible within this repo only.  Candidates are evaluated in batches so that the GPU
sees one ``qxb_amplitudes`` call per batch, not one per candidate.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np

AmplitudeFn = Callable[[np.ndarray], np.ndarray]      # uint8 [n][n_qubits] -> complex [n]


def bits_to_strings(bits: np.ndarray) -> List[str]:
    return ["".join("01"[b] for b in row) for row in bits]


def uniform_bitstrings(num_qubits: int, num_samples: int, seed: Optional[int]) -> np.ndarray:
    return np.random.default_rng(seed).integers(0, 2, (num_samples, num_qubits)).astype(np.uint8)


def rejection_sample(amplitudes: AmplitudeFn, num_qubits: int, num_samples: int, M: float = 0.0001,
                     fix_M: bool = False, seed: Optional[int] = None, batch: int = 1024,
                     max_batches: int = 10_000) -> Tuple[List[str], List[complex], dict]:
    """-> (accepted bitstrings, their amplitudes, info).  ``M`` follows outputs.jl:57-62."""
    if not float(M) > 0.0:
        raise ValueError(f"rejection sampling needs M > 0 (got {M}): the acceptance probability is p 2^n / M")
    rng = np.random.default_rng(seed)
    N = 2.0 ** num_qubits
    out_b: List[str] = []
    out_a: List[complex] = []
    drawn = 0
    m = float(M)
    for _ in range(max_batches):
        if len(out_b) >= num_samples:
            break
        cand = rng.integers(0, 2, (batch, num_qubits)).astype(np.uint8)
        amps = np.asarray(amplitudes(cand))
        ratio = (np.abs(amps) ** 2) * N
        u = rng.random(batch)
        for i in range(batch):
            drawn += 1
            if not fix_M and ratio[i] > m:
                m = float(ratio[i])                      # frugal: raise the bound, keep going
            if u[i] < ratio[i] / m:
                out_b.append("".join("01"[b] for b in cand[i]))
                out_a.append(complex(amps[i]))
                if len(out_b) >= num_samples:
                    break
    return out_b, out_a, {"M": m, "drawn": drawn, "accepted": len(out_b)}
