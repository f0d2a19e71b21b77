"""Generates tests/golden/golden_rqc.json (+ the .qx/.npz triples it refers to).

The reference ships no golden files (its tests write into mktempdir,
test/test_bin.jl:12-18) and cannot run here (no Julia), so these vectors are
produced by the pinned oracle on file triples emitted by this repo's host mirror:
they freeze the (program, tensors, bitstrings) -> amplitudes map so that later
changes to the planner or the oracle cannot silently move the target.
Run from the repo root:  python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import qx_oracle as orc   # noqa: E402
from cases import rqc_case            # noqa: E402

cases = []
for name, (r, c, d, ns, namp) in {"rqc_3x3_d8_s2": (3, 3, 8, 2, 6), "rqc_4x4_d12_s4": (4, 4, 12, 4, 6),
                                  "rqc_4x5_d14_s5": (4, 5, 14, 5, 4)}.items():
    txt, data, bs = rqc_case(r, c, d, ns, seed=42, n_amp=namp, amp_seed=11)
    open(os.path.join(HERE, name + ".qx"), "w").write(txt)
    np.savez(os.path.join(HERE, name + ".npz"), **data)
    cmds = orc.parse_dsl(txt)
    amps = orc.amplitudes(cmds, data, bs)
    per_slice = [complex(orc.amplitude(cmds, data, bs[0], slice_begin=s, slice_end=s + 1)) for s in range(2 ** ns)]
    cases.append({"name": name, "qx": name + ".qx", "npz": name + ".npz", "n_qubits": r * c, "bitstrings": bs,
                  "re": [float(a.real) for a in amps], "im": [float(a.imag) for a in amps],
                  "slice_re_bs0": [a.real for a in per_slice], "slice_im_bs0": [a.imag for a in per_slice]})
json.dump({"generator": "tests/golden/make_golden.py", "dtype": "complex128", "cases": cases},
          open(os.path.join(HERE, "golden_rqc.json"), "w"), indent=1)
print("wrote", len(cases), "cases")
