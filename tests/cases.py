"""Shared workload builders for the tests (seeded; small enough for the oracle)."""
import numpy as np

import qxb200 as q


def kat0():
    """2-qubit GHZ program of the reference docs with the tensor values of SURVEY.md Appendix B."""
    import os
    here = os.path.dirname(os.path.abspath(__file__))
    txt = open(os.path.join(here, "golden", "kat0_ghz2.qx")).read()
    s = 1 / np.sqrt(2)
    T = np.zeros((2, 2, 2))
    for o in range(2):
        for i in range(2):
            for c in range(2):
                T[o, i, c] = 1.0 if o == (i ^ c) else 0.0
    data = {"data_1": np.array([1, 0.0]), "data_2": np.array([[1, 1], [1, -1]]) * s,
            "data_3": T, "data_4": np.array([1.0, 1.0])}
    return txt, data


def rqc_case(rows, cols, depth, n_slice, seed=42, n_amp=8, amp_seed=1, **kw):
    circ = q.create_rqc_circuit(rows, cols, depth, seed)
    tnc = q.convert_to_tnc(circ)
    bond_groups, plan, meta = q.contraction_scheme(tnc, n_slice, time=0, **kw)
    cg = q.build_compute_graph(tnc, plan, bond_groups)
    bitstrings = list(q.amplitudes_uniform(rows * cols, amp_seed, n_amp))
    return cg.dsl(meta), dict(cg.tensors), bitstrings


def circuit_case(circ, n_slice=0, n_amp=8, amp_seed=1, decompose=True):
    tnc = q.convert_to_tnc(circ, decompose=decompose)
    if n_slice:
        bond_groups, plan, _ = q.contraction_scheme(tnc, n_slice, time=0)
    else:
        bond_groups, plan = None, q.min_fill_contraction_plan(tnc)
    cg = q.build_compute_graph(tnc, plan, bond_groups)
    bitstrings = list(q.amplitudes_uniform(circ.num_qubits, amp_seed, n_amp))
    return cg.dsl(), dict(cg.tensors), bitstrings


def rel_err(got, ref, n_qubits):
    """Relative to max(|amp_ref|, 2^{-n/2}) (zero amplitudes need an absolute floor)."""
    scale = max(float(np.max(np.abs(ref))), 2.0 ** (-n_qubits / 2))
    return float(np.max(np.abs(np.asarray(got) - np.asarray(ref)))) / scale


def random_program(seed, n_leaves=7, n_index=9, n_out=3, n_slice=2, max_ext=3):
    """A random closed tensor network as a .qx program: hyper-edges (an index on 2-4 tensors),
    outer products, scalars, non-power-of-two extents, sliced hyper-edges with views on loads and
    outputs, contracted in a random order with labels assigned like build_compute_graph does."""
    rng = np.random.default_rng(seed)
    ext = {i: int(rng.integers(2, max_ext + 1)) for i in range(n_index)}
    # output wires: extent 2, each on one output leaf + >= 1 other tensor
    leaves = [[] for _ in range(n_leaves)]
    for i in range(n_index):
        k = int(rng.integers(2, 5))
        owners = rng.choice(n_leaves, size=min(k, n_leaves), replace=False)
        for t in owners:
            leaves[t].append(i)
    out_idx = list(rng.choice(n_index, size=n_out, replace=False))
    for i in out_idx:
        ext[i] = 2
    names, idxs, lines, data = [], [], [], {}
    for t, ids in enumerate(leaves):
        if not ids:
            ids = [int(rng.integers(n_index))]
        ids = list(dict.fromkeys(ids))
        rng.shuffle(ids)
        shape = [ext[i] for i in ids]
        data[f"d{t}"] = rng.normal(size=shape) + 1j * rng.normal(size=shape)
        lines.append(f"load t{t} d{t} {','.join(map(str, shape))}")
        names.append(f"t{t}"); idxs.append(ids)
    for q_, i in enumerate(out_idx, start=1):
        lines.append(f"output o{q_} {q_} 2")
        names.append(f"o{q_}"); idxs.append([i])
    # slice some hyper-edges: every tensor carrying the index gets a view (compute_graph.jl:39-58)
    sliced = list(rng.choice(n_index, size=min(n_slice, n_index), replace=False))
    for v, i in enumerate(sliced, start=1):
        for t in range(len(names)):
            if i in idxs[t]:
                pos = idxs[t].index(i) + 1
                lines.append(f"view {names[t]}_s {names[t]} v{v} {pos} {ext[i]}")
                names[t] = names[t] + "_s"
    # random pairwise contraction order with the hyper-edge survival rule
    alive = list(range(len(names)))
    cnt = {}
    for t in alive:
        for i in idxs[t]:
            cnt[i] = cnt.get(i, 0) + 1
    k = 0
    while len(alive) > 1:
        a, b = rng.choice(len(alive), size=2, replace=False)
        a, b = alive[int(a)], alive[int(b)]
        ia, ib = idxs[a], idxs[b]
        label = {}
        for i in ia + ib:
            label.setdefault(i, len(label) + 1)
        others = lambda i: cnt[i] - (i in ia) - (i in ib)
        keep = [i for i in ia if others(i) > 0]
        keep += [i for i in ib if i not in keep and others(i) > 0]
        lab = lambda ls: ",".join(str(label[i]) for i in ls) if ls else "0"
        k += 1
        lines.append(f"ncon I{k} {lab(keep)} {names[a]} {lab(ia)} {names[b]} {lab(ib)}")
        for i in set(ia) | set(ib):
            cnt[i] -= (i in ia) + (i in ib)
        for i in keep:
            cnt[i] += 1
        names.append(f"I{k}"); idxs.append(keep)
        alive = [t for t in alive if t not in (a, b)] + [len(names) - 1]
    lines.append(f"save output {names[alive[0]]}")
    bitstrings = ["".join("01"[b] for b in rng.integers(0, 2, n_out)) for _ in range(4)] + ["+" * n_out, "-" * n_out]
    return "# version: 0.4.0\n" + "\n".join(lines) + "\n", data, bitstrings
