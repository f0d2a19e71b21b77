"""Random small tensor-network programs: the lowered program (CPU emulator) and, on the GPU
box, the library itself must reproduce the oracle on every one of them."""
import numpy as np
import pytest

from qxb200.executor import Graph, bits_from_strings
from oracle import qx_oracle as orc
import lowered_emulator as em
from cases import random_program

SEEDS = list(range(24))


@pytest.mark.parametrize("seed", SEEDS)
def test_random_program_lowering(lib_built, seed):
    txt, data, bs = random_program(seed)
    cmds = orc.parse_dsl(txt)
    ref = orc.amplitudes(cmds, data, bs)
    g = Graph.from_dsl(txt, data)
    n_out = g.n_outputs
    bits = bits_from_strings(bs, n_out)
    assert np.allclose(em.amplitudes(g, data, bits, shuffle_seed=seed), ref, atol=1e-10 * max(1.0, np.max(np.abs(ref))))
    S = g.n_slices
    if S > 2:
        ref2 = orc.amplitudes(cmds, data, bs, slice_begin=1, slice_end=S - 1)
        assert np.allclose(em.amplitudes(g, data, bits, 1, S - 1), ref2, atol=1e-10 * max(1.0, np.max(np.abs(ref))))
    # in-library re-planning must not change the value either
    g.replan(6, 8)
    new = g.program_text()
    assert np.allclose(orc.amplitudes(orc.parse_dsl(new), data, bs), ref, atol=1e-10 * max(1.0, np.max(np.abs(ref))))


@pytest.mark.gpu
@pytest.mark.parametrize("dtype,tol", [("c64", 1e-10), ("c32", 2e-4)])
def test_random_programs_on_gpu(gpu, dtype, tol):
    for seed in SEEDS:
        txt, data, bs = random_program(seed)
        cmds = orc.parse_dsl(txt)
        ref = orc.amplitudes(cmds, data, bs)
        scale = max(1e-30, float(np.max(np.abs(ref))))
        for replan in (0, 6):
            g = Graph.from_dsl(txt, data, dtype, replan=replan, replan_n_amp=8).compile()
            out = g.amplitudes(bs)
            assert np.max(np.abs(out - ref)) / scale < tol, (seed, replan)
            S = g.n_slices
            if S > 2:
                ref2 = orc.amplitudes(cmds, data, bs, slice_begin=1, slice_end=S - 1)
                assert np.max(np.abs(g.amplitudes(bs, 1, S - 1) - ref2)) / scale < tol, (seed, replan)
