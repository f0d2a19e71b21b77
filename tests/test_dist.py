"""N > 1 path on CPU: world_size-2 gloo processes exercise the slice/bitstring
partition and the single reduction, with the oracle standing in for the per-rank
contraction (the GPU version of the same flow is bench.py --gpus N / execute(use_mpi))."""
import os
import sys

import numpy as np
import pytest

from qxb200.dist import Distribution, partition_range

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_partition_range_covers():
    for n in (0, 1, 7, 64, 4096):
        for parts in (1, 2, 3, 8):
            got = []
            for i in range(parts):
                b, e = partition_range(n, parts, i)
                assert b <= e
                got.extend(range(b, e))
            assert got == list(range(n))


def test_distribution_two_level():
    # 8 ranks, sub-communicators of 4: 2 bitstring groups x 4 slice shares
    seen = set()
    for r in range(8):
        d = Distribution(10, 4096, world=8, rank=r, sub_comm_size=4)
        assert d.n_groups == 2 and d.group == r // 4 and d.rank_in_group == r % 4
        assert d.slice_end - d.slice_begin == 1024
        for a in range(d.amp_begin, d.amp_end):
            for s in (d.slice_begin, d.slice_end - 1):
                assert (a, s) not in seen
                seen.add((a, s))
    with pytest.raises(ValueError):
        Distribution(10, 64, world=8, rank=0, sub_comm_size=3)


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as td
    from oracle import qx_oracle as orc
    from cases import rqc_case
    from qxb200.dist import Distribution, reduce_partial_amplitudes
    td.init_process_group("gloo", rank=rank, world_size=world)
    txt, data, bs = rqc_case(3, 3, 8, 3, n_amp=4)
    cmds = orc.parse_dsl(txt)
    d = Distribution(len(bs), 8, world, rank)
    part = orc.amplitudes(cmds, data, bs, slice_begin=d.slice_begin, slice_end=d.slice_end)
    t = torch.from_numpy(np.ascontiguousarray(part))
    reduce_partial_amplitudes(t)
    if rank == 0:
        q.put(t.numpy())
    td.destroy_process_group()


def test_gloo_two_ranks_sum_slices():
    import torch.multiprocessing as mp
    from oracle import qx_oracle as orc
    from cases import rqc_case
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29650 + os.getpid() % 200
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    txt, data, bs = rqc_case(3, 3, 8, 3, n_amp=4)
    ref = orc.amplitudes(orc.parse_dsl(txt), data, bs)
    assert np.allclose(got, ref, atol=1e-14)
