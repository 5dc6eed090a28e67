"""Host-side logic of the sharded (N > 1) path under gloo, world_size 2, on the CPU: batch-global
weighted-BCE counts, shard bookkeeping of bench.py.  The device collectives themselves are checked on
GPUs by tools/mgpu_check.py (sharded vs full batch)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from optwboundeigenval_b200.hvp_operator import SpectralPlan
        g = torch.Generator().manual_seed(0)
        Y = (torch.rand(8, 14, generator=g) > 0.8).float()
        Y[1, 2] = float("nan")
        per = Y.shape[0] // world
        shard = Y[rank * per:(rank + 1) * per]

        class _P:      # only the method under test needs an instance
            pass
        sums = lambda p, s, n_c: SpectralPlan._global_sums(_P(), p, s, n_c)   # noqa: E731
        t, coef = SpectralPlan.wbce_coefficients(shard, sums)
        t_full, coef_full = SpectralPlan.wbce_coefficients(Y)
        ok = torch.allclose(coef, coef_full[rank * per:(rank + 1) * per], rtol=1e-6, atol=0) and \
            torch.equal(t, t_full[rank * per:(rank + 1) * per])
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_weighted_bce_counts_are_batch_global_under_gloo():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    assert dict(ret) == {0: True, 1: True}
