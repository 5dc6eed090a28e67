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


class _FakeSpectral:
    """Stands in for OptWBoundEignVal in rho_test: comp_rho is a deterministic function of the minibatch, so the test
    sees only the replicas-only bookkeeping (who takes which minibatch, gathering, weighting) -- no device work."""

    def __init__(self):
        self.rho = self.norm = 0.0
        self.calls = []

    def comp_rho(self, data):
        x, y = data
        self.calls.append(int(x[0, 0]))
        self.rho = float(x.sum())
        self.norm = float(x.abs().max())
        return int(x[0, 0]) % 5 + 1, 0.25 * float(x[0, 0]), len(y)


def _loader():
    # ragged: seven minibatches, the last one smaller (opt.py:882-910 weighs by batch size)
    return [(torch.full((4 if j < 6 else 3, 2), float(j)), torch.zeros(4 if j < 6 else 3)) for j in range(7)]


def _rho_sweep_worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from optwboundeigenval_b200 import hvp_operator, spectral
        me = _FakeSpectral()
        stats, avg = spectral.rho_test(me, _loader())
        ret[rank] = (me.calls, stats[:, :5].tolist(), avg[:4].tolist(), hvp_operator._DATA_PARALLEL)
    finally:
        dist.destroy_process_group()


def test_rho_sweep_is_replicas_only_under_gloo():
    """spectral.rho_test with two ranks: rank r takes minibatches j % 2 == r, nobody calls a collective inside the
    sweep, every rank ends with the rows of ALL minibatches in loader order and the size-weighted averages of the
    one-process sweep; the data-parallel mode is restored afterwards."""
    from optwboundeigenval_b200 import spectral
    single = _FakeSpectral()
    stats1, avg1 = spectral.rho_test(single, _loader())
    assert single.calls == list(range(7))
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_rho_sweep_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    out = dict(ret)
    assert out[0][0] == [0, 2, 4, 6] and out[1][0] == [1, 3, 5]
    for r in range(world):
        assert out[r][1] == stats1[:, :5].tolist()
        assert out[r][2] == pytest.approx(avg1[:4].tolist(), rel=1e-12)
        assert out[r][3] is True
