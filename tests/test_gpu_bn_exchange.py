"""The per-channel packet exchange of the BatchNorm sums (peer.cuh: peer_exchange_channel) on ONE GPU.

Under data parallelism the fused BatchNorm kernels exchange each channel's sums with the peers as 16-byte tagged
packets and use the local packet as the barrier between the channel's blocks (no grid.sync).  With world = 1
(``B2S_BN_CHANSYNC=1``) the same code path runs against a local buffer, so the single-GPU test box exercises the ticket /
packet / polling logic that ``tests/test_gpu_multi.py`` checks across GPUs: gradient, Hv and vGHv must equal the
grid-barrier form (the sums are the same fp64 atomics; only who waits for whom changes)."""
import os

import numpy as np
import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("kind,batch", [("cifar_densenet", 8), ("cifar_densenet", 32)])
def test_per_channel_barrier_equals_grid_barrier(kind, batch):
    from optwboundeigenval_b200 import zoo
    from optwboundeigenval_b200.hvp_operator import B200HVPOperator, clear_plans
    model, loss = zoo.build(kind)
    model.train()
    x, y = zoo.synthetic_batch(kind, batch)
    P = sum(p.numel() for p in model.parameters())
    g = torch.Generator().manual_seed(11)
    v = torch.randn(P, generator=g, dtype=torch.float64)
    v /= v.norm()
    res = {}
    old = os.environ.get("B2S_BN_CHANSYNC")
    try:
        for mode in ("0", "1"):
            os.environ["B2S_BN_CHANSYNC"] = mode          # read when the plan is created
            clear_plans()
            op = B200HVPOperator(model, [x, y], loss)
            grad = op.prepare_grad().cpu().numpy().copy()
            hv = op.Hv(v, storedGrad=True).cpu().numpy().copy()
            vg = op.vGHv(v, storedGrad=True).cpu().numpy().copy()
            # a second pass over the same plan: tickets and the sequence counter must have been left consistent
            hv2 = op.Hv(v, storedGrad=True).cpu().numpy().copy()
            assert rel_err(hv2, hv) < 1e-6
            res[mode] = (grad, hv, vg)
    finally:
        if old is None:
            os.environ.pop("B2S_BN_CHANSYNC", None)
        else:
            os.environ["B2S_BN_CHANSYNC"] = old
        clear_plans()
    for a, b, what in zip(res["1"], res["0"], ("grad", "Hv", "vGHv")):
        assert np.all(np.isfinite(a)), what
        # fp64 atomics in a different arrival order and fp32 atomics of the weight gradients: rounding-level differences
        assert rel_err(a, b) < 2e-5, (what, rel_err(a, b))
