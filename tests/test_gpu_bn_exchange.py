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


@pytest.mark.parametrize("kind,batch", [("cifar_densenet", 32), ("cifar_densenet", 16)])
def test_clustered_cached_form_equals_cooperative_form(kind, batch):
    """bn.cu: bn_fwd_clus_kernel / bn_bwd_clus_kernel (a thread-block cluster of 2 / 4 / 8 CTAs per channel, partial sums
    exchanged through distributed shared memory) against the cooperative grid-barrier kernels (``B2S_BN_CLUSTER=0``) on
    the layers whose channel does not fit one CTA's cache (DenseNet3's first dense block at batch 32 / 16)."""
    from optwboundeigenval_b200 import zoo
    from optwboundeigenval_b200.hvp_operator import B200HVPOperator, clear_plans
    model, loss = zoo.build(kind)
    model.train()
    x, y = zoo.synthetic_batch(kind, batch)
    P = sum(p.numel() for p in model.parameters())
    g = torch.Generator().manual_seed(12)
    v = torch.randn(P, generator=g, dtype=torch.float64)
    v /= v.norm()
    res = {}
    old = os.environ.get("B2S_BN_CLUSTER")
    buffers = [b.clone() for b in model.buffers()]
    try:
        for mode in ("0", "1"):
            os.environ["B2S_BN_CLUSTER"] = mode           # read at every launch, i.e. when the plan's graphs are captured
            clear_plans()
            with torch.no_grad():
                for b, saved in zip(model.buffers(), buffers):
                    b.copy_(saved)
            op = B200HVPOperator(model, [x, y], loss)
            grad = op.prepare_grad().cpu().numpy().copy()
            hv = op.Hv(v, storedGrad=True).cpu().numpy().copy()
            vg = op.vGHv(v, storedGrad=True).cpu().numpy().copy()
            hv2 = op.Hv(v, storedGrad=True).cpu().numpy().copy()
            assert rel_err(hv2, hv) < 1e-6
            stats = np.concatenate([b.detach().double().cpu().numpy().reshape(-1) for b in model.buffers()])
            res[mode] = (grad, hv, vg, stats)
    finally:
        if old is None:
            os.environ.pop("B2S_BN_CLUSTER", None)
        else:
            os.environ["B2S_BN_CLUSTER"] = old
        clear_plans()
    for a, b, what in zip(res["1"], res["0"], ("grad", "Hv", "vGHv", "running statistics")):
        assert np.all(np.isfinite(a)), what
        assert rel_err(a, b) < 2e-5, (what, rel_err(a, b))
