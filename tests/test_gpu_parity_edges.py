"""Layer geometries at the edges of what the five BASELINE configs use, through the C ABI on the GPU: the same small
networks ``tests/test_jet_oracle.py::test_edge_geometries_match_autograd_fp64`` pins on the CPU (pooling with padding and
overlap, rectangular kernels and inputs, 'same' / 'valid' padding, BatchNorm / ReLU on the input, BatchNorm1d, ...)."""
import copy

import pytest
import torch

from conftest import rel_err
from test_jet_oracle import _edge_models

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", sorted(_edge_models()[0]))
def test_edge_geometries_on_the_gpu(name):
    from optwboundeigenval_b200.hvp_operator import B200HVPOperator
    from oracle import autograd_oracle as ao
    torch.manual_seed(1)
    model, shape = _edge_models()[0][name]
    model.train()
    loss = torch.nn.CrossEntropyLoss()
    g = torch.Generator().manual_seed(2)
    x = torch.randn(16, *shape, generator=g)
    y = torch.randint(0, 4, (16,), generator=g)
    ref = ao.AutogradSpectralOperator(copy.deepcopy(model).double(), [x.double(), y], loss)
    P = sum(p.numel() for p in model.parameters())
    v = torch.randn(P, generator=g, dtype=torch.float64)
    v = (v / v.norm()).float().double()
    op = B200HVPOperator(model, [x, y], loss)
    hv = op.Hv(v, storedGrad=True)
    vg = op.vGHv(v, storedGrad=True)
    # fp64 autograd as the yard-stick: the GPU computes in fp32, 1e-4 is north_star's tolerance
    assert rel_err(op.stored_grad.cpu().numpy(), ref.gradient().detach().numpy()) < 1e-4
    assert rel_err(hv.cpu().numpy(), ref.hv(v).numpy()) < 1e-4
    assert rel_err(vg.cpu().numpy(), ref.vghv(v).numpy()) < 1e-4


@pytest.mark.parametrize("seed", [0, 1, 2, 3, 4, 5])
def test_random_architectures_on_the_gpu(seed):
    """tests/test_jet_oracle.py::_RandomNet (odd convolution geometry, DenseNet-style concatenation, residual add,
    BatchNorm without ReLU, a Linear used twice, optional softmax tail) through the C ABI against fp64 autograd."""
    import numpy as np
    from optwboundeigenval_b200.hvp_operator import B200HVPOperator
    from oracle import autograd_oracle as ao
    from test_jet_oracle import _RandomNet
    rng = np.random.default_rng(100 + seed)
    torch.manual_seed(200 + seed)
    model = _RandomNet(rng).train()
    with torch.no_grad():
        for m in model.modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                m.weight.uniform_(0.5, 1.5)
                m.bias.uniform_(-0.3, 0.3)
    hw = int(rng.choice([8, 12]))
    g = torch.Generator().manual_seed(300 + seed)
    batch = 8 + int(rng.choice([3, 5]))
    x = torch.randn(batch, model.c0, hw, hw, generator=g)
    y = torch.randint(0, 4, (batch,), generator=g)
    loss = torch.nn.CrossEntropyLoss()
    ref = ao.AutogradSpectralOperator(copy.deepcopy(model).double(), [x.double(), y], loss)
    P = sum(p.numel() for p in model.parameters())
    v = torch.randn(P, generator=g, dtype=torch.float64)
    v = (v / v.norm()).float().double()
    op = B200HVPOperator(model, [x, y], loss)
    hv = op.Hv(v, storedGrad=True)
    vg = op.vGHv(v, storedGrad=True)
    assert rel_err(op.stored_grad.cpu().numpy(), ref.gradient().detach().numpy()) < 1e-4
    assert rel_err(hv.cpu().numpy(), ref.hv(v).numpy()) < 1e-4
    assert rel_err(vg.cpu().numpy(), ref.vghv(v).numpy()) < 1e-4
