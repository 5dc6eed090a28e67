"""KLDivLoss on a one-hot target (opt.py:182-185 in prepare_grad, opt.py:566-569 in comp_f) through the C ABI: lowered
by the tracer to the cross-entropy head on the logits with the reduction's factor as the loss scale."""
import copy

import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("reduction", ["mean", "batchmean"])
def test_kldiv_one_hot_branch(reduction):
    from optwboundeigenval_b200.hvp_operator import B200HVPOperator
    from optwboundeigenval_b200.spectral import SpectralState
    from oracle import autograd_oracle as ao
    torch.manual_seed(3)
    model = torch.nn.Sequential(torch.nn.Linear(54, 20), torch.nn.ReLU(), torch.nn.Linear(20, 7),
                                torch.nn.LogSoftmax(dim=1))
    loss = torch.nn.KLDivLoss(reduction=reduction)
    x, y = torch.randn(64, 54), torch.randint(0, 7, (64,))
    ref = ao.AutogradSpectralOperator(copy.deepcopy(model), [x, y], loss)
    P = sum(p.numel() for p in model.parameters())
    v = torch.randn(P, generator=torch.Generator().manual_seed(4), dtype=torch.float64)
    v /= v.norm()
    evalm = copy.deepcopy(model).eval()
    with torch.no_grad():
        want = evalm(x)
        onehot = torch.zeros(want.shape).scatter_(1, y.view(-1, 1), 1)
        f_ref = loss(want, onehot).item()
    op = B200HVPOperator(model, [x, y], loss)
    hv = op.Hv(v, storedGrad=True)
    vg = op.vGHv(v, storedGrad=True)
    assert rel_err(op.stored_grad.cpu().numpy(), ref.gradient().detach().numpy()) < 1e-4
    assert abs(float(op.loss_value) - ref.loss_value) < 1e-5 * abs(ref.loss_value)
    assert rel_err(hv.cpu().numpy(), ref.hv(v).numpy()) < 1e-4
    assert rel_err(vg.cpu().numpy(), ref.vghv(v).numpy()) < 1e-4
    # comp_f: evaluation-mode loss and the model's own output (log-probabilities)
    st = SpectralState(model, loss)
    f, out = st.comp_f(x, y)
    assert abs(f - f_ref) < 1e-5 * abs(f_ref)
    assert rel_err(out.cpu().numpy(), want.numpy()) < 1e-5
