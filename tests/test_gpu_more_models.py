"""The remaining model families of the parameter files, built from the reference's OWN classes
(``baseline/_ref/optWBoundEigenval`` on the GPU box): ``dcnn.MyAlexNet`` (six chestxray_mu0_0*_K*.py files: 11x11 stride-4
and 5x5 convolutions, 3x3 stride-2 max pools), ``dcnn.DenseNet121`` on ``dnet.py``'s custom autograd Functions
(chestxray_best.py, chestxray_mu0.py) and residual networks (``dcnn.MyResNet50``, cifar100_ResNet_mu0.py), against the
CPU autograd oracle run on the same objects."""
import copy
import os
import sys

import numpy as np
import pytest
import torch

from conftest import rel_err

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu
RTOL_VEC = 1e-4


def _reference_model(kind):
    from oracle import reference_access as ra
    if ra.find_reference() is None:
        pytest.skip("no reference checkout (baseline/_ref/optWBoundEigenval)")
    ra.import_reference()
    torch.manual_seed(1226)
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        return ra.ref_model(kind)


def _check(model, loss, x, y, tol_vghv=RTOL_VEC):
    from optwboundeigenval_b200.hvp_operator import B200HVPOperator
    from oracle import autograd_oracle as ao
    model.train()
    P = sum(p.numel() for p in model.parameters())
    g = torch.Generator().manual_seed(7)
    v = torch.randn(P, generator=g, dtype=torch.float64)
    v /= v.norm()
    ref = ao.AutogradSpectralOperator(copy.deepcopy(model), [x, y], loss)
    g_ref, hv_ref, vg_ref = ref.gradient().detach().numpy(), ref.hv(v).numpy(), ref.vghv(v).numpy()
    op = B200HVPOperator(model, [x, y], loss)
    hv = op.Hv(v, storedGrad=True).cpu().numpy()
    grad = op.stored_grad.cpu().numpy()
    vg = op.vGHv(v, storedGrad=True).cpu().numpy()
    errs = (rel_err(grad, g_ref), rel_err(hv, hv_ref), rel_err(vg, vg_ref))
    print("grad %.2e hv %.2e vghv %.2e" % errs)
    assert abs(float(op.loss_value) - ref.loss_value) < 1e-5 * abs(ref.loss_value)
    if not (errs[0] < RTOL_VEC and errs[1] < RTOL_VEC and errs[2] < tol_vghv):
        from kinks import explain_by_kinks
        flips = explain_by_kinks(op, model, x, y, [("grad", "grad", None, grad), ("Hv", "hv", v, hv), ("vGHv", "vghv", v, vg)],
                                 rtol=RTOL_VEC)
        assert flips >= 1, errs
    return errs


def _multihot(batch, seed):
    g = torch.Generator().manual_seed(seed)
    return (torch.rand(batch, 14, generator=g) > 0.8).float()


def test_alexnet_chest_model():
    model, loss = _reference_model("chest_alexnet")
    g = torch.Generator().manual_seed(3)
    x = torch.randn(4, 3, 224, 224, generator=g)
    _check(model, loss, x, _multihot(4, 4))


def test_dnet_custom_function_densenet121():
    """dcnn.DenseNet121 = dnet.densenet121 (MyReLU / LinearFunction autograd Functions) + Linear -> Sigmoid head."""
    model, loss = _reference_model("chest_dnet121")
    g = torch.Generator().manual_seed(5)
    x = torch.randn(2, 3, 224, 224, generator=g)
    _check(model, loss, x, _multihot(2, 6))


def test_residual_network_bottlenecks():
    """torchvision Bottleneck stack as MyResNet50 uses it (identity and strided down-sample shortcuts, `out += identity`)
    at a size the CPU oracle finishes in seconds; the full dcnn.MyResNet50 is traced (tape structure) next to it."""
    import torch.nn as nn
    from torchvision.models.resnet import Bottleneck
    torch.manual_seed(9)

    class SmallResNet(nn.Module):
        def __init__(self):
            super().__init__()
            self.conv1 = nn.Conv2d(3, 16, 7, stride=2, padding=3, bias=False)
            self.bn1 = nn.BatchNorm2d(16)
            self.relu = nn.ReLU(inplace=True)
            self.maxpool = nn.MaxPool2d(3, stride=2, padding=1)
            d1 = nn.Sequential(nn.Conv2d(16, 64, 1, bias=False), nn.BatchNorm2d(64))
            d2 = nn.Sequential(nn.Conv2d(64, 128, 1, stride=2, bias=False), nn.BatchNorm2d(128))
            self.layer1 = nn.Sequential(Bottleneck(16, 16, downsample=d1), Bottleneck(64, 16))
            self.layer2 = nn.Sequential(Bottleneck(64, 32, stride=2, downsample=d2), Bottleneck(128, 32))
            self.transit = nn.Sequential(nn.Conv2d(128, 96, 3, padding=1), nn.BatchNorm2d(96), nn.ReLU(inplace=True),
                                         nn.MaxPool2d(2, padding=1))
            self.gpool = nn.MaxPool2d(3)
            self.classifier = nn.Linear(96, 14)

        def forward(self, x):
            h = self.maxpool(self.relu(self.bn1(self.conv1(x))))
            h = self.gpool(self.transit(self.layer2(self.layer1(h))))
            return self.classifier(h.view(-1, 96))

    from optwboundeigenval_b200 import zoo
    g = torch.Generator().manual_seed(10)
    x = torch.randn(6, 3, 64, 64, generator=g)
    _check(SmallResNet(), zoo.WeightedBCEWithLogits(), x, _multihot(6, 11))
    from optwboundeigenval_b200 import tracer
    model, loss = _reference_model("chest_resnet50")
    tape = tracer.trace(model, loss, (3, 1024, 1024))        # the input size at which MyResNet50's pooling tail is valid
    assert sum(1 for o in tape.ops if o.kind == tracer.OP_ADD) == 16 and tape.n_params == 42399822
