"""The jet recurrences over the traced tape (what the CUDA library implements) must agree with
nested autograd (what the reference does, pinned by golden vectors).  CPU only, fp64."""
import numpy as np
import pytest
import torch

from conftest import rel_err
from optwboundeigenval_b200 import tracer, zoo
from optwboundeigenval_b200.hvp_operator import SpectralPlan
from oracle import autograd_oracle as ao
from oracle.jet_oracle import JetTapeOracle

CASES = [("forest", 16), ("usps", 8), ("cifar_densenet", 4)]


def _setup(kind, batch, dtype=torch.float64):
    model, loss = zoo.build(kind)
    model.train()
    if kind == "cifar_densenet":      # make BN affine parameters non-trivial
        g = torch.Generator().manual_seed(5)
        for m in model.modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                m.weight.data = 1 + 0.3 * torch.randn(m.weight.shape, generator=g)
                m.bias.data = 0.2 * torch.randn(m.bias.shape, generator=g)
    x, y = zoo.synthetic_batch(kind, batch)
    return model, loss, x, y


def _fd_second(model, loss, x, y, v, h=1e-3):
    """central second difference of the gradient along v: d^2/dt^2 grad E(w + t v)"""
    w0 = ao.flat_params(model).clone()

    def grad_at(t):
        off = 0
        for p in model.parameters():
            k = p.numel()
            p.data = (w0 + t * v)[off:off + k].view(p.shape).clone()
            off += k
        out = torch.autograd.grad(loss(model(x), y), list(model.parameters()))
        return torch.cat([g.reshape(-1) for g in out])

    fd = (grad_at(h) - 2 * grad_at(0.0) + grad_at(-h)) / h ** 2
    grad_at(0.0)
    return fd


@pytest.mark.parametrize("kind,batch", CASES)
def test_jets_match_autograd_fp64(kind, batch):
    model, loss, x, y = _setup(kind, batch)
    tape = tracer.trace(model, loss, zoo.CONFIGS[kind][1])
    model64 = model.double()
    op = ao.AutogradSpectralOperator(model64, [x.double(), y], loss)
    P = tape.n_params
    g = torch.Generator().manual_seed(11)
    v = torch.randn(P, generator=g, dtype=torch.float64)
    v /= v.norm()
    v = v.float().double()          # fp32-representable so both sides see the same vector
    jo = JetTapeOracle(tape, ao.flat_params(model64), x, y)
    g0 = jo.run(0).clone()
    assert rel_err(g0.numpy(), op.gradient().detach().numpy()) < 1e-10
    assert abs(float(jo.loss) - op.loss_value) < 1e-10
    hv = jo.run(1, v).clone()
    assert rel_err(hv.numpy(), op.hv(v).numpy()) < 1e-9
    vghv = jo.vghv(v, mode="reference")
    assert rel_err(vghv.numpy(), op.vghv(v).numpy()) < 1e-8
    if kind == "cifar_densenet":
        # torch's third order through native batch norm is NOT the true derivative
        assert rel_err(jo.vghv(v, mode="exact").numpy(), op.vghv(v).numpy()) > 1e-3


@pytest.mark.parametrize("kind", ["chest_vgg_tiny", "chest_densenet_tiny"])
def test_jets_match_autograd_weighted_bce(kind):
    """Both chest heads (raw logits, and sigmoid in front of BCE-with-logits) on a small stand-in network
    with the same layer kinds (conv+bias, BN, ReLU, max pool with padding, concatenation)."""
    torch.manual_seed(3)
    nn = torch.nn

    class TinyVgg(nn.Module):
        def __init__(self):
            super().__init__()
            self.features = nn.Sequential(nn.Conv2d(3, 6, 3, padding=1), nn.BatchNorm2d(6), nn.ReLU(inplace=True),
                                          nn.MaxPool2d(2, 2), nn.Conv2d(6, 8, 3, padding=1), nn.BatchNorm2d(8),
                                          nn.ReLU(inplace=True), nn.MaxPool2d(2, padding=1), nn.MaxPool2d(3))
            self.classifier = nn.Linear(8, 5)

        def forward(self, x):
            return self.classifier(self.features(x).view(-1, 8))

    class TinyDense(nn.Module):
        def __init__(self):
            super().__init__()
            self.conv0 = nn.Conv2d(3, 4, 7, stride=2, padding=3, bias=False)
            self.norm0 = nn.BatchNorm2d(4)
            self.pool0 = nn.MaxPool2d(3, stride=2, padding=1)
            self.n1, self.c1 = nn.BatchNorm2d(4), nn.Conv2d(4, 3, 1, bias=False)
            self.n2, self.c2 = nn.BatchNorm2d(7), nn.Conv2d(7, 3, 3, padding=1, bias=False)
            self.tn, self.tc, self.tp = nn.BatchNorm2d(10), nn.Conv2d(10, 5, 1, bias=False), nn.AvgPool2d(2, 2)
            self.norm5 = nn.BatchNorm2d(5)
            self.classifier = nn.Sequential(nn.Linear(5, 5), nn.Sigmoid())

        def forward(self, x):
            f0 = self.pool0(torch.relu(self.norm0(self.conv0(x))))
            feats = [f0]
            feats.append(self.c1(torch.relu(self.n1(torch.cat(feats, 1)))))
            feats.append(self.c2(torch.relu(self.n2(torch.cat(feats, 1)))))
            h = self.tp(self.tc(torch.relu(self.tn(torch.cat(feats, 1)))))
            h = torch.nn.functional.relu(self.norm5(h), inplace=True)
            h = torch.flatten(torch.nn.functional.adaptive_avg_pool2d(h, (1, 1)), 1)
            return self.classifier(h)

    model = (TinyVgg() if kind == "chest_vgg_tiny" else TinyDense()).double().train()
    loss = zoo.WeightedBCEWithLogits()
    B = 6
    x = torch.randn(B, 3, 16, 16, dtype=torch.float64)
    y = (torch.rand(B, 5) > 0.7).double()
    y[0, 1] = float("nan")           # masked label
    tape = tracer.trace(model, loss, (3, 16, 16))
    assert not any(o.kind == tracer.OP_COPY for o in tape.ops)
    op = ao.AutogradSpectralOperator(model, [x, y], loss)
    t, coef = SpectralPlan.wbce_coefficients(y.float())
    jo = JetTapeOracle(tape, ao.flat_params(model), x, t, coef, loss_scale=1.0)
    P = tape.n_params
    v = torch.randn(P, dtype=torch.float64)
    v = (v / v.norm()).float().double()
    assert rel_err(jo.run(0).numpy(), op.gradient().detach().numpy()) < 1e-6     # coef is fp32
    assert abs(float(jo.loss) - op.loss_value) < 1e-6
    assert rel_err(jo.run(1, v).numpy(), op.hv(v).numpy()) < 1e-6
    assert rel_err(jo.vghv(v, mode="reference").numpy(), op.vghv(v).numpy()) < 1e-6


def test_full_size_chest_models_trace():
    for kind in ("chest_vgg", "chest_densenet121"):
        model, loss = zoo.build(kind)
        tape = tracer.trace(model, loss, zoo.CONFIGS[kind][1])
        assert tape.n_params == sum(p.numel() for p in model.parameters())
        assert not any(o.kind == tracer.OP_COPY for o in tape.ops)     # all concatenations alias
        assert not any(o.kind == tracer.OP_RELU for o in tape.ops)     # all ReLUs fused


def test_exact_third_order_is_the_true_derivative_and_torch_bn_is_not():
    """On a smooth network (no ReLU) with train-mode BatchNorm, d^2/dt^2 grad E(w+tv) by central
    differences agrees with the exact jets, while nested autograd through native_batch_norm does not:
    batchnorm_double_backward reads mean / invstd from saved non-differentiable tensors."""
    torch.manual_seed(0)
    nn = torch.nn

    class Smooth(nn.Module):
        def __init__(self):
            super().__init__()
            self.c = nn.Conv2d(2, 3, 3, padding=1, bias=False)
            self.b = nn.BatchNorm2d(3)
            self.f = nn.Linear(48, 4)

        def forward(self, x):
            return self.f(self.b(self.c(x)).view(-1, 48))

    model = Smooth().double().train()
    model.b.weight.data = 1 + 0.3 * torch.randn(3, dtype=torch.float64)
    model.b.bias.data = 0.2 * torch.randn(3, dtype=torch.float64)
    x = torch.randn(5, 2, 4, 4, dtype=torch.float64)
    y = torch.randint(0, 4, (5,))
    loss = nn.CrossEntropyLoss()
    tape = tracer.trace(model, loss, (2, 4, 4))
    v = torch.randn(tape.n_params, dtype=torch.float64)
    v = (v / v.norm()).float().double()
    op = ao.AutogradSpectralOperator(model, [x, y], loss)
    ref = op.vghv(v)
    jo = JetTapeOracle(tape, ao.flat_params(model), x, y)
    jo.run(0)
    exact = jo.vghv(v, mode="exact")
    compat = jo.vghv(v, mode="reference")
    fd = _fd_second(model, loss, x, y, v)
    assert rel_err(exact.numpy(), fd.numpy()) < 1e-5
    assert rel_err(compat.numpy(), ref.numpy()) < 1e-10
    assert rel_err(ref.numpy(), fd.numpy()) > 1e-2


class _TinyResNet(torch.nn.Module):
    """torchvision-style bottleneck blocks: identity and strided 1x1 down-sample shortcuts, `out += identity`, one
    shared in-place ReLU module per block (dcnn.py:219-236 builds MyResNet50 from these)."""

    def __init__(self):
        super().__init__()
        nn = torch.nn
        self.conv1 = nn.Conv2d(3, 8, 7, stride=2, padding=3, bias=False)
        self.bn1 = nn.BatchNorm2d(8)
        self.relu = nn.ReLU(inplace=True)
        self.maxpool = nn.MaxPool2d(3, stride=2, padding=1)
        from torchvision.models.resnet import Bottleneck
        down = nn.Sequential(nn.Conv2d(8, 16, 1, stride=2, bias=False), nn.BatchNorm2d(16))
        self.layer1 = nn.Sequential(Bottleneck(8, 4, stride=2, downsample=down), Bottleneck(16, 4))
        self.fc = nn.Linear(16, 5)

    def forward(self, x):
        h = self.maxpool(self.relu(self.bn1(self.conv1(x))))
        h = self.layer1(h)
        h = torch.nn.functional.adaptive_avg_pool2d(h, (1, 1))
        return self.fc(torch.flatten(h, 1))


def test_residual_add_jets_match_autograd_fp64():
    """OP_ADD (residual connections, fused ReLU, two-operand overwrite/accumulate analysis) against nested autograd."""
    torch.manual_seed(21)
    model = _TinyResNet().train()
    loss = torch.nn.CrossEntropyLoss()
    g = torch.Generator().manual_seed(22)
    x = torch.randn(6, 3, 32, 32, generator=g)
    y = torch.randint(0, 5, (6,), generator=g)
    tape = tracer.trace(model, loss, (3, 32, 32))
    assert sum(1 for o in tape.ops if o.kind == tracer.OP_ADD) == 2
    assert all(o.flags & tracer.F_RELU for o in tape.ops if o.kind == tracer.OP_ADD)
    model64 = model.double()
    op = ao.AutogradSpectralOperator(model64, [x.double(), y], loss)
    P = tape.n_params
    v = torch.randn(P, generator=g, dtype=torch.float64)
    v = (v / v.norm()).float().double()
    jo = JetTapeOracle(tape, ao.flat_params(model64), x, y)
    assert rel_err(jo.run(0).numpy(), op.gradient().detach().numpy()) < 1e-10
    assert rel_err(jo.run(1, v).numpy(), op.hv(v).numpy()) < 1e-9
    assert rel_err(jo.vghv(v).numpy(), op.vghv(v).numpy()) < 1e-8


def test_dnet_custom_function_modules_lower_to_relu_and_linear():
    """dnet.py's hand-written autograd Functions (MyReLU dnet.py:30-60, LinearFunction dnet.py:64-99) behind the
    `_relu` / `_linear` modules: same tape as the built-in layers, same jets as nested autograd through the custom
    backward passes (restated here; the reference's own classes are traced in tests/test_dropin.py)."""
    class MyReLU(torch.autograd.Function):
        @staticmethod
        def forward(ctx, inp):
            ctx.save_for_backward(inp)
            return inp.clamp(min=0)

        @staticmethod
        def backward(ctx, grad_output):
            inp, = ctx.saved_tensors
            gi = grad_output.clone()
            gi[inp < 0] = 0
            return gi

    class LinearFunction(torch.autograd.Function):
        @staticmethod
        def forward(ctx, inp, weight, bias=None):
            ctx.save_for_backward(inp, weight, bias)
            out = inp.mm(weight.t())
            if bias is not None:
                out += bias.unsqueeze(0).expand_as(out)
            return out

        @staticmethod
        def backward(ctx, go):
            inp, weight, bias = ctx.saved_tensors
            return go.mm(weight), go.t().mm(inp), go.sum(0)

    class _relu(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.f = MyReLU.apply

        def forward(self, x):
            return self.f(x)

    class _linear(torch.nn.Module):
        def __init__(self, n_in, n_out):
            super().__init__()
            self.weight = torch.nn.Parameter(torch.empty(n_out, n_in).uniform_(-0.3, 0.3))
            self.bias = torch.nn.Parameter(torch.empty(n_out).uniform_(-0.1, 0.1))

        def forward(self, x):
            return LinearFunction.apply(x, self.weight, self.bias)

    class Net(torch.nn.Module):
        def __init__(self):
            super().__init__()
            nn = torch.nn
            self.features = nn.Sequential(nn.Conv2d(3, 6, 3, padding=1, bias=False), nn.BatchNorm2d(6), _relu(),
                                          nn.Conv2d(6, 4, 1, bias=False), nn.BatchNorm2d(4))
            self.classifier = _linear(4, 3)

        def forward(self, x):
            h = torch.nn.functional.relu(self.features(x), inplace=True)
            h = torch.flatten(torch.nn.functional.adaptive_avg_pool2d(h, (1, 1)), 1)
            return self.classifier(h)

    torch.manual_seed(31)
    model = Net().train()
    loss = torch.nn.CrossEntropyLoss()
    g = torch.Generator().manual_seed(32)
    x = torch.randn(5, 3, 8, 8, generator=g)
    y = torch.randint(0, 3, (5,), generator=g)
    tape = tracer.trace(model, loss, (3, 8, 8))
    kinds = [o.kind for o in tape.ops]
    assert kinds == [tracer.OP_CONV, tracer.OP_BN, tracer.OP_CONV, tracer.OP_BN, tracer.OP_AVGPOOL, tracer.OP_CONV]
    model64 = model.double()
    op = ao.AutogradSpectralOperator(model64, [x.double(), y], loss)
    v = torch.randn(tape.n_params, generator=g, dtype=torch.float64)
    v = (v / v.norm()).float().double()
    jo = JetTapeOracle(tape, ao.flat_params(model64), x, y)
    assert rel_err(jo.run(0).numpy(), op.gradient().detach().numpy()) < 1e-10
    assert rel_err(jo.run(1, v).numpy(), op.hv(v).numpy()) < 1e-9
    assert rel_err(jo.vghv(v).numpy(), op.vghv(v).numpy()) < 1e-8


@pytest.mark.parametrize("reduction", ["mean", "batchmean", "sum"])
def test_kldiv_one_hot_branch_is_the_cross_entropy_head(reduction):
    """opt.py:182-185: KLDivLoss(model(x), one_hot(y)).  On log-probabilities that is the cross entropy of the logits
    times the reduction factor, which is how the tracer lowers it (Tape.head_scale); fp32 as the reference casts."""
    torch.manual_seed(3)
    model = torch.nn.Sequential(torch.nn.Linear(9, 12), torch.nn.ReLU(), torch.nn.Linear(12, 5),
                                torch.nn.LogSoftmax(dim=1))
    loss = torch.nn.KLDivLoss(reduction=reduction)
    x, y = torch.randn(6, 9), torch.randint(0, 5, (6,))
    tape = tracer.trace(model, loss, (9,))
    assert tape.head == tracer.HEAD_CE and tape.kl_reduction == reduction
    op = ao.AutogradSpectralOperator(model, [x, y], loss)
    P = tape.n_params
    v = torch.randn(P, generator=torch.Generator().manual_seed(4), dtype=torch.float64)
    v = (v / v.norm()).float().double()
    jo = JetTapeOracle(tape, ao.flat_params(model).double(), x, y, loss_scale=tape.head_scale(6, 5))
    assert rel_err(jo.run(0).numpy(), op.gradient().detach().numpy()) < 2e-6
    assert abs(float(jo.loss) - op.loss_value) < 1e-6 * abs(op.loss_value)
    assert rel_err(jo.run(1, v).numpy(), op.hv(v).numpy()) < 2e-5
    assert rel_err(jo.vghv(v).numpy(), op.vghv(v).numpy()) < 2e-4
    with pytest.raises(tracer.UnsupportedModel):
        tracer.trace(model[:3], loss, (9,))                  # raw logits into KLDivLoss
    with pytest.raises(tracer.UnsupportedModel):
        tracer.trace(model, torch.nn.CrossEntropyLoss(), (9,))


class _RandomNet(torch.nn.Module):
    """A seeded small network mixing everything the tracer lowers: biased / unbiased convolutions of odd geometry,
    BatchNorm with and without ReLU, both poolings, DenseNet-style concatenation, a residual add, a parameter shared
    by two call sites, a softmax tail."""

    def __init__(self, rng):
        super().__init__()
        nn = torch.nn
        pick = lambda *a: a[int(rng.integers(len(a)))]                      # noqa: E731
        self.c0 = int(pick(3, 5))
        c1, c2 = int(pick(6, 8)), int(pick(4, 7))
        self.k1, self.s1 = int(pick(1, 3, 5)), int(pick(1, 2))
        self.conv1 = nn.Conv2d(self.c0, c1, self.k1, stride=self.s1, padding=self.k1 // 2, bias=bool(pick(0, 1)))
        self.bn1 = nn.BatchNorm2d(c1)
        self.conv2 = nn.Conv2d(c1, c2, 3, padding=1, bias=False)
        self.bn2 = nn.BatchNorm2d(c1 + c2)
        self.conv3 = nn.Conv2d(c1 + c2, c1 + c2, int(pick(1, 3)), padding="same", bias=bool(pick(0, 1)))
        self.pool = nn.MaxPool2d(2) if pick(0, 1) else nn.AvgPool2d(2)
        self.shared = nn.Linear(c1 + c2, c1 + c2)
        self.fc = nn.Linear(c1 + c2, 4)
        self.tail = bool(pick(0, 1))

    def forward(self, x):
        F = torch.nn.functional
        h = F.relu(self.bn1(self.conv1(x)))
        h = torch.cat([h, self.conv2(h)], 1)                 # DenseNet-style growth
        h = self.bn2(h)                                       # BatchNorm without ReLU
        h = F.relu(h + self.conv3(h))                         # residual add
        h = self.pool(h)
        h = torch.flatten(F.adaptive_avg_pool2d(h, (1, 1)), 1)
        h = F.relu(self.shared(h))
        h = F.relu(self.shared(h))                            # same Linear twice (forest_data.py:75-89)
        h = self.fc(h)
        return F.softmax(h, dim=1) if self.tail else h


@pytest.mark.parametrize("seed", [0, 1, 2, 3, 4, 5])
def test_random_architectures_jets_match_autograd_fp64(seed):
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
    batch = int(rng.choice([3, 5]))
    x = torch.randn(batch, model.c0, hw, hw, generator=g)
    y = torch.randint(0, 4, (batch,), generator=g)
    loss = torch.nn.CrossEntropyLoss()
    tape = tracer.trace(model, loss, (model.c0, hw, hw))
    model64 = model.double()
    op = ao.AutogradSpectralOperator(model64, [x.double(), y], loss)
    P = tape.n_params
    assert P == sum(p.numel() for p in model.parameters())
    v = torch.randn(P, generator=g, dtype=torch.float64)
    v = (v / v.norm()).float().double()
    jo = JetTapeOracle(tape, ao.flat_params(model64), x, y)
    assert rel_err(jo.run(0).numpy(), op.gradient().detach().numpy()) < 1e-10
    assert abs(float(jo.loss) - op.loss_value) < 1e-10
    assert rel_err(jo.run(1, v).numpy(), op.hv(v).numpy()) < 1e-9
    assert rel_err(jo.vghv(v, mode="reference").numpy(), op.vghv(v).numpy()) < 1e-8


def _head(c):
    nn = torch.nn
    return [nn.AdaptiveAvgPool2d((1, 1)), nn.Flatten(), nn.Linear(c, 4)]


def _edge_models():
    nn = torch.nn
    conv = lambda *a, **k: nn.Conv2d(*a, **k)                                                   # noqa: E731
    ok = {
        "maxpool 3 s2 p1 (DenseNet121 stem)": (nn.Sequential(conv(3, 6, 3, padding=1), nn.ReLU(), nn.MaxPool2d(3, 2, 1), *_head(6)), (3, 9, 9)),
        "maxpool 3 s2 (AlexNet)": (nn.Sequential(conv(3, 6, 3, padding=1), nn.ReLU(), nn.MaxPool2d(3, 2), *_head(6)), (3, 9, 9)),
        "maxpool 2 p1 (VGG transit, dcnn.py:238-252)": (nn.Sequential(conv(3, 6, 3, padding=1), nn.ReLU(), nn.MaxPool2d(2, padding=1), *_head(6)), (3, 8, 8)),
        "avgpool 2 on an odd map": (nn.Sequential(conv(3, 6, 3, padding=1), nn.ReLU(), nn.AvgPool2d(2), *_head(6)), (3, 9, 9)),
        "conv 5 s2 p2 + BN + in-place ReLU": (nn.Sequential(conv(3, 6, 5, 2, 2), nn.BatchNorm2d(6), nn.ReLU(inplace=True), *_head(6)), (3, 9, 9)),
        "rectangular kernel": (nn.Sequential(conv(3, 6, (3, 1), padding=(1, 0)), nn.ReLU(), *_head(6)), (3, 8, 8)),
        "rectangular input": (nn.Sequential(conv(3, 6, 3, padding=1), nn.ReLU(), nn.MaxPool2d(2), nn.Flatten(), nn.Linear(90, 4)), (3, 6, 10)),
        "padding='same' / 'valid'": (nn.Sequential(conv(3, 6, 3, padding="same"), nn.ReLU(), conv(6, 6, 3, padding="valid"), *_head(6)), (3, 8, 8)),
        "BatchNorm on the input": (nn.Sequential(nn.BatchNorm2d(3), conv(3, 6, 3, padding=1), nn.ReLU(), *_head(6)), (3, 8, 8)),
        "ReLU on the input, ReLU twice": (nn.Sequential(nn.ReLU(), conv(3, 6, 3, padding=1), nn.ReLU(), nn.ReLU(), *_head(6)), (3, 8, 8)),
        "BatchNorm1d after Linear": (nn.Sequential(nn.Flatten(), nn.Linear(12, 8), nn.BatchNorm1d(8), nn.ReLU(), nn.Linear(8, 4)), (3, 2, 2)),
        "MLP with a softmax tail": (nn.Sequential(nn.Linear(7, 8), nn.ReLU(), nn.Linear(8, 4), nn.Softmax(dim=1)), (7,)),
        "Linear without bias, Dropout(0)": (nn.Sequential(nn.Linear(7, 8, bias=False), nn.Dropout(0.0), nn.ReLU(), nn.Linear(8, 4)), (7,)),
        "convolution output as logits": (nn.Sequential(conv(3, 4, 3, padding=1), nn.AdaptiveAvgPool2d((1, 1)), nn.Flatten()), (3, 8, 8)),
    }
    refused = {
        "max pooling with ceil_mode": (nn.Sequential(conv(3, 6, 3, padding=1), nn.MaxPool2d(2, ceil_mode=True), *_head(6)), (3, 9, 9)),
        "overlapping average pooling": (nn.Sequential(conv(3, 6, 3, padding=1), nn.AvgPool2d(3, 2, 1), *_head(6)), (3, 9, 9)),
        "adaptive pooling to 2x2": (nn.Sequential(conv(3, 6, 3, padding=1), nn.AdaptiveAvgPool2d((2, 2)), nn.Flatten(), nn.Linear(24, 4)), (3, 8, 8)),
        "grouped convolution": (nn.Sequential(conv(4, 8, 3, padding=1, groups=2), *_head(8)), (4, 8, 8)),
        "dilated convolution": (nn.Sequential(conv(3, 6, 3, padding=2, dilation=2), *_head(6)), (3, 8, 8)),
        "padding='same' on an even kernel": (nn.Sequential(conv(3, 6, 2, padding="same"), *_head(6)), (3, 8, 8)),
        "dropout with p > 0": (nn.Sequential(nn.Linear(7, 8), nn.Dropout(0.5), nn.Linear(8, 4)), (7,)),
        "an activation without a kernel": (nn.Sequential(nn.Linear(7, 8), nn.Tanh(), nn.Linear(8, 4)), (7,)),
    }
    return ok, refused


@pytest.mark.parametrize("name", sorted(_edge_models()[0]))
def test_edge_geometries_match_autograd_fp64(name):
    """Layer geometries at the edges of what the five configs use: each must reproduce nested autograd in fp64."""
    torch.manual_seed(1)
    model, shape = _edge_models()[0][name]
    model.train()
    loss = torch.nn.CrossEntropyLoss()
    g = torch.Generator().manual_seed(2)
    x = torch.randn(3, *shape, generator=g)
    y = torch.randint(0, 4, (3,), generator=g)
    tape = tracer.trace(model, loss, shape)
    m64 = model.double()
    op = ao.AutogradSpectralOperator(m64, [x.double(), y], loss)
    v = torch.randn(tape.n_params, generator=g, dtype=torch.float64)
    v = (v / v.norm()).float().double()
    jo = JetTapeOracle(tape, ao.flat_params(m64), x, y)
    assert rel_err(jo.run(0).numpy(), op.gradient().detach().numpy()) < 1e-10
    assert rel_err(jo.run(1, v).numpy(), op.hv(v).numpy()) < 1e-9
    assert rel_err(jo.vghv(v, mode="reference").numpy(), op.vghv(v).numpy()) < 1e-8


@pytest.mark.parametrize("name", sorted(_edge_models()[1]))
def test_unsupported_geometries_are_refused_by_name(name):
    """No silent approximation and no CPU fallback: what has no kernel raises UnsupportedModel at trace time."""
    model, shape = _edge_models()[1][name]
    with pytest.raises(tracer.UnsupportedModel):
        tracer.trace(model.train(), torch.nn.CrossEntropyLoss(), shape)
