"""Architectures of the five BASELINE configs, restated for offline use.

The GPU box has no copy of the reference checkout, so the parity tests, the
bench and ``smoke()`` build their networks from here.  Every class produces
the SAME ``state_dict`` keys, the same ``model.parameters()`` order (this is
the flat layout of the spectral vectors, reference ``opt.py:102,191``) and
the same forward graph as the reference class it stands in for, so that a
reference ``state_dict`` loads unchanged and the tape tracer sees the same
topology for either object:

    ForestMLP        <-> forest_data.Net            (forest_data.py:75-89)
    UspsCNN          <-> usps_data.CNN              (usps_data.py:298-336)
    CifarDenseNet    <-> densenet.DenseNet3         (densenet.py:70-121)
    ChestVGG16bn     <-> dcnn.MyVggNet16_bn         (dcnn.py:238-252)
    ChestDenseNet121 <-> dcnn.MyDenseNet121         (dcnn.py:268-278)
    WeightedBCEWithLogits <-> dcnn.W_BCEWithLogitsLoss (dcnn.py:375-400)

Nothing here is on the product hot path; the hot path accepts any
``nn.Module`` built from the supported layer set (see ``tracer.py``).
"""
from __future__ import annotations

import math
from collections import OrderedDict

import torch
import torch.nn as nn
import torch.nn.functional as F

SEED = 1226  # the reference's own convention (forest_data.py:27-28, usps_data.py:69)


class ForestMLP(nn.Module):
    """54-20-20-20-7 covertype MLP; the middle layer is applied twice with
    shared weights and the output is a softmax fed into CrossEntropyLoss."""

    def __init__(self, width: int = 20, n_in: int = 54, n_out: int = 7):
        super().__init__()
        self.fc1 = nn.Linear(n_in, width)
        self.fc2 = nn.Linear(width, width)
        self.fc3 = nn.Linear(width, n_out)

    def forward(self, x):
        h = torch.relu(self.fc1(x))
        for _ in range(2):  # shared fc2, used twice
            h = torch.relu(self.fc2(h))
        return torch.softmax(self.fc3(h), dim=1)


class UspsCNN(nn.Module):
    """3 x [conv3x3(p1) -> relu -> maxpool2] -> fc 128-64 -> relu -> fc 64-10 -> softmax."""

    def __init__(self):
        super().__init__()
        chans = (1, 8, 16, 32)
        for i in range(3):
            setattr(self, "conv%d" % (i + 1), nn.Conv2d(chans[i], chans[i + 1], 3, 1, 1))
        self.pool = nn.MaxPool2d(2, 2, 0)
        self.fc1 = nn.Linear(128, 64)
        self.fc2 = nn.Linear(64, 10)

    def forward(self, x):
        h = x.view(-1, 1, 16, 16)
        for conv in (self.conv1, self.conv2, self.conv3):
            h = self.pool(torch.relu(conv(h)))
        h = torch.relu(self.fc1(h.view(-1, 128)))
        return torch.softmax(self.fc2(h), dim=1)


class _DenseBottleneck(nn.Module):
    def __init__(self, c_in: int, growth: int):
        super().__init__()
        self.bn1 = nn.BatchNorm2d(c_in)
        self.relu = nn.ReLU(inplace=True)
        self.conv1 = nn.Conv2d(c_in, 4 * growth, 1, bias=False)
        self.bn2 = nn.BatchNorm2d(4 * growth)
        self.conv2 = nn.Conv2d(4 * growth, growth, 3, padding=1, bias=False)

    def forward(self, x):
        y = self.conv1(self.relu(self.bn1(x)))
        y = self.conv2(self.relu(self.bn2(y)))
        return torch.cat([x, y], 1)


class _DenseTransition(nn.Module):
    def __init__(self, c_in: int, c_out: int):
        super().__init__()
        self.bn1 = nn.BatchNorm2d(c_in)
        self.relu = nn.ReLU(inplace=True)
        self.conv1 = nn.Conv2d(c_in, c_out, 1, bias=False)

    def forward(self, x):
        return F.avg_pool2d(self.conv1(self.relu(self.bn1(x))), 2)


class _DenseStage(nn.Module):
    def __init__(self, n_layers: int, c_in: int, growth: int):
        super().__init__()
        self.layer = nn.Sequential(*[_DenseBottleneck(c_in + i * growth, growth) for i in range(n_layers)])

    def forward(self, x):
        return self.layer(x)


class CifarDenseNet(nn.Module):
    """DenseNet-BC(depth, growth) for 32x32 inputs, bottleneck blocks, reduction 0.5."""

    def __init__(self, depth: int = 40, num_classes: int = 10, growth_rate: int = 12):
        super().__init__()
        per_stage = (depth - 4) // 6
        c = 2 * growth_rate
        self.conv1 = nn.Conv2d(3, c, 3, padding=1, bias=False)
        self.block1 = _DenseStage(per_stage, c, growth_rate)
        c += per_stage * growth_rate
        self.trans1 = _DenseTransition(c, c // 2)
        c //= 2
        self.block2 = _DenseStage(per_stage, c, growth_rate)
        c += per_stage * growth_rate
        self.trans2 = _DenseTransition(c, c // 2)
        c //= 2
        self.block3 = _DenseStage(per_stage, c, growth_rate)
        c += per_stage * growth_rate
        self.bn1 = nn.BatchNorm2d(c)
        self.relu = nn.ReLU(inplace=True)
        self.fc = nn.Linear(c, num_classes)
        self.in_planes = c
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                fan = m.kernel_size[0] * m.kernel_size[1] * m.out_channels
                nn.init.normal_(m.weight, 0.0, math.sqrt(2.0 / fan))
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.ones_(m.weight)
                nn.init.zeros_(m.bias)
            elif isinstance(m, nn.Linear):
                nn.init.zeros_(m.bias)

    def forward(self, x):
        h = self.trans1(self.block1(self.conv1(x)))
        h = self.block3(self.trans2(self.block2(h)))
        h = F.avg_pool2d(self.relu(self.bn1(h)), 8)
        return self.fc(h.view(-1, self.in_planes))


def _chest_tail(c_in: int) -> "OrderedDict[str, nn.Module]":
    tail = OrderedDict()
    tail["transit"] = nn.Sequential(nn.Conv2d(c_in, 1024, 3, padding=1), nn.BatchNorm2d(1024),
                                    nn.ReLU(inplace=True), nn.MaxPool2d(2, padding=1))
    tail["gpool"] = nn.MaxPool2d(4)
    return tail


class ChestVGG16bn(nn.Module):
    """torchvision vgg16_bn feature stack + 512->1024 transit conv + global max pool + 14-way linear."""

    def __init__(self, outnum: int = 14):
        super().__init__()
        from torchvision import models
        self.features = models.vgg16_bn(weights=None).features
        for name, mod in _chest_tail(512).items():
            self.features.add_module(name, mod)
        self.classifier = nn.Linear(1024, outnum)

    def forward(self, x):
        return self.classifier(self.features(x).view(-1, 1024))


class ChestDenseNet121(nn.Module):
    """torchvision densenet121 whose classifier is Linear -> Sigmoid (the loss applies a sigmoid again)."""

    def __init__(self, class_count: int = 14):
        super().__init__()
        from torchvision import models
        self.densenet121 = models.densenet121(weights=None)
        width = self.densenet121.classifier.in_features
        self.densenet121.classifier = nn.Sequential(nn.Linear(width, class_count), nn.Sigmoid())

    def forward(self, x):
        return self.densenet121(x)


class WeightedBCEWithLogits(nn.Module):
    """Per-class positive/negative re-weighted BCE-with-logits, NaN labels masked,
    mean over classes that have at least one valid label.

    With p = number of positive labels and s = number of valid labels in the whole
    batch, a positive entry weighs s/p and a negative one s/(s-p); degenerate
    batches (p in {0, s}) weigh label+1.
    """

    def forward(self, logits, target):
        n_cls = logits.shape[1]
        if 10 * target.shape[0] == logits.shape[0]:  # ten-crop evaluation
            target = target.repeat(10, 1)
        valid = ~torch.isnan(target)
        p = int(target[valid].sum().item())
        s = int(valid.sum().item())
        per_class = []
        for c in range(n_cls):
            keep = valid[:, c]
            z, t = logits[:, c][keep], target[:, c][keep]
            if p != 0 and p != s:
                w = t * (s / p - s / (s - p)) + s / (s - p)
            else:
                w = t + 1
            per_class.append(F.binary_cross_entropy_with_logits(z, t, w))
        f = torch.stack(per_class)
        return f[~torch.isnan(f)].mean()


CONFIGS = {
    # name: (factory, input shape without batch, n_classes, default batch, label kind)
    "forest": (ForestMLP, (54,), 7, 128, "class"),
    "usps": (UspsCNN, (1, 16, 16), 10, 128, "class"),
    "cifar_densenet": (lambda: CifarDenseNet(40, 10, 12), (3, 32, 32), 10, 32, "class"),
    "chest_vgg": (lambda: ChestVGG16bn(14), (3, 224, 224), 14, 4, "multihot"),
    "chest_densenet121": (lambda: ChestDenseNet121(14), (3, 224, 224), 14, 4, "multihot"),
}


def build(name: str, seed: int = SEED):
    """Seeded model + matching loss for a BASELINE config name."""
    factory, _, _, _, kind = CONFIGS[name]
    torch.manual_seed(seed)
    model = factory()
    loss = nn.CrossEntropyLoss() if kind == "class" else WeightedBCEWithLogits()
    return model, loss


def synthetic_batch(name: str, batch: int | None = None, seed: int = SEED):
    """N(0,1) inputs and random labels of the config's shape."""
    _, shape, n_cls, default_b, kind = CONFIGS[name]
    b = default_b if batch is None else batch
    g = torch.Generator().manual_seed(seed + 1)
    x = torch.randn((b,) + shape, generator=g)
    if kind == "class":
        y = torch.randint(0, n_cls, (b,), generator=g)
    else:
        y = (torch.rand((b, n_cls), generator=g) > 0.8).float()
    return x, y
