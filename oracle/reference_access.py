"""Locate and import the UNMODIFIED reference checkout -- TEST / BASELINE INFRASTRUCTURE ONLY.

Used by ``oracle/make_golden.py`` (fixtures), ``bench.py --impl reference`` and its ``cpu_baseline`` /
``gpu_autograd_yardstick`` legs, and the drop-in tests.  The product package never imports this file.

Where the reference lives: ``$OPTW_REFERENCE``, else ``/root/reference`` (build container), else
``baseline/_ref/optWBoundEigenval`` (the copy that travels to the GPU box; git-ignored).

Import recipe: ``import pandas`` first, stub ``matplotlib``, ``matplotlib.pyplot`` and
``pytz`` (imported at module scope by opt.py:17,34 and dcnn.py:9), put the checkout on ``sys.path``.
"""
from __future__ import annotations

import contextlib
import io
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def find_reference():
    """Path of a directory holding the reference's opt.py, or None."""
    cands = [os.environ.get("OPTW_REFERENCE"), "/root/reference",
             os.path.join(ROOT, "baseline", "_ref", "optWBoundEigenval")]
    for c in cands:
        if c and os.path.exists(os.path.join(c, "opt.py")):
            return os.path.abspath(c)
    return None


def import_reference(path=None):
    """Returns the reference's ``opt`` module (unmodified)."""
    path = path or find_reference()
    if path is None:
        raise FileNotFoundError("no reference checkout (OPTW_REFERENCE, /root/reference, baseline/_ref/optWBoundEigenval)")
    import pandas  # noqa: F401  (must precede the pytz stub)
    for name in ("matplotlib", "matplotlib.pyplot", "pytz"):
        if name not in sys.modules:
            try:
                __import__(name)
            except ImportError:
                sys.modules[name] = types.ModuleType(name)
    if not hasattr(sys.modules["matplotlib"], "pyplot"):
        sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    if path not in sys.path:
        sys.path.insert(0, path)
    with contextlib.redirect_stdout(io.StringIO()):
        import opt  # noqa
    return opt


def ref_model(kind: str):
    """The reference's own model and loss classes; torchvision backbones with ``weights=None`` (the constructors
    hard-code ``pretrained=True``, dcnn.py:206,241,272, and there is no network)."""
    import torch
    if kind == "forest":
        import forest_data
        return forest_data.Net(), torch.nn.CrossEntropyLoss()
    if kind == "usps":
        import usps_data
        return usps_data.CNN(), torch.nn.CrossEntropyLoss()
    if kind == "cifar_densenet":
        import densenet
        return densenet.DenseNet3(depth=40, growth_rate=12, num_classes=10), torch.nn.CrossEntropyLoss()
    import dcnn
    from torchvision import models
    names = ("vgg16_bn", "densenet121", "alexnet", "resnet50")
    orig = {k: getattr(models, k) for k in names}
    try:
        for k, fn in orig.items():
            setattr(models, k, (lambda f: (lambda *a, **kw: f(weights=None)))(fn))
        if kind == "chest_vgg":
            return dcnn.MyVggNet16_bn(14), dcnn.W_BCEWithLogitsLoss()
        if kind == "chest_densenet121":
            return dcnn.MyDenseNet121(14), dcnn.W_BCEWithLogitsLoss()
        if kind == "chest_alexnet":
            return dcnn.MyAlexNet(14), dcnn.W_BCEWithLogitsLoss()
        if kind == "chest_resnet50":
            return dcnn.MyResNet50(14), dcnn.W_BCEWithLogitsLoss()
        if kind == "chest_dnet121":            # dnet.py's custom-Function DenseNet (chestxray_best.py)
            return dcnn.DenseNet121(14, isTrained=False), dcnn.W_BCEWithLogitsLoss()
    finally:
        for k, fn in orig.items():
            setattr(models, k, fn)
    raise KeyError(kind)


def ref_model_with_state(kind: str, like):
    """Reference model carrying the weights and buffers of ``like`` (a zoo model: same state_dict keys)."""
    model, loss = ref_model(kind)
    model.load_state_dict(like.state_dict())
    return model, loss
