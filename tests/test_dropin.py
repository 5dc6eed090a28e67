"""Host-side wiring of the drop-in (CPU): install/uninstall on the real reference when it is present
(build container), offline shims, introspection-compatible constructor."""
import inspect
import os
import sys

import pytest

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from optwboundeigenval_b200.dropin import find_reference  # noqa: E402

# $OPTW_REFERENCE, /root/reference (build container) or baseline/_ref/optWBoundEigenval (the copy on the GPU box)
REF = find_reference() or "/nonexistent"
needs_ref = pytest.mark.skipif(not os.path.exists(os.path.join(REF, "opt.py")), reason="reference checkout not present")


def dropin_saved_iter(opt):
    return opt._b200_saved["iter"]


@needs_ref
def test_install_patches_and_restores_the_reference_module():
    from optwboundeigenval_b200 import dropin, spectral
    from optwboundeigenval_b200.hvp_operator import B200HVPOperator
    dropin.stub_plotting_modules()
    sys.path.insert(0, REF)
    import opt
    orig_cls_init = opt.OptWBoundEignVal.__init__
    orig_iter, orig_comp_f = opt.OptWBoundEignVal.iter, opt.OptWBoundEignVal.comp_f
    spec_before = inspect.getfullargspec(opt.OptWBoundEignVal)
    dropin.install(opt)
    try:
        assert opt.HVPOperator is B200HVPOperator
        assert opt.OptWBoundEignVal.comp_rho is spectral.comp_rho
        assert opt.OptWBoundEignVal.comp_gradrho is spectral.comp_gradrho
        assert opt.OptWBoundEignVal.comp_f is spectral.comp_f
        assert opt.OptWBoundEignVal.iter is not dropin_saved_iter(opt)
        # opt.missing_params / arg_dic introspect the constructor (opt.py:1940-1965)
        assert inspect.getfullargspec(opt.OptWBoundEignVal) == spec_before
        o = opt.missing_params(opt.OptWBoundEignVal, {"model": None, "loss": None, "optimizer": None})
        assert o["use_gpu"] is False and o["pow_iter_eps"] == 1e-3
    finally:
        dropin.uninstall(opt)
    assert opt.OptWBoundEignVal.__init__ is orig_cls_init
    assert opt.OptWBoundEignVal.iter is orig_iter and opt.OptWBoundEignVal.comp_f is orig_comp_f
    assert opt.HVPOperator.__name__ == "HVPOperator"


@needs_ref
def test_offline_shims_keep_the_reference_models():
    from optwboundeigenval_b200 import dropin, tracer
    import torch
    dropin.stub_plotting_modules()
    sys.path.insert(0, REF)
    import opt  # noqa: F401
    dropin.offline_shims(REF, n_train=16, n_eval=8)
    import forest_data
    import usps_data
    import cifar10_data
    d = forest_data.get_data()
    assert d["inputs"].shape == (16, 54)
    tr, va = usps_data.get_train_valid_loader(batch_size=4)
    assert next(iter(tr))[0].shape == (4, 1, 16, 16)
    assert len(cifar10_data.get_train_valid_loader(batch_size=4)) == 3
    # the reference's own model classes trace to the same tapes as the zoo restatements
    from optwboundeigenval_b200 import zoo
    for ref_model, kind in ((forest_data.Net(), "forest"), (usps_data.CNN(), "usps")):
        a = tracer.trace(ref_model, torch.nn.CrossEntropyLoss(), zoo.CONFIGS[kind][1])
        b = tracer.trace(zoo.build(kind)[0], torch.nn.CrossEntropyLoss(), zoo.CONFIGS[kind][1])
        assert [(o.kind, o.flags, o.w_off, o.b_off, o.geom) for o in a.ops] == \
               [(o.kind, o.flags, o.w_off, o.b_off, o.geom) for o in b.ops]
        assert a.head == b.head == tracer.HEAD_SOFTMAX_CE
    import densenet
    a = tracer.trace(densenet.DenseNet3(40, 10, 12), torch.nn.CrossEntropyLoss(), (3, 32, 32))
    b = tracer.trace(zoo.build("cifar_densenet")[0], torch.nn.CrossEntropyLoss(), (3, 32, 32))
    assert [(o.kind, o.flags, o.w_off, o.inp, o.out) for o in a.ops] == [(o.kind, o.flags, o.w_off, o.inp, o.out) for o in b.ops]


def test_wbce_coefficients_match_the_loss():
    """the per-entry coefficients handed to the CUDA head reproduce W_BCEWithLogitsLoss (dcnn.py:375-400)"""
    import torch
    from optwboundeigenval_b200 import zoo
    from optwboundeigenval_b200.hvp_operator import SpectralPlan
    g = torch.Generator().manual_seed(0)
    z = torch.randn(9, 14, generator=g)
    y = (torch.rand(9, 14, generator=g) > 0.8).float()
    y[2, 3] = float("nan")
    y[:, 5] = float("nan")                      # a class without any valid label is dropped
    t, coef = SpectralPlan.wbce_coefficients(y)
    mine = (coef * torch.nn.functional.binary_cross_entropy_with_logits(z, t, reduction="none")).sum()
    ref = zoo.WeightedBCEWithLogits()(z, y)
    assert abs(float(mine) - float(ref)) < 1e-6
    for degenerate in (torch.zeros(4, 3), torch.ones(4, 3)):       # p == 0 and p == s (dcnn.py:397)
        t, coef = SpectralPlan.wbce_coefficients(degenerate)
        z = torch.randn(4, 3, generator=g)
        mine = (coef * torch.nn.functional.binary_cross_entropy_with_logits(z, t, reduction="none")).sum()
        assert abs(float(mine) - float(zoo.WeightedBCEWithLogits()(z, degenerate))) < 1e-6


@needs_ref
def test_remaining_reference_model_classes_trace():
    """dcnn.MyAlexNet (six chestxray_mu0_0*_K*.py files), dcnn.DenseNet121 on dnet.py's custom autograd Functions
    (chestxray_best.py, chestxray_mu0.py) and dcnn.MyResNet50 (cifar100_ResNet_mu0.py) lower to tapes: residual adds,
    `_relu` / `_linear` leaf modules, 11x11 stride-4 and 7x7 stride-2 convolutions, 3x3 stride-2 max pools."""
    import contextlib
    import io
    import torch
    from optwboundeigenval_b200 import tracer
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from oracle import reference_access as ra
    ra.import_reference(REF)
    want = {"chest_alexnet": ((3, 224, 224), 4846414, 0), "chest_dnet121": ((3, 224, 224), 6968206, 0),
            "chest_resnet50": ((3, 1024, 1024), 42399822, 16)}
    for kind, (shape, n_params, n_add) in want.items():
        with contextlib.redirect_stdout(io.StringIO()):
            model, loss = ra.ref_model(kind)
        tape = tracer.trace(model, loss, shape)
        assert tape.n_params == n_params == sum(p.numel() for p in model.parameters())
        assert sum(1 for o in tape.ops if o.kind == tracer.OP_ADD) == n_add
        assert tape.head in (tracer.HEAD_WBCE, tracer.HEAD_SIGMOID_WBCE)
    # dnet's classifier is Linear -> Sigmoid behind a custom Function: the sigmoid is folded into the head
    assert tracer.trace(ra.ref_model("chest_dnet121")[0], loss, (3, 224, 224)).head == tracer.HEAD_SIGMOID_WBCE
