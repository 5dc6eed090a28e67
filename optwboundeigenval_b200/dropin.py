"""Install the B200 path behind the reference's own entry points.

    import opt                                   # the reference module (opt.py)
    from optwboundeigenval_b200 import dropin
    dropin.install(opt)                          # opt.main('<params file>') now runs the hot path on the GPU

What is replaced (all resolved at call time by the reference, so no source edit is needed):

* ``opt.HVPOperator``  ->  ``B200HVPOperator``           (constructed per minibatch at opt.py:424)
* ``OptWBoundEignVal.comp_rho / comp_gradrho``  ->  fused device loop (``spectral.py``)
* ``OptWBoundEignVal.comp_f``  ->  forward-only evaluation pass (``b2s_eval_pass``); ``test_model`` reaches it unchanged
* ``OptWBoundEignVal.iter``  ->  ``spectral.iter_epoch``: the same epoch with the minibatch body's clip, step assembly
  and SGD / Adam update fused on flat vectors (opt.py:535-542, 616-659, 696-699)
* ``OptWBoundEignVal.__init__``: ``use_gpu`` is forced on -- many parameter files say
  ``use_gpu=False`` (forest_best.py:43, usps_CNN_lobpcg.py:52, chestxray_best_reg.py:110) and the
  B200 path has no CPU variant; with it the trainer keeps model, ``self.v`` and gradients on the device.

``offline_shims`` provides synthetic stand-ins for the data modules the parameter files import
(every one of them downloads data or reads private paths) while re-exporting the
reference's model and loss classes unchanged, so ``python main.py <pfile>`` works without network.
"""
from __future__ import annotations

import functools
import importlib.util
import os
import sys
import types

import numpy as np
import torch


def find_reference():
    """A directory holding the reference's ``opt.py``: ``$OPTW_REFERENCE``, ``/root/reference``, or the copy under
    ``baseline/_ref/optWBoundEigenval`` next to this package (None when there is none)."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for c in (os.environ.get("OPTW_REFERENCE"), "/root/reference", os.path.join(root, "baseline", "_ref", "optWBoundEigenval")):
        if c and os.path.exists(os.path.join(c, "opt.py")):
            return os.path.abspath(c)
    return None


def install(opt_module, fused_loop: bool = True, force_gpu: bool = True):
    """Patch the reference module in place; returns a dict of the originals for ``uninstall``."""
    from . import spectral
    from .hvp_operator import B200HVPOperator
    cls = opt_module.OptWBoundEignVal
    saved = {"HVPOperator": opt_module.HVPOperator, "comp_rho": cls.comp_rho, "comp_gradrho": cls.comp_gradrho,
             "__init__": cls.__init__}
    opt_module.HVPOperator = B200HVPOperator
    if fused_loop:
        from . import kfac as _kfac
        saved.update({"kfac": cls.kfac, "init_kfac": cls.init_kfac})
        cls.comp_rho = spectral.comp_rho
        cls.comp_gradrho = spectral.comp_gradrho
        saved["comp_f"] = cls.comp_f
        cls.comp_f = spectral.comp_f
        cls.kfac = _kfac.kfac
        cls.init_kfac = _kfac.init_kfac
        # additions (no reference method is replaced): the fused step assembly of iter()'s minibatch body and the
        # replicas-only multi-GPU form of the rho_test sweep
        cls.assemble_step = spectral.assemble_step
        cls.fused_step = spectral.fused_step
        cls.rho_sweep = spectral.rho_test
        # iter() (opt.py:580-763): same epoch, minibatch body = comp_g + fused clip / assembly / optimizer update; the
        # branches that are not the plain power-iteration step run the reference's own iter()
        saved["iter"] = reference_iter = cls.iter

        def iter(self):
            import types
            self._b200_reference_iter = types.MethodType(reference_iter, self)
            return spectral.iter_epoch(self)
        cls.iter = iter
    if force_gpu:
        orig_init = cls.__init__

        @functools.wraps(orig_init)
        def init(self, *args, **kwargs):
            import inspect
            names = inspect.getfullargspec(orig_init).args[1:]
            if "use_gpu" in names:
                idx = names.index("use_gpu")
                if len(args) > idx:
                    args = args[:idx] + (True,) + args[idx + 1:]
                else:
                    kwargs["use_gpu"] = True
            orig_init(self, *args, **kwargs)

        # missing_params()/arg_dic() introspect the constructor signature (opt.py:1940-1965)
        import inspect
        init.__signature__ = inspect.signature(orig_init)
        cls.__init__ = init
    opt_module._b200_saved = saved
    return saved


def uninstall(opt_module):
    saved = getattr(opt_module, "_b200_saved", None)
    if not saved:
        return
    cls = opt_module.OptWBoundEignVal
    opt_module.HVPOperator = saved["HVPOperator"]
    cls.comp_rho, cls.comp_gradrho, cls.__init__ = saved["comp_rho"], saved["comp_gradrho"], saved["__init__"]
    if "kfac" in saved:
        cls.kfac, cls.init_kfac = saved["kfac"], saved["init_kfac"]
        cls.iter, cls.comp_f = saved["iter"], saved["comp_f"]
        for extra in ("assemble_step", "fused_step", "rho_sweep"):
            if extra in cls.__dict__:
                delattr(cls, extra)
    del opt_module._b200_saved


# ---------------------------------------------------------------------------------------------------
# offline environment for the reference (stubs + synthetic data modules)
# ---------------------------------------------------------------------------------------------------
def stub_plotting_modules():
    """opt.py:17,34 and dcnn.py:9 import matplotlib / pytz at module scope; neither is needed by the hot
    path.  pandas must be imported before the pytz stub."""
    import pandas  # noqa: F401
    for name in ("matplotlib", "matplotlib.pyplot", "pytz"):
        if name not in sys.modules:
            try:
                __import__(name)
            except ImportError:
                sys.modules[name] = types.ModuleType(name)
    if isinstance(sys.modules.get("matplotlib"), types.ModuleType) and not hasattr(sys.modules["matplotlib"], "pyplot"):
        sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]


def _load_real(reference_dir: str, name: str):
    path = os.path.join(reference_dir, name + ".py")
    spec = importlib.util.spec_from_file_location("_reference_" + name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _loader(x, y, batch_size, as_dict=False):
    import torch.utils.data as D
    if as_dict:
        class _DictSet(D.Dataset):
            def __len__(self):
                return x.shape[0]

            def __getitem__(self, i):
                return {"image": x[i], "label": y[i], "name": "synthetic_%d" % i}
        return D.DataLoader(_DictSet(), batch_size=batch_size, shuffle=False)
    return D.DataLoader(D.TensorDataset(x, y), batch_size=batch_size, shuffle=False)


def offline_shims(reference_dir: str, n_train: int = 512, n_eval: int = 128, seed: int = 1226):
    """Insert synthetic-data modules into ``sys.modules`` under the names the parameter files import.
    Model / loss classes are the reference's own objects."""
    g = torch.Generator().manual_seed(seed)

    def tensors(n, shape, n_cls, multihot=False):
        x = torch.randn((n,) + shape, generator=g)
        if multihot:
            return x, (torch.rand((n, n_cls), generator=g) > 0.8).float()
        return x, torch.randint(0, n_cls, (n,), generator=g)

    def clone_public(real, name):
        shim = types.ModuleType(name)
        for k, v in vars(real).items():
            if not k.startswith("__"):
                setattr(shim, k, v)
        return shim

    # forest_data: get_data() -> dict of tensors (forest_data.py:30-71)
    real = _load_real(reference_dir, "forest_data")
    shim = clone_public(real, "forest_data")

    def get_data(*a, **k):
        x, y = tensors(n_train, (54,), 7)
        xv, yv = tensors(n_eval, (54,), 7)
        xt, yt = tensors(n_eval, (54,), 7)
        return {"inputs": x, "target": y, "inputs_valid": xv, "target_valid": yv, "inputs_test": xt, "target_test": yt}
    shim.get_data = get_data
    sys.modules["forest_data"] = shim

    # usps_data: loaders (usps_data.py:69-295)
    real = _load_real(reference_dir, "usps_data")
    shim = clone_public(real, "usps_data")

    def usps_train_valid(batch_size=1, **k):
        return _loader(*tensors(n_train, (1, 16, 16), 10), batch_size), _loader(*tensors(n_eval, (1, 16, 16), 10), batch_size)

    def usps_eval(batch_size=1, **k):
        return _loader(*tensors(n_eval, (1, 16, 16), 10), batch_size)
    shim.get_train_valid_loader = usps_train_valid
    shim.get_test_loader = usps_eval
    shim.get_mnist_loader = usps_eval
    shim.get_gan_loader = usps_eval
    sys.modules["usps_data"] = shim

    # cifar: the parameter files import a module that does not exist in the reference
    def cifar_module(name, n_cls):
        m = types.ModuleType(name)

        def get_norm(*a, **k):
            return np.zeros(3), np.ones(3)

        def get_train_valid_loader(batch_size=1, **k):
            tr = _loader(*tensors(n_train, (3, 32, 32), n_cls), batch_size)
            va = _loader(*tensors(n_eval, (3, 32, 32), n_cls), batch_size)
            na = _loader(*tensors(n_eval, (3, 32, 32), n_cls), batch_size)
            return tr, va, na

        def get_test_loader(batch_size=1, **k):
            return _loader(*tensors(n_eval, (3, 32, 32), n_cls), batch_size)
        m.get_norm, m.get_train_valid_loader, m.get_test_loader = get_norm, get_train_valid_loader, get_test_loader
        return m
    sys.modules["cifar10_data"] = cifar_module("cifar10_data", 10)
    sys.modules["cifar100_data"] = cifar_module("cifar100_data", 100)
    sys.modules["cifar_data"] = cifar_module("cifar_data", 100)

    # dcnn: private data sets + ImageNet downloads (dcnn.py:23-200,206,241,272)
    from torchvision import models as tvm
    for fn in ("alexnet", "resnet50", "vgg16_bn", "densenet121", "densenet161", "densenet201"):
        if hasattr(tvm, fn) and not getattr(getattr(tvm, fn), "_b200_offline", False):
            orig = getattr(tvm, fn)

            def make(orig):
                def offline(*a, **k):
                    return orig(weights=None)
                offline._b200_offline = True
                return offline
            setattr(tvm, fn, make(orig))
    real = sys.modules.get("dcnn") or _load_real(reference_dir, "dcnn")
    shim = clone_public(real, "dcnn")

    class SyntheticChest(torch.utils.data.Dataset):
        def __init__(self, *a, use="train", **k):
            n = n_train if use == "train" else n_eval
            self.x, self.y = tensors(min(n, 64), (3, 224, 224), 14, multihot=True)

        def __len__(self):
            return self.x.shape[0]

        def __getitem__(self, i):
            return {"image": self.x[i], "label": self.y[i], "name": "synthetic_%d" % i}
    for cname in ("ChestXray_Dataset", "CheXpert_Dataset", "MIMICCXR_Dataset"):
        setattr(shim, cname, SyntheticChest)
    sys.modules["dcnn"] = shim
