"""Generate tests/golden/*.npz by running the UNMODIFIED reference on CPU.

Run in the build container only (``/root/reference`` does not exist on the GPU
box):   python oracle/make_golden.py [--out tests/golden] [--only forest,usps,...]

Recipe: ``import pandas`` first, stub ``matplotlib``,
``matplotlib.pyplot`` and ``pytz`` (imported at module scope by opt.py:17,34 and
dcnn.py:9), put the reference on ``sys.path``, build the reference's own model
classes, feed seeded synthetic batches, and record what ``opt.HVPOperator`` /
``opt.OptWBoundEignVal.comp_rho`` / ``comp_gradrho`` return.  The stored arrays
are what the reference computed -- nothing from this repo is on that path except
the synthetic input generator.
"""
from __future__ import annotations

import argparse
import contextlib
import io
import os
import sys
import tempfile
import types

import numpy as np

REF = os.environ.get("OPTW_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def import_reference():
    import pandas  # noqa: F401  (must precede the pytz stub)
    for name in ("matplotlib", "matplotlib.pyplot", "pytz"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    if REF not in sys.path:
        sys.path.insert(0, REF)
    with contextlib.redirect_stdout(io.StringIO()):
        import opt  # noqa
    return opt


def ref_model(kind: str):
    """The reference's own classes, weights=None for torchvision backbones."""
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
    orig = {k: getattr(models, k) for k in ("vgg16_bn", "densenet121")}
    try:
        for k, fn in orig.items():
            setattr(models, k, (lambda f: (lambda *a, **kw: f(weights=None)))(fn))
        if kind == "chest_vgg":
            return dcnn.MyVggNet16_bn(14), dcnn.W_BCEWithLogitsLoss()
        if kind == "chest_densenet121":
            return dcnn.MyDenseNet121(14), dcnn.W_BCEWithLogitsLoss()
    finally:
        for k, fn in orig.items():
            setattr(models, k, fn)
    raise KeyError(kind)


def flat_state(model):
    import torch
    names, arrs = [], []
    for k, v in model.state_dict().items():
        names.append(k)
        arrs.append(v.detach().reshape(-1).double().numpy())
    return names, np.concatenate(arrs).astype(np.float32)


def run_case(opt, kind: str, batch: int, *, eps: float, max_pow_iter: int, lobpcg: bool = False,
             alpha=1, kfac_batch: int = 1, clip=None, second_batch: bool = True, seed: int = 1226):
    import torch
    sys.path.insert(0, ROOT)
    from optwboundeigenval_b200 import zoo

    torch.manual_seed(seed)
    np.random.seed(seed)
    model, loss = ref_model(kind)
    model.train()
    names, state0 = flat_state(model)
    x, y = zoo.synthetic_batch(kind, batch, seed)
    x2, y2 = zoo.synthetic_batch(kind, batch, seed + 7)
    out = {"state_names": np.array(names), "state0": state0, "x": x.numpy(), "y": y.numpy(),
           "x2": x2.numpy(), "y2": y2.numpy()}

    ndim = sum(p.numel() for p in model.parameters())
    v0 = torch.from_numpy(np.ones(ndim) / np.sqrt(ndim))
    gen = torch.Generator().manual_seed(seed + 3)
    vr = torch.randn(ndim, generator=gen, dtype=torch.float64)
    vr /= vr.norm()
    out["v_rand"] = vr.numpy()

    # --- HVPOperator on batch 1 (fresh operator per vGHv: it is one-shot)
    op = opt.HVPOperator(model, [x, y], loss, use_gpu=False)
    out["hv_v0"] = op.Hv(v0, storedGrad=True).numpy().astype(np.float32)
    out["grad"] = op.stored_grad.detach().numpy().astype(np.float32)
    out["hv_vrand"] = op.Hv(vr.numpy(), storedGrad=True).numpy().astype(np.float32)   # ndarray input path
    out["vghv_vrand"] = op.vGHv(vr, storedGrad=True).numpy().astype(np.float32)
    # module buffers after ONE train-mode forward (BN running stats side effect)
    model2, _ = ref_model(kind)
    model2.load_state_dict(dict(zip(names, _unflatten(state0, model2))))
    model2.train()
    op2 = opt.HVPOperator(model2, [x, y], loss, use_gpu=False)
    op2.Hv(v0, storedGrad=True)
    out["state_after_one_pass"] = flat_state(model2)[1]
    out["loss"] = np.array(_train_loss(kind, names, state0, x, y))

    # --- comp_rho / comp_gradrho through OptWBoundEignVal on a pristine copy of the weights
    model3, loss3 = ref_model(kind)
    model3.load_state_dict(dict(zip(names, _unflatten(state0, model3))))
    calls = []

    class Recording(opt.HVPOperator):           # records, does not alter, the reference operator
        def Hv(self, vec, storedGrad=False):
            r = super().Hv(vec, storedGrad)
            calls.append((torch.as_tensor(vec).detach().clone().double(), r.clone()))
            return r

    real = opt.HVPOperator
    opt.HVPOperator = Recording
    cwd = os.getcwd()
    tmp = tempfile.mkdtemp()
    os.makedirs(os.path.join(tmp, "logs"))
    os.chdir(tmp)
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            o = opt.OptWBoundEignVal(model3, loss3, torch.optim.SGD(model3.parameters(), lr=0.1), mu=0.01, K=0,
                                     pow_iter_eps=eps, max_pow_iter=max_pow_iter, use_gpu=False,
                                     ignore_bad_vals=False, verbose=True, header="golden", lobpcg=lobpcg,
                                     pow_iter_alpha=alpha, kfac_batch=kfac_batch, kfac_rand=False,
                                     gradg_clip=clip, batch_size=batch)
            i, rn, size = o.comp_rho([x, y])
        out["rho1_iters"] = np.array(i)
        out["rho1_rn"] = np.array(float(rn))
        out["rho1_rho"] = np.array(float(o.rho))
        out["rho1_norm"] = np.array(float(o.norm))
        out["rho1_v"] = o.v.numpy().copy()
        out["rho1_traj"] = _trajectory(calls)
        with contextlib.redirect_stdout(io.StringIO()):
            o.g = np.max([0.0, o.rho - o.K, o.Kmin - o.rho])
            o.comp_gradrho()
        out["rho1_gradrho"] = o.gradrho.numpy().astype(np.float32)
        out["rho1_gradf"] = o.hvp_op.stored_grad.detach().numpy().astype(np.float32)
        if second_batch:                        # warm start from batch 1's vector (opt.py:432)
            calls.clear()
            with contextlib.redirect_stdout(io.StringIO()):
                i2, rn2, _ = o.comp_rho([x2, y2])
            out["rho2_iters"] = np.array(i2)
            out["rho2_rho"] = np.array(float(o.rho))
            out["rho2_norm"] = np.array(float(o.norm))
            out["rho2_v"] = o.v.numpy().copy()
            out["rho2_traj"] = _trajectory(calls)
        with open(os.path.join(tmp, "logs", os.path.basename(o.verbose_log_file))) as fh:
            out["verbose_log"] = np.array(fh.read())
    finally:
        os.chdir(cwd)
        opt.HVPOperator = real
    return out


def _train_loss(kind, names, state0, x, y):
    import torch
    m, l = ref_model(kind)
    m.load_state_dict(dict(zip(names, _unflatten(state0, m))))
    m.train()
    with torch.no_grad():
        return float(l(m(x), y))


def _unflatten(flat, model):
    import torch
    res, j = [], 0
    for k, v in model.state_dict().items():
        n = v.numel()
        res.append(torch.from_numpy(np.asarray(flat[j:j + n])).to(v.dtype).view(v.shape))
        j += n
    return res


def _trajectory(calls):
    """(i, lam, n, rn, vnn) per iteration, recomputed in fp64 from the recorded (v, Hv) pairs exactly as
    opt.py:455-464 does."""
    import torch
    rows, r_old = [], 0.0
    for i, (v, w) in enumerate(calls):
        lam = float(torch.dot(w, v))
        if lam < 0:
            lam, w = -lam, -w
        r = w - lam * v
        n = float(torch.norm(r))
        rn = float(min(torch.norm(r - r_old), torch.norm(r + r_old)))
        rows.append((i, lam, n, rn, float(torch.norm(w))))
        r_old = r
    return np.array(rows, dtype=np.float64)


CASES = {
    # name: kwargs  (sizes keep the fixtures a few MB in total)
    "forest": dict(kind="forest", batch=128, eps=1e-3, max_pow_iter=1000),
    "forest_lobpcg": dict(kind="forest", batch=128, eps=1e-3, max_pow_iter=1000, lobpcg=True,
                          alpha=lambda k: np.exp(-4 * k - 2), kfac_batch=2),
    "usps": dict(kind="usps", batch=64, eps=1e-3, max_pow_iter=30),
    "usps_lobpcg": dict(kind="usps", batch=64, eps=1e-3, max_pow_iter=1000, lobpcg=True,
                        alpha=lambda k: np.exp(-4 * k), kfac_batch=4),
    "cifar_densenet": dict(kind="cifar_densenet", batch=8, eps=5e-2, max_pow_iter=100, second_batch=False),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "tests", "golden"))
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    opt = import_reference()
    os.makedirs(args.out, exist_ok=True)
    only = [s for s in args.only.split(",") if s]
    for name, kw in CASES.items():
        if only and name not in only:
            continue
        res = run_case(opt, **kw)
        meta = {k: v for k, v in kw.items() if not callable(v)}
        res["meta"] = np.array(repr(meta))
        path = os.path.join(args.out, name + ".npz")
        np.savez_compressed(path, **res)
        print("%-16s -> %s  (%.2f MB)  rho=%g iters=%d" % (
            name, path, os.path.getsize(path) / 1e6, float(res["rho1_rho"]), int(res["rho1_iters"])))


if __name__ == "__main__":
    main()
